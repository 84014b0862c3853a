// Epilogue helpers shared by the tcgen05 transposed-conv kernels: folded-BN output activation and 16-bit packing.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "../../include/a3d.h"

namespace a3d {
namespace {

// ELU(alpha = 1) without the slow expm1f, branch- AND predicate-free: max(v, 2^min(v log2e, 0) - 1) is v for v > 0 (the
// exponential term is then 0) and exp(v) - 1 below (exp(v) - 1 >= v everywhere).  ex2.approx.ftz: 2^-22 relative error,
// i.e. <= 2.4e-7 absolute on exp(v) <= 1, far below the fp16 quantisation of the stored activation.  __expf(v) compiled
// to a predicated denormal-scaling sequence whose single predicate register serialised the whole epilogue (ncu source
// page, round 2): FMUL, FMNMX, MUFU.EX2, FADD, FMNMX instead.
template <int ACT>
__device__ __forceinline__ float activate(float v) {
  if constexpr (ACT == A3D_ACT_ELU) {
    float t = fminf(v * 1.4426950408889634f, 0.f);
    asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(t));
    return fmaxf(v, t - 1.f);
  } else if constexpr (ACT == A3D_ACT_RELU) {
    return fmaxf(v, 0.f);
  } else if constexpr (ACT == A3D_ACT_LRELU) {
    return v > 0.f ? v : 0.3f * v;
  } else {
    return v;
  }
}

template <int FMT>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if constexpr (FMT == A3D_DTYPE_F16) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  } else {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
}

}  // namespace
}  // namespace a3d
