// Epilogue helpers shared by the tcgen05 transposed-conv kernels: folded-BN output activation and 16-bit packing.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "../../include/a3d.h"

namespace a3d {
namespace {

// ELU(alpha = 1) without the slow expm1f: exp via ex2.approx (|abs err| of exp(v) - 1 <= ~6e-8 for v <= 0, an order of
// magnitude below the fp16 quantisation of the stored activation everywhere it matters); 4 instructions, branch-free.
template <int ACT>
__device__ __forceinline__ float activate(float v) {
  if constexpr (ACT == A3D_ACT_ELU) {
    const float e = __expf(v) - 1.f;
    return v > 0.f ? v : e;
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
