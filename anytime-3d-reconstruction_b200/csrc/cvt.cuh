// 16-bit operand <-> fp32 conversion helpers shared by the kernels (fmt: A3D_DTYPE_F16 / A3D_DTYPE_BF16).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "../../include/a3d.h"

namespace a3d {

template <int FMT>
__device__ __forceinline__ float to_f32(uint16_t v) {
  if constexpr (FMT == A3D_DTYPE_F16)
    return __half2float(__ushort_as_half(v));
  else
    return __bfloat162float(__ushort_as_bfloat16(v));
}
template <int FMT>
__device__ __forceinline__ uint16_t from_f32(float v) {
  if constexpr (FMT == A3D_DTYPE_F16) {
    // saturate instead of overflowing to inf: activations are BN-normalised, this is a guard only
    v = fminf(fmaxf(v, -65504.f), 65504.f);
    return __half_as_ushort(__float2half_rn(v));
  } else {
    return __bfloat16_as_ushort(__float2bfloat16_rn(v));
  }
}
template <int FMT>
__device__ __forceinline__ float2 unpack2(uint32_t v) {
  if constexpr (FMT == A3D_DTYPE_F16) {
    return __half22float2(*reinterpret_cast<const __half2*>(&v));
  } else {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v));
  }
}
__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == A3D_ACT_ELU) return v > 0.f ? v : expm1f(v);
  if (act == A3D_ACT_RELU) return fmaxf(v, 0.f);
  if (act == A3D_ACT_LRELU) return v > 0.f ? v : 0.3f * v;
  return v;
}

}  // namespace a3d
