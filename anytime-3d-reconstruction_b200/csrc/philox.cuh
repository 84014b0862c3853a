// Philox4x32-10 counter-based generator + Box-Muller, shared by the imputation sampler (aux_kernels.cu) and the
// encoder's latent sampler (enc2d_kernels.cu).  Pinned by the Random123 known-answer vectors in tests/test_oracle.py.
#pragma once
#include <stdint.h>

namespace a3d {
namespace {

// Philox4x32-10 (Salmon et al. 2011).  Counter (c0..c3), key (k0, k1).
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ void box_muller(uint32_t w0, uint32_t w1, float& n0, float& n1) {
  const float u0 = fmaf((float)w0, 2.3283064365386963e-10f, 1.1641532182693481e-10f);  // (w + 0.5) * 2^-32
  const float u1 = fmaf((float)w1, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
  const float r = sqrtf(-2.f * logf(u0));
  float s, c;
  sincospif(2.f * u1, &s, &c);
  n0 = r * c;
  n1 = r * s;
}

}  // namespace
}  // namespace a3d
