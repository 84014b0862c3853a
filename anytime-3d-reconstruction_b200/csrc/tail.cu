// Fused tail of the anytime path:
//   final Conv3DTranspose(64 -> 1, k4, s2, 'same', no bias, no BN) + tf.sigmoid   autoencoder3D.py:129-136
//   mean over the K post-sigmoid grids of an object                                 nolbo_test.py:167-177
//   yPred = (mean >= thr), TP / FP / FN against the bit-packed target               function.py:100-115
// One pass over the 32^3 x 64 activations of the K samples; nothing but the counts (and, on request, the fp32 mean
// grid) is written back.
//
// v1 (this file): CUDA-core gather form.  One thread per INPUT voxel j computes its 8 output parities
// out[2j + p] = sum_{delta, ci} x[j + delta, ci] * w5[t(p, delta), ci]; 27 neighbour vectors are read once and
// shared by the parities that use them.
#include "cvt.cuh"
#include "internal.h"

namespace a3d {
namespace {

constexpr int G4 = 32;   // input grid of the final layer
constexpr int C4 = 64;   // its channels
constexpr int GO = 64;   // output grid

template <int FMT>
__global__ void __launch_bounds__(256)
tail_simt_kernel(const uint16_t* __restrict__ a4, const float* __restrict__ w5 /*[64 taps][64 ci]*/, int K,
                 int final_sigmoid, const uint8_t* __restrict__ target_bits, float thr,
                 unsigned long long* __restrict__ counts, float* __restrict__ mean_prob, float gamma,
                 double* __restrict__ loss) {
  __shared__ float ws[64 * C4];
  for (int i = threadIdx.x; i < 64 * C4; i += blockDim.x) ws[i] = w5[i];
  __syncthreads();
  const int64_t b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;  // input voxel, w fastest
  const int jw = j & 31, jh = (j >> 5) & 31, jd = j >> 10;
  float psum[8];
#pragma unroll
  for (int p = 0; p < 8; ++p) psum[p] = 0.f;

  for (int k = 0; k < K; ++k) {
    const uint16_t* base = a4 + ((size_t)(b * K + k) * G4 * G4 * G4) * C4;
    float acc[8];
#pragma unroll
    for (int p = 0; p < 8; ++p) acc[p] = 0.f;
#pragma unroll
    for (int dd = -1; dd <= 1; ++dd) {
      const int id = jd + dd;
      if (id < 0 || id >= G4) continue;
#pragma unroll
      for (int dh = -1; dh <= 1; ++dh) {
        const int ih = jh + dh;
        if (ih < 0 || ih >= G4) continue;
#pragma unroll
        for (int dw = -1; dw <= 1; ++dw) {
          const int iw = jw + dw;
          if (iw < 0 || iw >= G4) continue;
          const uint4* xp = reinterpret_cast<const uint4*>(base + ((size_t)(id * G4 + ih) * G4 + iw) * C4);
#pragma unroll
          for (int c8 = 0; c8 < 8; ++c8) {
            const uint4 q = __ldg(xp + c8);
            const float2 f0 = unpack2<FMT>(q.x), f1 = unpack2<FMT>(q.y), f2 = unpack2<FMT>(q.z),
                         f3 = unpack2<FMT>(q.w);
            const float xv[8] = {f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, f3.x, f3.y};
#pragma unroll
            for (int p = 0; p < 8; ++p) {
              const int pd = p >> 2, ph = (p >> 1) & 1, pw = p & 1;
              // parity p uses delta in {p-1, p}; tap = p + 1 - 2*delta
              if ((dd != pd - 1 && dd != pd) || (dh != ph - 1 && dh != ph) || (dw != pw - 1 && dw != pw)) continue;
              const int tap = ((pd + 1 - 2 * dd) * 4 + (ph + 1 - 2 * dh)) * 4 + (pw + 1 - 2 * dw);
              const float4* wp = reinterpret_cast<const float4*>(ws + tap * C4 + c8 * 8);
              const float4 w0 = wp[0], w1 = wp[1];
              float a = acc[p];
              a = fmaf(xv[0], w0.x, a); a = fmaf(xv[1], w0.y, a); a = fmaf(xv[2], w0.z, a); a = fmaf(xv[3], w0.w, a);
              a = fmaf(xv[4], w1.x, a); a = fmaf(xv[5], w1.y, a); a = fmaf(xv[6], w1.z, a); a = fmaf(xv[7], w1.w, a);
              acc[p] = a;
            }
          }
        }
      }
    }
#pragma unroll
    for (int p = 0; p < 8; ++p) psum[p] += final_sigmoid ? 1.f / (1.f + expf(-acc[p])) : acc[p];
  }

  const float invk = 1.f / (float)K;
  int tp = 0, fp = 0, fn = 0;
  float lsum = 0.f;
#pragma unroll
  for (int p = 0; p < 8; p += 2) {
    const int pd = p >> 2, ph = (p >> 1) & 1;
    const int od = 2 * jd + pd, oh = 2 * jh + ph, ow = 2 * jw;
    const size_t v = ((size_t)od * GO + oh) * GO + ow;
    const float m0 = psum[p] * invk, m1 = psum[p + 1] * invk;
    A3D_DEV_CHECK(v + 1 < (size_t)A3D_VOXELS);
    if (mean_prob) *reinterpret_cast<float2*>(mean_prob + (size_t)b * A3D_VOXELS + v) = make_float2(m0, m1);
    if (target_bits) {
      const uint32_t byte = target_bits[(size_t)b * (A3D_VOXELS / 8) + (v >> 3)];
      const int t0 = (byte >> (v & 7)) & 1, t1 = (byte >> ((v & 7) + 1)) & 1;
      const int y0 = m0 >= thr, y1 = m1 >= thr;
      tp += (t0 & y0) + (t1 & y1);
      fp += ((1 - t0) & y0) + ((1 - t1) & y1);
      fn += (t0 & (1 - y0)) + (t1 & (1 - y1));
      if (loss) {
        const float c0 = fminf(fmaxf(m0, 1e-7f), 1.f - 1e-7f), c1 = fminf(fmaxf(m1, 1e-7f), 1.f - 1e-7f);
        lsum -= t0 ? gamma * logf(c0) : (1.f - gamma) * logf(1.f - c0);
        lsum -= t1 ? gamma * logf(c1) : (1.f - gamma) * logf(1.f - c1);
      }
    }
  }
  if (target_bits) {
    if (loss) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
      if ((threadIdx.x & 31) == 0) atomicAdd(loss + b, (double)lsum);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      tp += __shfl_xor_sync(0xffffffffu, tp, o);
      fp += __shfl_xor_sync(0xffffffffu, fp, o);
      fn += __shfl_xor_sync(0xffffffffu, fn, o);
    }
    if ((threadIdx.x & 31) == 0) {
      if (tp) atomicAdd(counts + b * 3 + 0, (unsigned long long)tp);
      if (fp) atomicAdd(counts + b * 3 + 1, (unsigned long long)fp);
      if (fn) atomicAdd(counts + b * 3 + 2, (unsigned long long)fn);
    }
  }
}

}  // namespace

int launch_tail(const void* a4, const float* w5, int64_t B, int K, int fmt, int final_sigmoid,
                const uint8_t* target_bits, float thr, unsigned long long* counts, float* mean_prob, float gamma,
                double* loss, cudaStream_t st, int64_t* launches) {
  if (B <= 0) return A3D_OK;
  dim3 grid(G4 * G4 * G4 / 256, (unsigned)B);
  if (fmt == A3D_DTYPE_F16)
    tail_simt_kernel<A3D_DTYPE_F16><<<grid, 256, 0, st>>>((const uint16_t*)a4, w5, K, final_sigmoid, target_bits, thr,
                                                         counts, mean_prob, gamma, loss);
  else
    tail_simt_kernel<A3D_DTYPE_BF16><<<grid, 256, 0, st>>>((const uint16_t*)a4, w5, K, final_sigmoid, target_bits, thr,
                                                          counts, mean_prob, gamma, loss);
  A3D_CUDA_OK(cudaGetLastError());
  if (launches) ++*launches;
  return A3D_OK;
}

}  // namespace a3d
