// CUDA-core kernels: Dense + BN + act (linearTransform, autoencoder3D.py:56-70), the stride-1 ConvT layer
// (conv3DDec with strides=1, autoencoder3D.py:41-54,127-128; 0.2 % of the decoder FLOPs) and a direct stride-2
// ConvT used only as an on-device diagnostic for the tcgen05 kernel (A3D_IMPL_SIMT).
#include "cvt.cuh"
#include "internal.h"
#include "ptx.cuh"

namespace a3d {
namespace {

// ---------------------------------------------------------------------------------------------------------
// Dense(D -> 512) + bias + folded BN + activation.  One block per decode, one thread per output unit.
// out a0[n][512] is the NDHWC tensor [n,4,4,4,8] (tf.reshape at autoencoder3D.py:125 keeps the flat order).
template <int FMT>
__global__ void dense_kernel(const float* __restrict__ z, int D, const float* __restrict__ wd,
                             const float* __restrict__ bias, const float* __restrict__ scale,
                             const float* __restrict__ shift, uint16_t* __restrict__ a0, int units, int act) {
  extern __shared__ float zs[];
  const int64_t n = blockIdx.x;
  ptx::pdl_sync();     // z may be the imputation kernel's output; a0 was read by the previous call's chain
  for (int i = threadIdx.x; i < D; i += blockDim.x) zs[i] = z[n * D + i];
  __syncthreads();
  // one thread per output unit; the weight column is fetched 16 rows at a time (16 independent L2 loads in flight per
  // thread: the kernel is pure load latency -- with one load per FMA it took 15 us for 32 latents, 5 % of the call)
  for (int j = threadIdx.x; j < units; j += blockDim.x) {
    float acc = 0.f;
    int d = 0;
    for (; d + 16 <= D; d += 16) {
      float w[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) w[i] = __ldg(wd + (size_t)(d + i) * units + j);
#pragma unroll
      for (int i = 0; i < 16; ++i) acc = fmaf(zs[d + i], w[i], acc);
    }
    for (; d < D; ++d) acc = fmaf(zs[d], wd[(size_t)d * units + j], acc);
    acc += bias[j];
    a0[n * units + j] = from_f32<FMT>(apply_act(acc * scale[j] + shift[j], act));
  }
}

// ---------------------------------------------------------------------------------------------------------
// ConvT3D k=4 s=1 'same' (pad 1 before / 2 after): out[o, co] = sum_{t, ci} in[o - t + 1, ci] * W[t, co, ci].
// 4^3 x 8 -> 4^3 x 512.  Block = NB decodes; thread = 2 adjacent output channels x NB decodes (register tile); the
// input voxels sit in shared memory as [pos][ci][NB] so one (pos, ci) is NB/4 broadcast LDS.128; weights [tap][ci][co]
// are read once per NB decodes as packed 16-bit pairs.
template <int FMT, int NB>
__global__ void __launch_bounds__(256)
convt_s1_kernel(const uint16_t* __restrict__ a0, const uint16_t* __restrict__ w_tco, const float* __restrict__ scale,
                const float* __restrict__ shift, uint16_t* __restrict__ a1, int64_t n_total, int act) {
  constexpr int CIN = 8, COUT = 512;
  __shared__ __align__(16) float xs[64 * CIN * NB];  // [pos][ci][b]
  const int64_t n0 = (int64_t)blockIdx.x * NB;
  for (int i = threadIdx.x; i < NB * 64 * CIN; i += blockDim.x) {
    const int b = i / (64 * CIN), pc = i % (64 * CIN);
    const int64_t n = n0 + b;
    xs[pc * NB + b] = n < n_total ? to_f32<FMT>(a0[n * (64 * CIN) + pc]) : 0.f;
  }
  __syncthreads();
  const int co = threadIdx.x * 2;
  const float sc0 = scale[co], sc1 = scale[co + 1], sh0 = shift[co], sh1 = shift[co + 1];
  const uint32_t* w32 = reinterpret_cast<const uint32_t*>(w_tco);
  for (int o = 0; o < 64; ++o) {
    const int od = o >> 4, oh = (o >> 2) & 3, ow = o & 3;
    float acc0[NB], acc1[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) { acc0[b] = 0.f; acc1[b] = 0.f; }
    for (int td = 0; td < 4; ++td) {
      const int id = od - td + 1;
      if (id < 0 || id > 3) continue;
      for (int th = 0; th < 4; ++th) {
        const int ih = oh - th + 1;
        if (ih < 0 || ih > 3) continue;
        for (int tw = 0; tw < 4; ++tw) {
          const int iw = ow - tw + 1;
          if (iw < 0 || iw > 3) continue;
          const int tap = (td * 4 + th) * 4 + tw;
          const int ipos = (id * 4 + ih) * 4 + iw;
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci) {
            const float2 wv = unpack2<FMT>(__ldg(w32 + (((size_t)tap * CIN + ci) * COUT + co) / 2));
            const float4* xp = reinterpret_cast<const float4*>(xs + (ipos * CIN + ci) * NB);
#pragma unroll
            for (int q = 0; q < NB / 4; ++q) {
              const float4 x = xp[q];
              acc0[4 * q + 0] = fmaf(x.x, wv.x, acc0[4 * q + 0]); acc1[4 * q + 0] = fmaf(x.x, wv.y, acc1[4 * q + 0]);
              acc0[4 * q + 1] = fmaf(x.y, wv.x, acc0[4 * q + 1]); acc1[4 * q + 1] = fmaf(x.y, wv.y, acc1[4 * q + 1]);
              acc0[4 * q + 2] = fmaf(x.z, wv.x, acc0[4 * q + 2]); acc1[4 * q + 2] = fmaf(x.z, wv.y, acc1[4 * q + 2]);
              acc0[4 * q + 3] = fmaf(x.w, wv.x, acc0[4 * q + 3]); acc1[4 * q + 3] = fmaf(x.w, wv.y, acc1[4 * q + 3]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const int64_t n = n0 + b;
      if (n < n_total) {
        const uint32_t v = (uint32_t)from_f32<FMT>(apply_act(acc0[b] * sc0 + sh0, act)) |
                           ((uint32_t)from_f32<FMT>(apply_act(acc1[b] * sc1 + sh1, act)) << 16);
        *reinterpret_cast<uint32_t*>(a1 + (n * 64 + o) * COUT + co) = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Direct stride-2 ConvT (diagnostic).  One block per (decode, output voxel group); thread = output channel.
template <int FMT>
__global__ void convt_s2_simt_kernel(const uint16_t* __restrict__ in, const uint16_t* __restrict__ w_tco,
                                     const float* __restrict__ scale, const float* __restrict__ shift,
                                     uint16_t* __restrict__ out, int win, int cin, int cout, int act,
                                     int64_t total_vox) {
  extern __shared__ float xs[];  // [8 taps][cin]
  const int od_ = 2 * win;
  for (int64_t v = blockIdx.x; v < total_vox; v += gridDim.x) {
    const int ow = v % od_, oh = (v / od_) % od_, od = (v / ((int64_t)od_ * od_)) % od_;
    const int64_t n = v / ((int64_t)od_ * od_ * od_);
    int taps[8];
    __syncthreads();
    for (int s = 0; s < 8; ++s) {
      const int sd = s >> 2, sh = (s >> 1) & 1, sw = s & 1;
      const int pd = od & 1, ph = oh & 1, pw = ow & 1;
      const int id = (od >> 1) + sd - 1 + pd, ih = (oh >> 1) + sh - 1 + ph, iw = (ow >> 1) + sw - 1 + pw;
      // tap = p + 1 - 2*delta, delta = s - 1 + p
      const int td = pd + 1 - 2 * (sd - 1 + pd), th = ph + 1 - 2 * (sh - 1 + ph), tw = pw + 1 - 2 * (sw - 1 + pw);
      const bool ok = id >= 0 && id < win && ih >= 0 && ih < win && iw >= 0 && iw < win;
      taps[s] = ok ? (td * 4 + th) * 4 + tw : -1;
      for (int c = threadIdx.x; c < cin; c += blockDim.x)
        xs[s * cin + c] = ok ? to_f32<FMT>(in[((((size_t)n * win + id) * win + ih) * win + iw) * cin + c]) : 0.f;
    }
    __syncthreads();
    for (int co = threadIdx.x; co < cout; co += blockDim.x) {
      float acc = 0.f;
      for (int s = 0; s < 8; ++s) {
        if (taps[s] < 0) continue;
        const uint16_t* wp = w_tco + (size_t)taps[s] * cin * cout + co;
        for (int c = 0; c < cin; ++c) acc = fmaf(xs[s * cin + c], to_f32<FMT>(wp[(size_t)c * cout]), acc);
      }
      out[(size_t)v * cout + co] = from_f32<FMT>(apply_act(acc * scale[co] + shift[co], act));
    }
  }
}

template <int FMT>
__global__ void to_f32_kernel(const uint16_t* __restrict__ src, float* __restrict__ dst, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = to_f32<FMT>(src[i]);
}

}  // namespace

int launch_dense_l1(const float* z, int64_t n, int D, const float* wd, const float* bd, const float* s0,
                    const float* h0, void* a0, const void* w1_tco, const float* s1, const float* h1, void* a1,
                    int fmt, int act, bool do_s1, cudaStream_t st, int64_t* launches) {
  if (n <= 0) return A3D_OK;
  constexpr int NB = 16;
  const int blocks1 = (int)((n + NB - 1) / NB);
  if (fmt == A3D_DTYPE_F16) {
    A3D_CUDA_OK(launch_chain(dense_kernel<A3D_DTYPE_F16>, dim3((unsigned)n), dim3(512), D * sizeof(float), st, 1, z, D, wd, bd,
                             s0, h0, (uint16_t*)a0, 512, act));
    if (do_s1) convt_s1_kernel<A3D_DTYPE_F16, NB><<<blocks1, 256, 0, st>>>((const uint16_t*)a0, (const uint16_t*)w1_tco, s1, h1,
                                                               (uint16_t*)a1, n, act);
  } else {
    A3D_CUDA_OK(launch_chain(dense_kernel<A3D_DTYPE_BF16>, dim3((unsigned)n), dim3(512), D * sizeof(float), st, 1, z, D, wd, bd,
                             s0, h0, (uint16_t*)a0, 512, act));
    if (do_s1) convt_s1_kernel<A3D_DTYPE_BF16, NB><<<blocks1, 256, 0, st>>>((const uint16_t*)a0, (const uint16_t*)w1_tco, s1, h1,
                                                                (uint16_t*)a1, n, act);
  }
  A3D_CUDA_OK(cudaGetLastError());
  if (launches) *launches += do_s1 ? 2 : 1;
  return A3D_OK;
}

int launch_convt_s2_simt(const ConvLayer& L, const void* in, void* out, int64_t n, int fmt, int act, cudaStream_t st,
                         int64_t* launches) {
  const int od = 2 * L.win;
  const int64_t total = n * od * od * od;
  const int grid = (int)(total < 148 * 16 ? total : 148 * 16);
  const int threads = L.cout < 256 ? (L.cout < 64 ? 64 : L.cout) : 256;
  const size_t smem = (size_t)8 * L.cin * sizeof(float);
  if (fmt == A3D_DTYPE_F16)
    convt_s2_simt_kernel<A3D_DTYPE_F16><<<grid, threads, smem, st>>>((const uint16_t*)in, (const uint16_t*)L.wgt_tco, L.scale,
                                                                    L.shift, (uint16_t*)out, L.win, L.cin, L.cout, act, total);
  else
    convt_s2_simt_kernel<A3D_DTYPE_BF16><<<grid, threads, smem, st>>>((const uint16_t*)in, (const uint16_t*)L.wgt_tco, L.scale,
                                                                     L.shift, (uint16_t*)out, L.win, L.cin, L.cout, act, total);
  A3D_CUDA_OK(cudaGetLastError());
  if (launches) ++*launches;
  return A3D_OK;
}

int launch_to_f32(const void* src, float* dst, int64_t n, int fmt, cudaStream_t st) {
  const int grid = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  if (fmt == A3D_DTYPE_F16)
    to_f32_kernel<A3D_DTYPE_F16><<<grid, 256, 0, st>>>((const uint16_t*)src, dst, n);
  else
    to_f32_kernel<A3D_DTYPE_BF16><<<grid, 256, 0, st>>>((const uint16_t*)src, dst, n);
  A3D_CUDA_OK(cudaGetLastError());
  return A3D_OK;
}

}  // namespace a3d
