// CUDA-core kernels of the image encoder (Darknet19 + head2D, src/net_core/darknet.py:83-168):
//   conv2d_first_pool -- Conv2D(3 -> 32, k3, 'same') + BN + act + MaxPool2D(2,2) on the fp32 NHWC image (darknet.py:99-100):
//                        K = 27 is too thin for a tcgen05 tile; the layer is 1.6 % of the encoder's MACs
//   maxpool2d         -- stand-alone MaxPool2D(2, 2) on 16-bit NHWC (only used when a pool does not follow a conv)
//   global_pool       -- tf.reduce_max / reduce_mean over (H, W), darknet.py:158-163
//   import / export   -- fp32 <-> 16-bit NHWC copies with channel padding between user buffers and the arena
//   split_sample      -- mean / clipped log-variance split + sampling(), nolbo.py:869-875, function.py:35-38
#include "cvt.cuh"
#include "epilogue.cuh"
#include "internal.h"
#include "ptx.cuh"
#include "philox.cuh"

namespace a3d {
namespace {

template <int ACT>
__device__ __forceinline__ float act2d(float v) {
  if constexpr (ACT == A3D_ACT_LRELU01) return v > 0.f ? v : 0.1f * v;
  else return activate<ACT>(v);
}

// Block = 128 pooled pixels x 2 channel halves (warps 0-3: channels 0-15, warps 4-7: 16-31).  A thread holds the
// 4 x 4 x 3 input patch of its pooled pixel and 4 x 16 accumulators; weights [27][32] come from shared memory as
// warp-uniform 128-bit broadcasts (16 FMAs per load).
template <int FMT, int ACT>
__global__ void __launch_bounds__(256)
conv2d_first_pool_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ scale,
                         const float* __restrict__ shift, uint16_t* __restrict__ out, int64_t n_pooled, int H, int W,
                         int cout_pad) {
  __shared__ __align__(16) float ws[27 * 32];
  __shared__ float ss[32], sh[32];
  for (int i = threadIdx.x; i < 27 * 32; i += 256) ws[i] = w[i];
  if (threadIdx.x < 32) { ss[threadIdx.x] = scale[threadIdx.x]; sh[threadIdx.x] = shift[threadIdx.x]; }
  __syncthreads();
  const int half = threadIdx.x >> 7;
  const int64_t q = (int64_t)blockIdx.x * 128 + (threadIdx.x & 127);
  if (q >= n_pooled) return;
  const int Wq = W >> 1, Hq = H >> 1;
  const int wq = (int)(q % Wq), hq = (int)((q / Wq) % Hq);
  const int64_t img = q / ((int64_t)Wq * Hq);
  const float* base = in + img * (int64_t)H * W * 3;
  float patch[4][4][3];
#pragma unroll
  for (int y = 0; y < 4; ++y) {
    const int iy = 2 * hq - 1 + y;
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const int ix = 2 * wq - 1 + x;
      const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
      const float* p = base + ((int64_t)iy * W + ix) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) patch[y][x][c] = ok ? __ldg(p + c) : 0.f;
    }
  }
  float acc[4][16];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[p][c] = 0.f;
#pragma unroll
  for (int dy = 0; dy < 3; ++dy)
#pragma unroll
    for (int dx = 0; dx < 3; ++dx)
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        const float4* wv = reinterpret_cast<const float4*>(&ws[((dy * 3 + dx) * 3 + ci) * 32 + half * 16]);
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const float4 wq4 = wv[c4];
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const float x = patch[(p >> 1) + dy][(p & 1) + dx][ci];
            acc[p][c4 * 4 + 0] = fmaf(x, wq4.x, acc[p][c4 * 4 + 0]);
            acc[p][c4 * 4 + 1] = fmaf(x, wq4.y, acc[p][c4 * 4 + 1]);
            acc[p][c4 * 4 + 2] = fmaf(x, wq4.z, acc[p][c4 * 4 + 2]);
            acc[p][c4 * 4 + 3] = fmaf(x, wq4.w, acc[p][c4 * 4 + 3]);
          }
        }
      }
  uint32_t o[8];
#pragma unroll
  for (int c = 0; c < 16; c += 2) {
    float m[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float s = ss[half * 16 + c + e], t = sh[half * 16 + c + e];
      float v = act2d<ACT>(fmaf(acc[0][c + e], s, t));
      v = fmaxf(v, act2d<ACT>(fmaf(acc[1][c + e], s, t)));
      v = fmaxf(v, act2d<ACT>(fmaf(acc[2][c + e], s, t)));
      v = fmaxf(v, act2d<ACT>(fmaf(acc[3][c + e], s, t)));
      m[e] = v;
    }
    o[c >> 1] = pack2<FMT>(m[0], m[1]);
  }
  uint4* dst = reinterpret_cast<uint4*>(out + q * cout_pad + half * 16);
  dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
  dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
}

// thread = (pooled pixel, 8-channel group)
template <int FMT>
__global__ void __launch_bounds__(256)
maxpool2d_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, int64_t total, int H, int W, int C) {
  ptx::pdl_sync();
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int cg = C >> 3, Wq = W >> 1, Hq = H >> 1;
  const int c8 = (int)(i % cg);
  const int64_t q = i / cg;
  const int wq = (int)(q % Wq), hq = (int)((q / Wq) % Hq);
  const int64_t img = q / ((int64_t)Wq * Hq);
  const uint16_t* p = in + ((img * H + 2 * hq) * W + 2 * wq) * (int64_t)C + c8 * 8;
  const uint4 a = *reinterpret_cast<const uint4*>(p), b = *reinterpret_cast<const uint4*>(p + C);
  const uint4 c = *reinterpret_cast<const uint4*>(p + (int64_t)W * C), d = *reinterpret_cast<const uint4*>(p + (int64_t)W * C + C);
  const uint32_t av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w}, cv[4] = {c.x, c.y, c.z, c.w},
                 dv[4] = {d.x, d.y, d.z, d.w};
  uint32_t r[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 fa = unpack2<FMT>(av[e]), fb = unpack2<FMT>(bv[e]), fc = unpack2<FMT>(cv[e]), fd = unpack2<FMT>(dv[e]);
    r[e] = pack2<FMT>(fmaxf(fmaxf(fa.x, fb.x), fmaxf(fc.x, fd.x)), fmaxf(fmaxf(fa.y, fb.y), fmaxf(fc.y, fd.y)));
  }
  *reinterpret_cast<uint4*>(out + q * C + c8 * 8) = make_uint4(r[0], r[1], r[2], r[3]);
}

// in [n, HW, C] fp32 -> out [n, C]; thread = (image, channel); reduce_mean sums in index order like a serial loop
__global__ void __launch_bounds__(128)
global_pool_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t total, int HW, int C, int is_max) {
  ptx::pdl_sync();
  const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const int64_t img = i / C;
  const float* p = in + img * HW * (int64_t)C + c;
  float m = p[0];
  if (is_max) {
    for (int k = 1; k < HW; ++k) m = fmaxf(m, p[(int64_t)k * C]);
  } else {
    for (int k = 1; k < HW; ++k) m += p[(int64_t)k * C];
    m /= (float)HW;
  }
  out[i] = m;
}

template <int FMT>
__global__ void __launch_bounds__(256)
import_kernel(const void* __restrict__ in, int in_is_f32, uint16_t* __restrict__ out, int64_t total, int C, int C_pad) {
  ptx::pdl_sync();
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C_pad);
  const int64_t p = i / C_pad;
  uint16_t v = from_f32<FMT>(0.f);
  if (c < C) v = in_is_f32 ? from_f32<FMT>(reinterpret_cast<const float*>(in)[p * C + c])
                           : reinterpret_cast<const uint16_t*>(in)[p * C + c];
  out[i] = v;
}

template <int FMT>
__global__ void __launch_bounds__(256)
export_kernel(const uint16_t* __restrict__ in, void* __restrict__ out, int out_is_f32, int64_t total, int C, int C_pad) {
  ptx::pdl_sync();
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const int64_t p = i / C;
  const uint16_t v = in[p * C_pad + c];
  if (out_is_f32) reinterpret_cast<float*>(out)[i] = to_f32<FMT>(v);
  else reinterpret_cast<uint16_t*>(out)[i] = v;
}

// thread = (image, quad of latent dims)
__global__ void __launch_bounds__(128)
split_sample_kernel(const float* __restrict__ enc_out, int64_t n, int D, int out_stride, float clip, int seed_enable,
                    uint64_t seed, uint64_t obj_offset, float* __restrict__ mean, float* __restrict__ logvar,
                    float* __restrict__ z) {
  ptx::pdl_sync();
  const int nq = (D + 3) >> 2;
  const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (i >= n * nq) return;
  const int q = (int)(i % nq);
  const int64_t b = i / nq;
  float nrm[4] = {0.f, 0.f, 0.f, 0.f};
  if (seed_enable) {
    const uint64_t obj = obj_offset + (uint64_t)b;
    const uint4 w = philox4x32_10(make_uint4((uint32_t)q, 0x5A4D504Cu, (uint32_t)obj, (uint32_t)(obj >> 32)),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    box_muller(w.x, w.y, nrm[0], nrm[1]);
    box_muller(w.z, w.w, nrm[2], nrm[3]);
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int d = q * 4 + e;
    if (d >= D) break;
    const float mu = enc_out[b * out_stride + d];
    const float lv = fminf(fmaxf(enc_out[b * out_stride + D + d], -clip), clip);   // nolbo.py:873
    if (mean) mean[b * D + d] = mu;
    if (logvar) logvar[b * D + d] = lv;
    if (z) z[b * D + d] = mu + sqrtf(expf(lv)) * nrm[e];                           // function.py:37-38
  }
}

__global__ void __launch_bounds__(256) sigmoid_kernel(float* __restrict__ x, int64_t n) {
  ptx::pdl_sync();
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) x[i] = 1.f / (1.f + __expf(-x[i]));
}

}  // namespace

int launch_sigmoid_inplace(float* x, int64_t n, cudaStream_t st, int64_t* launches) {
  if (n <= 0) return A3D_OK;
  A3D_CUDA_OK(launch_chain(sigmoid_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, 1, x, n));
  if (launches) ++*launches;
  return A3D_OK;
}

int launch_conv2d_first_pool(const float* in, const float* w27x32, const float* scale, const float* shift, void* out,
                             int64_t n, int H, int W, int cout_pad, int fmt, int act, cudaStream_t st,
                             int64_t* launches) {
  if (n <= 0) return A3D_OK;
  const int64_t n_pooled = n * (H / 2) * (W / 2);
  const unsigned grid = (unsigned)((n_pooled + 127) / 128);
  uint16_t* o = reinterpret_cast<uint16_t*>(out);
#define A3D_FIRST(FMT_, ACT_) conv2d_first_pool_kernel<FMT_, ACT_><<<grid, 256, 0, st>>>(in, w27x32, scale, shift, o, n_pooled, H, W, cout_pad)
#define A3D_FIRST_ACT(FMT_)                                          \
  switch (act) {                                                     \
    case A3D_ACT_ELU: A3D_FIRST(FMT_, A3D_ACT_ELU); break;           \
    case A3D_ACT_RELU: A3D_FIRST(FMT_, A3D_ACT_RELU); break;         \
    case A3D_ACT_LRELU: A3D_FIRST(FMT_, A3D_ACT_LRELU); break;       \
    case A3D_ACT_LRELU01: A3D_FIRST(FMT_, A3D_ACT_LRELU01); break;   \
    case A3D_ACT_NONE: A3D_FIRST(FMT_, A3D_ACT_NONE); break;         \
    default: set_error("conv2d_first: unsupported activation %d", act); return A3D_ERR_INVALID; \
  }
  if (fmt == A3D_DTYPE_F16) { A3D_FIRST_ACT(A3D_DTYPE_F16) } else { A3D_FIRST_ACT(A3D_DTYPE_BF16) }
#undef A3D_FIRST_ACT
#undef A3D_FIRST
  A3D_CUDA_OK(cudaGetLastError());
  if (launches) ++*launches;
  return A3D_OK;
}

int launch_maxpool2d(const void* in, void* out, int64_t n, int H, int W, int C, int fmt, cudaStream_t st,
                     int64_t* launches) {
  if (n <= 0) return A3D_OK;
  const int64_t total = n * (H / 2) * (W / 2) * (C / 8);
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (fmt == A3D_DTYPE_F16)
    A3D_CUDA_OK(launch_chain(maxpool2d_kernel<A3D_DTYPE_F16>, dim3(grid), dim3(256), 0, st, 1, (const uint16_t*)in, (uint16_t*)out, total, H, W, C));
  else
    A3D_CUDA_OK(launch_chain(maxpool2d_kernel<A3D_DTYPE_BF16>, dim3(grid), dim3(256), 0, st, 1, (const uint16_t*)in, (uint16_t*)out, total, H, W, C));
  if (launches) ++*launches;
  return A3D_OK;
}

int launch_global_pool(const float* in, float* out, int64_t n, int HW, int C, int is_max, cudaStream_t st,
                       int64_t* launches) {
  if (n <= 0) return A3D_OK;
  const int64_t total = n * C;
  A3D_CUDA_OK(launch_chain(global_pool_kernel, dim3((unsigned)((total + 127) / 128)), dim3(128), 0, st, 1, in, out, total, HW, C, is_max));
  if (launches) ++*launches;
  return A3D_OK;
}

// uint8 image bytes -> fp32 * scale (the loader's `image / 255.`, src/dataset_loader/pascal3D.py:242, moved onto the
// device so that the host -> device copy carries 1 byte per sample instead of 4): 16 bytes in, 64 bytes out per thread
__global__ void __launch_bounds__(256) u8_to_f32_kernel(const uint8_t* __restrict__ in, float* __restrict__ out,
                                                        int64_t total, float scale) {
  ptx::pdl_sync();
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
  if (i + 16 <= total) {
    const uint4 v = *reinterpret_cast<const uint4*>(in + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q)
      reinterpret_cast<float4*>(out + i)[q] =
          make_float4((float)(w[q] & 255u) * scale, (float)((w[q] >> 8) & 255u) * scale,
                      (float)((w[q] >> 16) & 255u) * scale, (float)(w[q] >> 24) * scale);
  } else {
    for (int64_t j = i; j < total; ++j) out[j] = (float)in[j] * scale;
  }
}

int launch_u8_to_f32(const uint8_t* in, float* out, int64_t total, float scale, cudaStream_t st, int64_t* launches) {
  if (total <= 0) return A3D_OK;
  const int64_t threads = (total + 15) / 16;
  A3D_CUDA_OK(launch_chain(u8_to_f32_kernel, dim3((unsigned)((threads + 255) / 256)), dim3(256), 0, st, 1, in, out, total, scale));
  if (launches) ++*launches;
  return A3D_OK;
}

int launch_import_nhwc(const void* in, int in_is_f32, void* out, int64_t pixels, int C, int C_pad, int fmt,
                       cudaStream_t st, int64_t* launches) {
  if (pixels <= 0) return A3D_OK;
  const int64_t total = pixels * C_pad;
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (fmt == A3D_DTYPE_F16) A3D_CUDA_OK(launch_chain(import_kernel<A3D_DTYPE_F16>, dim3(grid), dim3(256), 0, st, 1, in, in_is_f32, (uint16_t*)out, total, C, C_pad));
  else A3D_CUDA_OK(launch_chain(import_kernel<A3D_DTYPE_BF16>, dim3(grid), dim3(256), 0, st, 1, in, in_is_f32, (uint16_t*)out, total, C, C_pad));
  if (launches) ++*launches;
  return A3D_OK;
}

int launch_export_nhwc(const void* in, void* out, int out_is_f32, int64_t pixels, int C, int C_pad, int fmt,
                       cudaStream_t st, int64_t* launches) {
  if (pixels <= 0) return A3D_OK;
  const int64_t total = pixels * C;
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (fmt == A3D_DTYPE_F16) A3D_CUDA_OK(launch_chain(export_kernel<A3D_DTYPE_F16>, dim3(grid), dim3(256), 0, st, 1, (const uint16_t*)in, out, out_is_f32, total, C, C_pad));
  else A3D_CUDA_OK(launch_chain(export_kernel<A3D_DTYPE_BF16>, dim3(grid), dim3(256), 0, st, 1, (const uint16_t*)in, out, out_is_f32, total, C, C_pad));
  if (launches) ++*launches;
  return A3D_OK;
}

int launch_split_sample(const float* enc_out, int64_t n, int D, int out_stride, float clip, int seed_enable,
                        uint64_t seed, uint64_t obj_offset, float* mean, float* logvar, float* z, cudaStream_t st,
                        int64_t* launches) {
  if (n <= 0) return A3D_OK;
  const int64_t total = n * ((D + 3) / 4);
  A3D_CUDA_OK(launch_chain(split_sample_kernel, dim3((unsigned)((total + 127) / 128)), dim3(128), 0, st, 1, enc_out, n, D, out_stride,
                           clip, seed_enable, seed, obj_offset, mean, logvar, z));
  if (launches) ++*launches;
  return A3D_OK;
}

}  // namespace a3d
