// First Darknet19 layer on the tensor cores: Conv2D(3 -> 32, k3, 'same') + BN + act + MaxPool2D(2,2) on the fp32 NHWC image
// (src/net_core/darknet.py:99-100).
//
// K = 3 x 3 x 3 = 27 is too thin for a TMA-fed implicit GEMM (a 3-channel pixel is 12 bytes), so the A operand is built
// by the CTA itself in the SWIZZLE_64B K-major layout (64-byte rows of 27 + 5 zero columns; Swizzle<2,4,3>: 16-byte
// chunk c of row r lands at chunk c ^ ((r >> 1) & 3)) and multiplied against the resident 32 x 32 weight tile with
// tcgen05.mma (M = 128, N = 32, K = 16 x 2).  GEMM rows are POOLED pixels: each of the four positions of the 2 x 2 pool
// window gets its own A tile and its own 32-column TMEM block, so the pool is a per-thread max (no shuffles) taken before
// the BN shift and the activation.  The CUDA-core version (enc2d_kernels.cu) needed 864 FMAs per pixel and was 30 % of
// the encoder.  Several CTAs per SM hide the gather -> MMA -> epilogue dependency chain (128 TMEM columns each).
#include "epilogue.cuh"
#include "internal.h"
#include "ptx.cuh"

namespace a3d {
namespace {

template <int ACT>
__device__ __forceinline__ float act2d(float v) {
  if constexpr (ACT == A3D_ACT_LRELU01) return v > 0.f ? v : 0.1f * v;
  else return activate<ACT>(v);
}

struct FirstGeom {
  int H, W, lw, lh, tiles_w, tiles_h, total_tiles, n_images, cout_pad;   // lw / lh / tiles_*: brick of POOLED pixels
};

// Thread t owns POOLED pixel t of a (1 << lw) x (1 << lh) brick (x nt images): it gathers the 4 x 4 x 3 input patch of
// its 2 x 2 window once (12 loads per conv pixel instead of 27), writes one K-major row into each of FOUR A tiles
// (one per window position), and the four accumulators land in four 32-column TMEM blocks of the SAME lane -- so the
// max-pool is a per-thread max over four registers: no shuffles, and BN shift + activation run once per pooled value.
template <int FMT, int ACT>
__global__ void __launch_bounds__(128, 4)
conv2d_first_tc_kernel(const float* __restrict__ in, const uint16_t* __restrict__ w32x32, const float* __restrict__ scale,
                       const float* __restrict__ shift, uint16_t* __restrict__ out, FirstGeom g) {
  __shared__ __align__(1024) uint8_t sA[4][128 * 64];
  __shared__ __align__(1024) uint8_t sB[32 * 64];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float ss[32], sh[32];
  const int tid = threadIdx.x, warp = tid >> 5;
  {
    const int row = tid >> 2, c = tid & 3;   // 32 rows x 4 chunks of the weight tile [co][k]
    const uint4 v = *reinterpret_cast<const uint4*>(w32x32 + row * 32 + c * 8);
    *reinterpret_cast<uint4*>(sB + row * 64 + ((c ^ ((row >> 1) & 3)) << 4)) = v;
  }
  if (tid < 32) { ss[tid] = scale[tid]; sh[tid] = shift[tid]; }
  if (tid == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  if (warp == 0) { ptx::tmem_alloc<1>(&tmem_slot, 128); ptx::tmem_relinquish<1>(); }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  ptx::pdl_sync();   // programmatic dependent launch: the prologue above overlaps the previous kernel (ptx.cuh)
  constexpr uint32_t idesc = ptx::make_idesc_f16(128, 32, FMT);
  const uint32_t a_lo = ptx::sw128_desc_lo(ptx::smem_u32(&sA[0][0])), b_lo = ptx::sw128_desc_lo(ptx::smem_u32(sB));
  const int wi = tid & ((1 << g.lw) - 1), hi = (tid >> g.lw) & ((1 << g.lh) - 1), ni = tid >> (g.lw + g.lh);
  const int sw = (tid >> 1) & 3;
  const int Hq = g.H >> 1, Wq = g.W >> 1;
  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
    const int tw = tile % g.tiles_w, th = (tile / g.tiles_w) % g.tiles_h, nb = tile / (g.tiles_w * g.tiles_h);
    const int img = (nb << (7 - g.lw - g.lh)) + ni;
    const int hq = (th << g.lh) + hi, wq = (tw << g.lw) + wi;
    const bool valid = img < g.n_images && hq < Hq && wq < Wq;   // the power-of-two brick may overhang the pooled grid
    float patch[4][4][3];
    const float* base = in + (int64_t)img * g.H * g.W * 3;
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      const int iy = 2 * hq - 1 + y;
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const int ix = 2 * wq - 1 + x;
        const bool ok = valid && iy >= 0 && iy < g.H && ix >= 0 && ix < g.W;
        const float* p = base + ((int64_t)iy * g.W + ix) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) patch[y][x][c] = ok ? __ldg(p + c) : 0.f;
      }
    }
#pragma unroll
    for (int sub = 0; sub < 4; ++sub) {
      const int py = sub >> 1, px = sub & 1;
      uint32_t u[16];
#pragma unroll
      for (int i = 0; i < 14; ++i) {
        const int k0 = 2 * i, k1 = 2 * i + 1;
        const float v0 = patch[py + k0 / 9][px + (k0 / 3) % 3][k0 % 3];
        const float v1 = k1 < 27 ? patch[py + k1 / 9][px + (k1 / 3) % 3][k1 % 3] : 0.f;
        u[i] = pack2<FMT>(v0, v1);
      }
      u[14] = 0u; u[15] = 0u;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4*>(&sA[sub][tid * 64 + ((c ^ sw) << 4)]) = make_uint4(u[4 * c], u[4 * c + 1], u[4 * c + 2], u[4 * c + 3]);
    }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
#pragma unroll
        for (int sub = 0; sub < 4; ++sub) {
          const uint32_t a = a_lo + sub * ((128 * 64) >> 4);
          ptx::umma_f16<1>(tmem + sub * 32, ptx::sw64_desc(a), ptx::sw64_desc(b_lo), idesc, 0u);
          ptx::umma_f16<1>(tmem + sub * 32, ptx::sw64_desc(a + 2), ptx::sw64_desc(b_lo + 2), idesc, 1u);
        }
        ptx::umma_commit<1>(&bar);
      }
      __syncwarp();
    }
    ptx::mbar_wait(&bar, phase);
    phase ^= 1u;
    ptx::tc_fence_after();
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t acc[4][16];
#pragma unroll
      for (int sub = 0; sub < 4; ++sub) ptx::tmem_ld16(taddr + sub * 32 + half * 16, acc[sub]);
      ptx::tmem_ld_wait();
      uint32_t o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float y[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c = 2 * i + e;
          const float s = ss[half * 16 + c];
          // max over the window of s * a (then + shift, activation): valid for either sign of s, both maps are monotone
          float m = fmaxf(fmaxf(__uint_as_float(acc[0][c]) * s, __uint_as_float(acc[1][c]) * s),
                          fmaxf(__uint_as_float(acc[2][c]) * s, __uint_as_float(acc[3][c]) * s));
          y[e] = act2d<ACT>(m + sh[half * 16 + c]);
        }
        o[i] = pack2<FMT>(y[0], y[1]);
      }
      // stage the pooled 64-byte row in the (now free) first A tile, same 64-byte XOR swizzle
      *reinterpret_cast<uint4*>(&sA[0][tid * 64 + (((half * 2) ^ sw) << 4)]) = make_uint4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<uint4*>(&sA[0][tid * 64 + (((half * 2 + 1) ^ sw) << 4)]) = make_uint4(o[4], o[5], o[6], o[7]);
    }
    ptx::tc_fence_before();
    __syncwarp();
    {
      // coalesced write-out: 8 rows x 64 B per store instruction (full sectors) instead of 16 B per lane per row
      const int lane = tid & 31;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int rr = warp * 32 + 8 * j + (lane >> 2), ch = lane & 3;
        const int rwi = rr & ((1 << g.lw) - 1), rhi = (rr >> g.lw) & ((1 << g.lh) - 1), rni = rr >> (g.lw + g.lh);
        const int rimg = (nb << (7 - g.lw - g.lh)) + rni;
        const uint4 q = *reinterpret_cast<const uint4*>(&sA[0][rr * 64 + ((ch ^ ((rr >> 1) & 3)) << 4)]);
        const int rhq = (th << g.lh) + rhi, rwq = (tw << g.lw) + rwi;
        if (rimg < g.n_images && rhq < Hq && rwq < Wq) {
          const int64_t rp = ((int64_t)rimg * Hq + rhq) * Wq + rwq;
          *reinterpret_cast<uint4*>(out + rp * g.cout_pad + ch * 8) = q;
        }
      }
    }
    __syncwarp();   // the next tile's A rows of this warp overwrite the staging rows it just read
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<1>(tmem, 128);
}

}  // namespace

int launch_conv2d_first_tc(const float* in, const void* w32x32, const float* scale, const float* shift, void* out,
                           int64_t n, int H, int W, int cout_pad, int fmt, int act, int num_sms, cudaStream_t st,
                           int64_t* launches) {
  if (n <= 0) return A3D_OK;
  FirstGeom g;
  if ((H & 1) || (W & 1) || H < 2 || W < 2) { set_error("conv2d_first: needs even H, W >= 2"); return A3D_ERR_INVALID; }
  const int Wq = W / 2, Hq = H / 2;   // the brick tiles the POOLED grid
  int lw = 0, lh = 0;   // power-of-two brick of pooled pixels, at most 16 wide; it may overhang the pooled grid
  while (lw < 4 && (2 << lw) <= Wq) ++lw;
  while (lw + lh < 7 && (2 << lh) <= Hq) ++lh;
  const int wt = 1 << lw, ht = 1 << lh;
  const int nt = 128 >> (lw + lh);
  g.H = H; g.W = W; g.lw = lw; g.lh = lh; g.tiles_w = (Wq + wt - 1) / wt; g.tiles_h = (Hq + ht - 1) / ht;
  g.total_tiles = (int)((n + nt - 1) / nt) * g.tiles_w * g.tiles_h;
  g.n_images = (int)n; g.cout_pad = cout_pad;
  const int grid = g.total_tiles < num_sms * 4 ? g.total_tiles : num_sms * 4;
  const uint16_t* w = reinterpret_cast<const uint16_t*>(w32x32);
  uint16_t* o = reinterpret_cast<uint16_t*>(out);
#define A3D_FTC(FMT_, ACT_) A3D_CUDA_OK(launch_chain(conv2d_first_tc_kernel<FMT_, ACT_>, dim3(grid), dim3(128), 0, st, 1, in, w, scale, shift, o, g))
#define A3D_FTC_ACT(FMT_)                                            \
  switch (act) {                                                     \
    case A3D_ACT_ELU: A3D_FTC(FMT_, A3D_ACT_ELU); break;             \
    case A3D_ACT_RELU: A3D_FTC(FMT_, A3D_ACT_RELU); break;           \
    case A3D_ACT_LRELU: A3D_FTC(FMT_, A3D_ACT_LRELU); break;         \
    case A3D_ACT_LRELU01: A3D_FTC(FMT_, A3D_ACT_LRELU01); break;     \
    case A3D_ACT_NONE: A3D_FTC(FMT_, A3D_ACT_NONE); break;           \
    default: set_error("conv2d_first: unsupported activation %d", act); return A3D_ERR_INVALID; \
  }
  if (fmt == A3D_DTYPE_F16) { A3D_FTC_ACT(A3D_DTYPE_F16) } else { A3D_FTC_ACT(A3D_DTYPE_BF16) }
#undef A3D_FTC_ACT
#undef A3D_FTC
  A3D_CUDA_OK(cudaGetLastError());
  if (launches) ++*launches;
  return A3D_OK;
}

}  // namespace a3d
