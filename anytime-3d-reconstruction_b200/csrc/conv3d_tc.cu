// Conv3D(k = 4, strides 2 or 1, 'same', no bias) + folded BatchNorm + activation of the voxel encoder as tcgen05 implicit
// GEMMs (conv3DEnc / encoder3D, src/net_core/autoencoder3D.py:26-39,72-102).
//
//   D[(n, o), co] = sum_{kd,kh,kw,ci} X[n, s*o + k - 1, ci] * W[kd, kh, kw, ci, co]      ('same': pad_before = 1 for k = 4)
//
// conv3d_tc_kernel (Cin % 64 == 0): the M tile is a brick of bw x bh x bd OUTPUT voxels x nt objects (128 rows).  For a
// tap (kd, kh, kw) the rows of the A operand are the input voxels s*o + k - 1: ONE 5-D TMA box of the (c, w, h, d, n)
// view with element strides (1, s, s, s, 1) started at (s*o0 + k - 1): the stride-2 gather and the zero padding (TMA
// out-of-bounds fill) both happen inside the copy engine.  K = 64 taps x Cin in chunks of 64; weights repacked to
// [tap][co][ci]; fp32 accumulators double-buffered in TMEM; same warp roles as conv2d_tc.cu.
//
// conv3d_first_tc_kernel (Cin = 1, the occupancy grid): K = 64 taps exactly fills one 128-byte K-major row, so thread r
// gathers the 4 x 4 x 4 neighbourhood of output voxel r from the fp32 grid, rounds it to the operand type and stores
// row r of the A tile itself (SWIZZLE_128B: 16-byte chunk c of row r at chunk c ^ (r & 7)); four tcgen05.mma
// (M = 128, N = 64, K = 16) against the resident 64 x 64 weight tile.
#include <cstdlib>

#include "epilogue.cuh"
#include "internal.h"
#include "ptx.cuh"

namespace a3d {
namespace {

constexpr int BM = 128;
constexpr int kEpiWarps = 16;
constexpr int kThreads = 128 + 32 * kEpiWarps;

// MT = M tiles per unit.  MT = 2 (128-wide N tiles only): two 128-voxel bricks share every weight tile of a K step (two
// MMAs per K substep against the same B descriptor, two accumulators of BN columns) -- the 64 -> 128 layer streams its
// whole 1 MB weight set per brick otherwise and is bound by L2 -> SM operand traffic, not by the tensor pipe.
template <int BN, int MT>
struct Cfg3 {
  static constexpr int A_BYTES = MT * BM * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGES = BN == 256 ? 4 : (MT == 2 ? 4 : 6);
  static constexpr int NUM_BARS = 2 * STAGES + 4;
  static constexpr int STAGING_BYTES = kEpiWarps * 32 * 64;   // 2 KB per epilogue warp: 32 rows x 32 columns x 16 bit
  static constexpr int SMEM_BYTES = 1024 + STAGES * (A_BYTES + B_BYTES) + STAGING_BYTES + NUM_BARS * 8 + 16;
};

// MODE 0: 16-bit [voxels, cout_pad]; 1: fp32 [voxels, cout_real] (final conv, feeds the global pool)
template <int BN, int FMT, int ACT, int MODE, int MT>
__global__ void __launch_bounds__(kThreads, 1)
conv3d_tc_kernel(const __grid_constant__ CUtensorMap tmap_act, const __grid_constant__ CUtensorMap tmap_wgt,
                 void* __restrict__ out, const float* __restrict__ scale, const float* __restrict__ shift,
                 Conv3dGeom g) {
  using C = Cfg3<BN, MT>;
  static_assert(MT == 1 || (MT == 2 && BN == 128), "two M tiles per unit are built for 128-wide N tiles");
  constexpr int CW = MT * BN / 4;   // columns per epilogue warp: 4 warps per TMEM lane quarter, split over (M tile, columns)
  constexpr int A_BYTES = C::A_BYTES;
  constexpr int TBUF = MT * BN;     // TMEM columns per accumulator buffer
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::STAGES * A_BYTES;
  uint8_t* smem_stage = smem_b + C::STAGES * C::B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stage + C::STAGING_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = full + C::STAGES;
  uint64_t* t_full = empty + C::STAGES;
  uint64_t* t_empty = t_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_units = (g.m_tiles / MT) * g.n_tiles;   // the host launches MT = 2 only for an even tile count
  const int ksteps = 64 * g.cin_chunks;
  const int tiles_per_obj = g.tiles_w * g.tiles_h * g.tiles_d;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_act);
    ptx::prefetch_tmap(&tmap_wgt);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::STAGES; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&t_full[i], 1); ptx::mbar_init(&t_empty[i], kEpiWarps); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc<1>(tmem_slot, 512);
    ptx::tmem_relinquish<1>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_sync();   // programmatic dependent launch: the prologue above overlaps the previous layer (ptx.cuh)

  if (warp == 0 || warp == 3) {
    // ===================================================== TMA producers: warp 0 feeds the even K steps, warp 3 the odd ones
    // (see conv2d_tc.cu: one warp's ~550 clk of dependent issue work per K step starves the MMA warp)
    const int par = warp == 3 ? 1 : 0;
    int s = par;
    uint32_t ph = 0, cnt = 0;
    static_assert(C::STAGES % 2 == 0, "two producer warps need an even stage count");
    for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
      const int nt = u % g.n_tiles, mu = u / g.n_tiles;
      int w0[MT], h0[MT], d0[MT], n0[MT];
#pragma unroll
      for (int t = 0; t < MT; ++t) {
        const int mt = mu * MT + t;
        const int tb = mt % tiles_per_obj, nb = mt / tiles_per_obj;
        const int tw = tb % g.tiles_w, th = (tb / g.tiles_w) % g.tiles_h, td = tb / (g.tiles_w * g.tiles_h);
        w0[t] = ((tw << g.lw) * g.stride) - 1; h0[t] = ((th << g.lh) * g.stride) - 1; d0[t] = ((td << g.ld) * g.stride) - 1;
        n0[t] = nb << (7 - g.lw - g.lh - g.ld);
      }
      int brow = nt * BN;
      for (int kd = 0; kd < 4; ++kd)
        for (int kh = 0; kh < 4; ++kh)
          for (int kw = 0; kw < 4; ++kw) {
            for (int kc = 0; kc < g.cin_chunks; ++kc, ++cnt) {
              if ((cnt & 1u) != (uint32_t)par) continue;
              ptx::mbar_wait(&empty[s], ph ^ 1);
              if (ptx::elect_one()) {
                ptx::mbar_expect_tx(&full[s], A_BYTES + C::B_BYTES);
#pragma unroll
                for (int t = 0; t < MT; ++t)
                  ptx::tma_load_5d(smem_a + s * A_BYTES + t * (BM * 128), &tmap_act, &full[s], kc * 64, w0[t] + kw, h0[t] + kh,
                                   d0[t] + kd, n0[t]);
                ptx::tma_load_2d(smem_b + s * C::B_BYTES, &tmap_wgt, &full[s], kc * 64, brow);
              }
              __syncwarp();
              s += 2;
              if (s >= C::STAGES) { s -= C::STAGES; ph ^= 1; }
            }
            brow += g.cout_pad;
          }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (converged warp, elected-lane issue)
    constexpr uint32_t idesc = ptx::make_idesc_f16(BM, BN, FMT);
    const uint32_t a_lo0 = ptx::sw128_desc_lo(ptx::smem_u32(smem_a));
    const uint32_t b_lo0 = ptx::sw128_desc_lo(ptx::smem_u32(smem_b));
    uint32_t unit_it = 0, ph = 0;
    int s = 0;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++unit_it) {
      const int buf = unit_it & 1;
      ptx::mbar_wait(&t_empty[buf], ((unit_it >> 1) & 1) ^ 1);
      ptx::tc_fence_after();
      const uint32_t tacc = tmem_base + buf * TBUF;
      for (int ks = 0; ks < ksteps; ++ks) {
        ptx::mbar_wait(&full[s], ph);
        ptx::tc_fence_after();
        const uint32_t a_lo = a_lo0 + s * (A_BYTES >> 4), b_lo = b_lo0 + s * (C::B_BYTES >> 4);
        if (ptx::elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int t = 0; t < MT; ++t)
              ptx::umma_f16<1>(tacc + t * BN, ptx::sw128_desc(a_lo + t * ((BM * 128) >> 4) + kk * 2), ptx::sw128_desc(b_lo + kk * 2),
                               idesc, (ks | kk) != 0);
          ptx::umma_commit<1>(&empty[s]);
          if (ks == ksteps - 1) ptx::umma_commit<1>(&t_full[buf]);
        }
        __syncwarp();
        if (++s == C::STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue: TMEM -> BN -> act -> global
    const int e = warp - 4;
    const int quarter = e & 3;
    const int half = MT == 2 ? (e >> 2) & 1 : 0;      // which M tile of the unit
    const int cgrp = MT == 2 ? e >> 3 : e >> 2;       // column group of CW columns
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    constexpr int GROUPS = CW / 32;
    uint32_t unit_it = 0;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++unit_it) {
      const int nt = u % g.n_tiles, mt = (u / g.n_tiles) * MT + half;
      const int buf = unit_it & 1;
      ptx::mbar_wait(&t_full[buf], (unit_it >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t tacc = tmem_base + lane_base + buf * TBUF + half * BN + cgrp * CW;
      const int tb = mt % tiles_per_obj, nb = mt / tiles_per_obj;
      const int tw = tb % g.tiles_w, th = (tb / g.tiles_w) % g.tiles_h, td = tb / (g.tiles_w * g.tiles_h);
      const int r = quarter * 32 + lane;
      const int wi = r & ((1 << g.lw) - 1), hi = (r >> g.lw) & ((1 << g.lh) - 1);
      const int di = (r >> (g.lw + g.lh)) & ((1 << g.ld) - 1), ni = r >> (g.lw + g.lh + g.ld);
      const int obj = (nb << (7 - g.lw - g.lh - g.ld)) + ni;
      const int ow = (tw << g.lw) + wi, oh = (th << g.lh) + hi, od = (td << g.ld) + di;
      const bool row_ok = obj < g.n_objects;
      const int64_t p = (((int64_t)obj * g.G + od) * g.G + oh) * g.G + ow;
      int64_t prow[4];   // MODE 0: voxel index of the 4 rows this lane writes out (8 * j + lane / 4), -1 if past the batch
      if constexpr (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int rr = quarter * 32 + 8 * j + (lane >> 2);
          const int rwi = rr & ((1 << g.lw) - 1), rhi = (rr >> g.lw) & ((1 << g.lh) - 1);
          const int rdi = (rr >> (g.lw + g.lh)) & ((1 << g.ld) - 1), rni = rr >> (g.lw + g.lh + g.ld);
          const int robj = (nb << (7 - g.lw - g.lh - g.ld)) + rni;
          prow[j] = robj < g.n_objects
                        ? (((int64_t)robj * g.G + (td << g.ld) + rdi) * g.G + (th << g.lh) + rhi) * g.G + (tw << g.lw) + rwi
                        : -1;
        }
      }
#pragma unroll 1
      for (int gi = 0; gi < GROUPS; ++gi) {
        const int co0 = nt * BN + cgrp * CW + gi * 32;
        uint32_t v[32];
        ptx::tmem_ld16(tacc + gi * 32, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
        ptx::tmem_ld16(tacc + gi * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
        ptx::tmem_ld_wait();
        const float4* sc4 = reinterpret_cast<const float4*>(scale + co0);
        const float4* sh4 = reinterpret_cast<const float4*>(shift + co0);
        if constexpr (MODE == 1) {
          float* dst = reinterpret_cast<float*>(out) + p * g.cout_real + co0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 sc = __ldg(sc4 + i), sh = __ldg(sh4 + i);
            float x[4];
            x[0] = activate<ACT>(fmaf(__uint_as_float(v[4 * i]), sc.x, sh.x));
            x[1] = activate<ACT>(fmaf(__uint_as_float(v[4 * i + 1]), sc.y, sh.y));
            x[2] = activate<ACT>(fmaf(__uint_as_float(v[4 * i + 2]), sc.z, sh.z));
            x[3] = activate<ACT>(fmaf(__uint_as_float(v[4 * i + 3]), sc.w, sh.w));
            if (row_ok) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (co0 + 4 * i + j < g.cout_real) dst[4 * i + j] = x[j];
            }
          }
        } else {
          uint32_t o[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 sc = __ldg(sc4 + i), sh = __ldg(sh4 + i);
            const float x0 = activate<ACT>(fmaf(__uint_as_float(v[4 * i]), sc.x, sh.x));
            const float x1 = activate<ACT>(fmaf(__uint_as_float(v[4 * i + 1]), sc.y, sh.y));
            const float x2 = activate<ACT>(fmaf(__uint_as_float(v[4 * i + 2]), sc.z, sh.z));
            const float x3 = activate<ACT>(fmaf(__uint_as_float(v[4 * i + 3]), sc.w, sh.w));
            o[2 * i] = pack2<FMT>(x0, x1);
            o[2 * i + 1] = pack2<FMT>(x2, x3);
          }
          // coalesced write-out through a per-warp swizzled staging tile (32 rows x 64 B): every store instruction
          // then covers 8 rows x 64 contiguous bytes (full sectors) instead of 16 bytes per lane in 32 different lines
          uint8_t* stg = smem_stage + e * 2048;
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4)
            *reinterpret_cast<uint4*>(stg + lane * 64 + ((c4 ^ ((lane >> 1) & 3)) << 4)) =
                make_uint4(o[4 * c4], o[4 * c4 + 1], o[4 * c4 + 2], o[4 * c4 + 3]);
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int rr = 8 * j + (lane >> 2), ch = lane & 3;
            const uint4 q = *reinterpret_cast<const uint4*>(stg + rr * 64 + ((ch ^ ((rr >> 1) & 3)) << 4));
            if (prow[j] >= 0)
              *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(out) + prow[j] * g.cout_pad + co0 + ch * 8) = q;
          }
          __syncwarp();
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&t_empty[buf]);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<1>(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// First layer: fp32 occupancy grid [n, Gin^3] -> 16-bit [n, (Gin/2)^3, 64]; thread = output voxel of a 16 x 8 x 1 brick.
template <int FMT, int ACT>
__global__ void __launch_bounds__(128, 4)
conv3d_first_tc_kernel(const float* __restrict__ in, const uint16_t* __restrict__ w64x64, const float* __restrict__ scale,
                       const float* __restrict__ shift, uint16_t* __restrict__ out, int n_objects, int Gin, int total_tiles) {
  __shared__ __align__(1024) uint8_t sA[128 * 128];
  __shared__ __align__(1024) uint8_t sB[64 * 128];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float ss[64], sh[64];
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 64 * 8; i += 128) {   // 64 rows x 8 chunks of the weight tile [co][k]
    const int row = i >> 3, c = i & 7;
    const uint4 v = *reinterpret_cast<const uint4*>(w64x64 + row * 64 + c * 8);
    *reinterpret_cast<uint4*>(sB + row * 128 + ((c ^ (row & 7)) << 4)) = v;
  }
  if (tid < 64) { ss[tid] = scale[tid]; sh[tid] = shift[tid]; }
  if (tid == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  if (warp == 0) { ptx::tmem_alloc<1>(&tmem_slot, 64); ptx::tmem_relinquish<1>(); }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  ptx::pdl_sync();   // programmatic dependent launch: the prologue above overlaps the previous kernel (ptx.cuh)
  constexpr uint32_t idesc = ptx::make_idesc_f16(128, 64, FMT);
  const uint32_t a_lo = ptx::sw128_desc_lo(ptx::smem_u32(sA)), b_lo = ptx::sw128_desc_lo(ptx::smem_u32(sB));
  const int G = Gin >> 1;                       // output grid
  const int tiles_w = G >> 4, tiles_h = G >> 3; // 16 x 8 x 1 bricks
  const int wi = tid & 15, hi = tid >> 4;
  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int tw = tile % tiles_w, th = (tile / tiles_w) % tiles_h;
    const int od = (tile / (tiles_w * tiles_h)) % G, obj = tile / (tiles_w * tiles_h * G);
    const int ow = (tw << 4) + wi, oh = (th << 3) + hi;
    const float* base = in + (int64_t)obj * Gin * Gin * Gin;
    const int iw0 = 2 * ow - 1;
    // every thread gathers the 4 x 4 x 4 neighbourhood of its output voxel (48 independent loads in flight; L1 serves the
    // 8-fold reuse).  Staging the 4 x 18 x 34 patch of the tile in shared memory first was measured twice in round 2: 2.3 x
    // slower with a rolled copy loop (20 serial L2 latencies per tile), no faster than this gather with the loads unrolled.
#pragma unroll
    for (int kd = 0; kd < 4; ++kd) {
      const int id = 2 * od + kd - 1;
      uint32_t u[8];
#pragma unroll
      for (int kh = 0; kh < 4; ++kh) {
        const int ih = 2 * oh + kh - 1;
        const bool ok = id >= 0 && id < Gin && ih >= 0 && ih < Gin;
        const float* row = base + ((int64_t)id * Gin + ih) * Gin;
        // w taps iw0 .. iw0+3: iw0 is odd, so the middle pair is an aligned float2
        const float a = (ok && iw0 >= 0) ? __ldg(row + iw0) : 0.f;
        float2 m = make_float2(0.f, 0.f);
        if (ok) m = __ldg(reinterpret_cast<const float2*>(row + iw0 + 1));
        const float d = (ok && iw0 + 3 < Gin) ? __ldg(row + iw0 + 3) : 0.f;
        u[2 * kh] = pack2<FMT>(a, m.x);
        u[2 * kh + 1] = pack2<FMT>(m.y, d);
      }
      // taps (kd, 0..3, 0..3) = K columns kd*16 .. kd*16+15 = 16-byte chunks 2*kd, 2*kd+1 of row tid
      *reinterpret_cast<uint4*>(sA + tid * 128 + (((2 * kd) ^ (tid & 7)) << 4)) = make_uint4(u[0], u[1], u[2], u[3]);
      *reinterpret_cast<uint4*>(sA + tid * 128 + (((2 * kd + 1) ^ (tid & 7)) << 4)) = make_uint4(u[4], u[5], u[6], u[7]);
    }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          ptx::umma_f16<1>(tmem, ptx::sw128_desc(a_lo + kk * 2), ptx::sw128_desc(b_lo + kk * 2), idesc, kk != 0);
        ptx::umma_commit<1>(&bar);
      }
      __syncwarp();
    }
    ptx::mbar_wait(&bar, phase);
    phase ^= 1u;
    ptx::tc_fence_after();
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    // the MMAs have completed: the A tile is free and becomes the output staging buffer (row tid = voxel tid, 128 bytes,
    // same XOR swizzle), so that the global stores below are full 512-byte runs instead of 16 bytes per lane per row
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t v[32];
      ptx::tmem_ld16(taddr + half * 32, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
      ptx::tmem_ld16(taddr + half * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
      ptx::tmem_ld_wait();
      uint32_t o[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int c = half * 32 + 2 * i;
        o[i] = pack2<FMT>(activate<ACT>(fmaf(__uint_as_float(v[2 * i]), ss[c], sh[c])),
                          activate<ACT>(fmaf(__uint_as_float(v[2 * i + 1]), ss[c + 1], sh[c + 1])));
      }
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4)
        *reinterpret_cast<uint4*>(sA + tid * 128 + (((half * 4 + c4) ^ (tid & 7)) << 4)) =
            make_uint4(o[4 * c4], o[4 * c4 + 1], o[4 * c4 + 2], o[4 * c4 + 3]);
    }
    ptx::tc_fence_before();
    __syncwarp();
    {
      const int lane = tid & 31;
      const int64_t tile_p = (((int64_t)obj * G + od) * G + (th << 3)) * G + (tw << 4);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int rr = warp * 32 + 4 * j + (lane >> 3), ch = lane & 7;   // 4 rows x 128 B = 512 contiguous bytes
        const uint4 q = *reinterpret_cast<const uint4*>(sA + rr * 128 + ((ch ^ (rr & 7)) << 4));
        *reinterpret_cast<uint4*>(out + (tile_p + (int64_t)(rr >> 4) * G + (rr & 15)) * 64 + ch * 8) = q;
      }
    }
    __syncwarp();   // the next tile's A rows of this warp overwrite the staging rows it just read
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<1>(tmem, 64);
}

template <int BN, int FMT, int MODE, int MT>
int launch_act(const CUtensorMap& ta, const CUtensorMap& tw, void* out, const float* scale, const float* shift,
               const Conv3dGeom& g, int act, int grid, cudaStream_t st) {
  auto launch = [&](auto kern) -> int {
    A3D_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg3<BN, MT>::SMEM_BYTES));
    A3D_CUDA_OK(launch_chain(kern, dim3(grid), dim3(kThreads), Cfg3<BN, MT>::SMEM_BYTES, st, 1, ta, tw, out, scale, shift, g));
    return A3D_OK;
  };
  switch (act) {
    case A3D_ACT_ELU: return launch(conv3d_tc_kernel<BN, FMT, A3D_ACT_ELU, MODE, MT>);
    case A3D_ACT_RELU: return launch(conv3d_tc_kernel<BN, FMT, A3D_ACT_RELU, MODE, MT>);
    case A3D_ACT_LRELU: return launch(conv3d_tc_kernel<BN, FMT, A3D_ACT_LRELU, MODE, MT>);
    case A3D_ACT_NONE: return launch(conv3d_tc_kernel<BN, FMT, A3D_ACT_NONE, MODE, MT>);
    default: set_error("conv3d: unsupported activation %d", act); return A3D_ERR_INVALID;
  }
}

}  // namespace

int launch_conv3d_tc(const CUtensorMap& tmap_act, const CUtensorMap& tmap_wgt, void* out, const float* scale,
                     const float* shift, const Conv3dGeom& g, int bn, int fmt, int act, bool out_f32, int num_sms,
                     cudaStream_t st, int64_t* launches) {
  if (g.n_objects <= 0) return A3D_OK;
  // two bricks per unit share the weight tiles when the N tile is 128 wide and the brick count is even
  static const bool no_mt2 = [] { const char* e = getenv("A3D_ENC3D_MT2"); return e && e[0] == '0'; }();
  const bool mt2 = bn == 128 && (g.m_tiles % 2) == 0 && !no_mt2 && (g.m_tiles / 2) * g.n_tiles >= num_sms;   // keep every SM busy
  const int total = (g.m_tiles / (mt2 ? 2 : 1)) * g.n_tiles;
  const int grid = total < num_sms ? total : num_sms;
  int rc;
#define A3D_C3D(BN_, MT_)                                                                                               \
  (fmt == A3D_DTYPE_F16                                                                                                  \
       ? (out_f32 ? launch_act<BN_, A3D_DTYPE_F16, 1, MT_>(tmap_act, tmap_wgt, out, scale, shift, g, act, grid, st)       \
                  : launch_act<BN_, A3D_DTYPE_F16, 0, MT_>(tmap_act, tmap_wgt, out, scale, shift, g, act, grid, st))      \
       : (out_f32 ? launch_act<BN_, A3D_DTYPE_BF16, 1, MT_>(tmap_act, tmap_wgt, out, scale, shift, g, act, grid, st)      \
                  : launch_act<BN_, A3D_DTYPE_BF16, 0, MT_>(tmap_act, tmap_wgt, out, scale, shift, g, act, grid, st)))
  if (bn == 256) rc = A3D_C3D(256, 1);
  else if (bn == 128 && mt2) rc = A3D_C3D(128, 2);
  else if (bn == 128) rc = A3D_C3D(128, 1);
  else { set_error("conv3d: N tile must be 128 or 256"); return A3D_ERR_INVALID; }
#undef A3D_C3D
  if (rc == A3D_OK && launches) ++*launches;
  return rc;
}

int launch_conv3d_first_tc(const float* in, const void* w64x64, const float* scale, const float* shift, void* out,
                           int64_t n, int Gin, int fmt, int act, int num_sms, cudaStream_t st, int64_t* launches) {
  if (n <= 0) return A3D_OK;
  const int G = Gin / 2;
  if (G % 16 != 0) { set_error("conv3d_first: the output grid must be a multiple of 16"); return A3D_ERR_INVALID; }
  const int64_t tiles = n * G * (G / 8) * (G / 16);
  const int grid = (int)(tiles < (int64_t)num_sms * 4 ? tiles : (int64_t)num_sms * 4);
  const uint16_t* w = reinterpret_cast<const uint16_t*>(w64x64);
  uint16_t* o = reinterpret_cast<uint16_t*>(out);
#define A3D_F3(FMT_, ACT_) A3D_CUDA_OK(launch_chain(conv3d_first_tc_kernel<FMT_, ACT_>, dim3(grid), dim3(128), 0, st, 1, in, w, scale, shift, o, (int)n, Gin, (int)tiles))
#define A3D_F3_ACT(FMT_)                                       \
  switch (act) {                                               \
    case A3D_ACT_ELU: A3D_F3(FMT_, A3D_ACT_ELU); break;        \
    case A3D_ACT_RELU: A3D_F3(FMT_, A3D_ACT_RELU); break;      \
    case A3D_ACT_LRELU: A3D_F3(FMT_, A3D_ACT_LRELU); break;    \
    case A3D_ACT_NONE: A3D_F3(FMT_, A3D_ACT_NONE); break;      \
    default: set_error("conv3d_first: unsupported activation %d", act); return A3D_ERR_INVALID; \
  }
  if (fmt == A3D_DTYPE_F16) { A3D_F3_ACT(A3D_DTYPE_F16) } else { A3D_F3_ACT(A3D_DTYPE_BF16) }
#undef A3D_F3_ACT
#undef A3D_F3
  A3D_CUDA_OK(cudaGetLastError());
  if (launches) ++*launches;
  return A3D_OK;
}

}  // namespace a3d
