// Stride-1 'same' Conv2D (k = 1 or 3, no bias) + folded BatchNorm + activation as a tcgen05 implicit GEMM
// (Darknet19Conv / convHead / the final head conv, src/net_core/darknet.py:83-94,135-147,155-157).
//
//   D[p, co] = sum_{dy,dx,ci} X[n, h+dy-1, w+dx-1, ci] * W[dy, dx, ci, co]        p = (n, h, w) flattened, NHWC
//
// M tile = a brick of wt x ht pixels x nt images (wt * ht * nt = 128, powers of two; the encoder's grids are powers of
// two), N tile = BN output channels, K = taps x Cin in chunks of 64 channels.  A operand: one 4-D TMA box
// (64 ch, wt, ht, nt) of the (c, w, h, n) view per (tap, chunk), started at (w0+dx-1, h0+dy-1): the out-of-bounds zero
// fill of TMA *is* the 'same' padding, per image.  B operand: 2-D TMA box (64 ci, BN rows) of the weights repacked to
// [tap][co][ci].  fp32 accumulators double-buffered in TMEM, so the BN + activation epilogue of unit i overlaps the
// main loop of unit i+1; persistent CTAs, units ordered n-tile fastest so the CTAs that share an activation tile run
// at the same time (L2 reuse).
//
// POOL = true fuses the MaxPool2D(2, 2) that follows the convolution (darknet.py:100,104,110,116,124): the brick is
// then at most 16 pixels wide, so the 2 x 2 partners of a pixel sit in lanes (lane ^ 1) and (lane ^ wt) of the same
// epilogue warp: two shuffles + max per accumulator, taken BEFORE the BN shift and the activation (both monotone), so
// the activation work drops 4x; each lane of the quad finishes and stores one 16-byte quarter of the pooled 64-byte
// segment.
#include <cstdlib>

#include "epilogue.cuh"
#include "internal.h"
#include "ptx.cuh"

namespace a3d {
namespace {

constexpr int BM = 128;
constexpr int kEpiWarps = 16;   // 4 per TMEM lane quarter: the shallow layers (K = 9 steps) are epilogue-latency bound
constexpr int kThreads = 128 + 32 * kEpiWarps;

// KC = channels per K step: 64 (128-byte rows, SWIZZLE_128B) or 32 (64-byte rows, SWIZZLE_64B; the 32-channel input of
// the second Darknet19 conv, which would waste half of every MMA if padded to 64)
template <int BN, int KC>
struct Cfg {
  static constexpr int A_BYTES = BM * KC * 2;
  static constexpr int B_BYTES = BN * KC * 2;
  static constexpr int STAGES = KC == 32 ? 12 : (BN == 256 ? 4 : (BN == 128 ? 4 : 6));
  static constexpr int NUM_BARS = 2 * STAGES + 5;
  static constexpr int STAGING_BYTES = kEpiWarps * 32 * 64;   // 2 KB per epilogue warp: 32 rows x 32 columns x 16 bit
  static constexpr int SS_BYTES = BN <= 128 ? kEpiWarps * 256 : 0;   // per-warp copy of its 32 scale + 32 shift values
  static constexpr int SMEM_BYTES = 1024 + STAGES * (A_BYTES + B_BYTES) + STAGING_BYTES + SS_BYTES + NUM_BARS * 8 + 16;
};

// RH variant ("resident weights + haloed activation box") for the shallow 3 x 3 layers whose whole weight set fits in
// shared memory (9 x BN x KC x 2 bytes, one Cin chunk, one N tile): these layers are bound by L2 -> SM operand traffic
// (~42 B/clk/SM), not by the tensor pipe.  (1) the nine weight tiles are loaded once per CTA and stay resident;
// (2) the brick is 16 x 8 pixels and ONE TMA box of 16 x 10 pixels per dx serves the three dy taps: rows are ordered
// (h, w) so a shift of one image row is 16 GEMM rows = whole swizzle atoms = just another descriptor start address.
// Operand bytes per 128-pixel tile: 288 KB -> 60 KB (64 -> 128 channels), 108 KB -> 30 KB (32 -> 64 channels).
template <int BN, int KC, int MODE>
struct CfgRH {
  // MODE 3 (pool through four accumulators): GEMM rows are POOLED pixels; a stage is one element-strided box of
  // 16 x 9 pooled positions (one of 4 x-offsets x 2 row parities), two row-shifted views each
  static constexpr int HALO_ROWS = MODE == 3 ? 16 * 9 : 16 * 10;
  static constexpr int A_BYTES = HALO_ROWS * KC * 2;
  static constexpr int W_TILE = BN * KC * 2;
  static constexpr int RES_BYTES = 9 * W_TILE;
  static constexpr int STAGING_BYTES = (MODE == 0 || MODE == 3) ? kEpiWarps * 32 * 64 : 0;
  // 128-wide N tile: 144 KB of weights are resident, so the plain 16-bit mode (which needs the 32 KB store staging) keeps
  // two 20 KB activation stages, the pooled mode four
  static constexpr int STAGES = KC == 32 ? 8 : (MODE == 0 ? 2 : 4);
  static constexpr int SS_BYTES = 4 * 256;   // one N tile: the four lane-quarter warps of a column group share one copy
  static constexpr int NUM_BARS = 2 * STAGES + 5;
  static constexpr int SMEM_BYTES = 1024 + STAGES * A_BYTES + RES_BYTES + STAGING_BYTES + SS_BYTES + NUM_BARS * 8 + 16;
};

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1),
      "r"(c2), "r"(c3)
      : "memory");
}

template <int ACT>
__device__ __forceinline__ float act2d(float v) {
  if constexpr (ACT == A3D_ACT_LRELU01) return v > 0.f ? v : 0.1f * v;
  else return activate<ACT>(v);
}

// MODE 0: 16-bit [pixels, cout_pad]; 1: fp32 [pixels, cout_real] (final head conv; feeds the global pool);
// 2: 16-bit with the 2 x 2 max-pool fused, [n, H/2, W/2, cout_pad] (pool by warp shuffles);
// 3 (RH, BN = 64 only): same output as 2, but GEMM rows are pooled pixels and each of the four pool-window positions has its
//    own accumulator (4 x 64 TMEM columns per buffer): the pool is a per-thread max, no shuffles
template <int BN, int KC, int FMT, int ACT, int MODE, bool RH>
__global__ void __launch_bounds__(kThreads, 1)
conv2d_tc_kernel(const __grid_constant__ CUtensorMap tmap_act, const __grid_constant__ CUtensorMap tmap_wgt,
                 void* __restrict__ out, const float* __restrict__ scale, const float* __restrict__ shift,
                 Conv2dGeom g) {
  using C = Cfg<BN, KC>;
  constexpr int WPQ = BN >= 128 ? 4 : 2;   // epilogue warps per lane quarter; each owns CW = BN / WPQ columns
  constexpr int CW = BN / WPQ;
  using R = CfgRH<BN, KC, MODE>;
  constexpr int A_BYTES = RH ? R::A_BYTES : C::A_BYTES;
  constexpr int STAGES = RH ? R::STAGES : C::STAGES;
  constexpr int B_REGION = RH ? R::RES_BYTES : C::STAGES * C::B_BYTES;     // resident weight tiles / per-stage weight tiles
  constexpr int STAGING = RH ? R::STAGING_BYTES : C::STAGING_BYTES;
  constexpr bool SS_SMEM = BN <= 128 && MODE != 2;
  constexpr int TBUF = MODE == 3 ? 4 * BN : BN;   // TMEM columns per accumulator buffer   // scale / shift of a warp's 32 columns live in shared memory (LDS broadcast) instead
                                        // of 16 LDG.128 per tile behind the store queue (stall_lg / long_scoreboard); not in
                                        // the pooled mode, whose shuffles already load the MIO pipe (measured: slower)
  constexpr int SS_BYTES = SS_SMEM ? (RH ? R::SS_BYTES : C::SS_BYTES) : 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_BYTES;
  uint8_t* smem_stage = smem_b + B_REGION;
  float* smem_ss = reinterpret_cast<float*>(smem_stage + STAGING);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stage + STAGING + SS_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = full + STAGES;
  uint64_t* t_full = empty + STAGES;
  uint64_t* t_empty = t_full + 2;
  uint64_t* w_full = t_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_units = g.m_tiles * g.n_tiles;
  const int ksteps = g.taps * g.cin_chunks;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_act);
    ptx::prefetch_tmap(&tmap_wgt);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&t_full[i], 1); ptx::mbar_init(&t_empty[i], 4 * WPQ); }
    ptx::mbar_init(w_full, 1);
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
    // ===================================================== TMA producers (two converged warps, elected-lane issue): warp 0
    // feeds the even K steps, warp 3 the odd ones (STAGES is even, so each warp owns the stages of its parity).  One
    // warp needs ~550 clk per K step for its ~90 dependent uniform-datapath instructions (ncu source page), which starved
    // the MMA warp on every layer whose K step holds less than that much tensor work.  Stage / phase / tap offsets
    // advance incrementally (the first version's div / mod per K step cost another ~150 clk).
    const int par = warp == 3 ? 1 : 0;
    int s = par;
    uint32_t ph = 0, cnt = 0;
    static_assert(STAGES % 2 == 0, "two producer warps need an even stage count");
    if constexpr (RH) {
      if (par == 0) {   // resident weights: nine [BN x KC] tiles, once per CTA
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(w_full, R::RES_BYTES);
          for (int tap = 0; tap < 9; ++tap)
            ptx::tma_load_2d(smem_b + tap * R::W_TILE, &tmap_wgt, w_full, 0, tap * g.cout_pad);
        }
        __syncwarp();
      }
      if constexpr (MODE == 3) {
        for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
          const int tw = u % g.tiles_w, th = (u / g.tiles_w) % g.tiles_h, nb = u / (g.tiles_w * g.tiles_h);
          const int w0 = (tw << 5) - 1, h0 = (th << 4) - 1;       // conv-pixel origin of the 16 x 8 pooled brick, minus the pad
          for (int box = 0; box < 8; ++box, ++cnt) {               // box = x-offset (0..3) * 2 + row parity
            if ((cnt & 1u) != (uint32_t)par) continue;
            ptx::mbar_wait(&empty[s], ph ^ 1);
            if (ptx::elect_one()) {
              ptx::mbar_expect_tx(&full[s], A_BYTES);
              // element strides (1, 2, 2, 1): pixels (w0 + xo + 2*i, h0 + yp + 2*j), i < 16, j < 9
              tma_load_4d(smem_a + s * A_BYTES, &tmap_act, &full[s], 0, w0 + (box >> 1), h0 + (box & 1), nb);
            }
            __syncwarp();
            s += 2;
            if (s >= STAGES) { s -= STAGES; ph ^= 1; }
          }
        }
      } else
      for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
        const int tw = u % g.tiles_w, th = (u / g.tiles_w) % g.tiles_h, nb = u / (g.tiles_w * g.tiles_h);
        const int w0 = (tw << 4) - 1, h0 = (th << 3) - 1;
        for (int dx = 0; dx < 3; ++dx, ++cnt) {
          if ((cnt & 1u) != (uint32_t)par) continue;
          ptx::mbar_wait(&empty[s], ph ^ 1);
          if (ptx::elect_one()) {
            ptx::mbar_expect_tx(&full[s], A_BYTES);
            tma_load_4d(smem_a + s * A_BYTES, &tmap_act, &full[s], 0, w0 + dx, h0, nb);   // box (KC, 16, 10, 1)
          }
          __syncwarp();
          s += 2;
          if (s >= STAGES) { s -= STAGES; ph ^= 1; }
        }
      }
    } else {
    for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
      const int nt = u % g.n_tiles, mt = u / g.n_tiles;
      const int tw = mt % g.tiles_w, th = (mt / g.tiles_w) % g.tiles_h, nb = mt / (g.tiles_w * g.tiles_h);
      const int w0 = tw << g.lw, h0 = th << g.lh, n0 = nb << (7 - g.lw - g.lh);
      int dy = g.taps == 9 ? -1 : 0, dx = dy;
      int brow = nt * BN;
      for (int tap = 0; tap < g.taps; ++tap) {
        for (int kc = 0; kc < g.cin_chunks; ++kc, ++cnt) {
          if ((cnt & 1u) != (uint32_t)par) continue;
          ptx::mbar_wait(&empty[s], ph ^ 1);
          if (ptx::elect_one()) {
            ptx::mbar_expect_tx(&full[s], A_BYTES + C::B_BYTES);
            tma_load_4d(smem_a + s * A_BYTES, &tmap_act, &full[s], kc * KC, w0 + dx, h0 + dy, n0);
            ptx::tma_load_2d(smem_b + s * C::B_BYTES, &tmap_wgt, &full[s], kc * KC, brow);
          }
          __syncwarp();
          s += 2;
          if (s >= STAGES) { s -= STAGES; ph ^= 1; }
        }
        brow += g.cout_pad;
        if (++dx > 1) { dx = -1; ++dy; }
      }
    }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (converged warp, elected-lane issue)
    constexpr uint32_t idesc = ptx::make_idesc_f16(BM, BN, FMT);
    const uint32_t a_lo0 = ptx::sw128_desc_lo(ptx::smem_u32(smem_a));
    const uint32_t b_lo0 = ptx::sw128_desc_lo(ptx::smem_u32(smem_b));
    uint32_t unit_it = 0, ph = 0;
    int s = 0;
    if constexpr (RH) { ptx::mbar_wait(w_full, 0); ptx::tc_fence_after(); }
    for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++unit_it) {
      const int buf = unit_it & 1;
      ptx::mbar_wait(&t_empty[buf], ((unit_it >> 1) & 1) ^ 1);
      ptx::tc_fence_after();
      const uint32_t tacc = tmem_base + buf * TBUF;
      if constexpr (RH && MODE == 3) {
        constexpr uint32_t DY_STEP = (16 * KC * 2) >> 4;   // one pooled row = 16 GEMM rows
        uint32_t touched = 0;
        for (int box = 0; box < 8; ++box) {
          const int xo = box >> 1, yp = box & 1;
          ptx::mbar_wait(&full[s], ph);
          ptx::tc_fence_after();
          const uint32_t a_lo = a_lo0 + s * (A_BYTES >> 4);
          if (ptx::elect_one()) {
            for (int sh = 0; sh < 2; ++sh) {
              const int yo = yp + 2 * sh;                     // row offset of this view inside the 4 x 4 patch
              for (int py = 0; py < 2; ++py) {
                const int dy = yo - py;
                if (dy < 0 || dy > 2) continue;
                for (int px = 0; px < 2; ++px) {
                  const int dx = xo - px;
                  if (dx < 0 || dx > 2) continue;
                  const int a = py * 2 + px;                  // accumulator of pool-window position (py, px)
                  const uint32_t b_lo = b_lo0 + (dy * 3 + dx) * (R::W_TILE >> 4);
#pragma unroll
                  for (int kk = 0; kk < KC / 16; ++kk) {
                    const uint32_t acc = ((touched >> a) & 1u) | (uint32_t)(kk != 0);
                    if constexpr (KC == 64)
                      ptx::umma_f16<1>(tacc + a * BN, ptx::sw128_desc(a_lo + sh * DY_STEP + kk * 2), ptx::sw128_desc(b_lo + kk * 2), idesc, acc);
                    else
                      ptx::umma_f16<1>(tacc + a * BN, ptx::sw64_desc(a_lo + sh * DY_STEP + kk * 2), ptx::sw64_desc(b_lo + kk * 2), idesc, acc);
                  }
                  touched |= 1u << a;
                }
              }
            }
            ptx::umma_commit<1>(&empty[s]);
            if (box == 7) ptx::umma_commit<1>(&t_full[buf]);
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      } else if constexpr (RH) {
        constexpr uint32_t DY_STEP = (16 * KC * 2) >> 4;   // one image row of the haloed box = 16 GEMM rows
        for (int dx = 0; dx < 3; ++dx) {
          ptx::mbar_wait(&full[s], ph);
          ptx::tc_fence_after();
          const uint32_t a_lo = a_lo0 + s * (A_BYTES >> 4);
          if (ptx::elect_one()) {
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              const uint32_t b_lo = b_lo0 + (dy * 3 + dx) * (R::W_TILE >> 4);
#pragma unroll
              for (int kk = 0; kk < KC / 16; ++kk) {
                const uint32_t acc = (dx | dy | kk) != 0;
                if constexpr (KC == 64)
                  ptx::umma_f16<1>(tacc, ptx::sw128_desc(a_lo + dy * DY_STEP + kk * 2), ptx::sw128_desc(b_lo + kk * 2), idesc, acc);
                else
                  ptx::umma_f16<1>(tacc, ptx::sw64_desc(a_lo + dy * DY_STEP + kk * 2), ptx::sw64_desc(b_lo + kk * 2), idesc, acc);
              }
            }
            ptx::umma_commit<1>(&empty[s]);
            if (dx == 2) ptx::umma_commit<1>(&t_full[buf]);
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      } else {
      for (int ks = 0; ks < ksteps; ++ks) {
        ptx::mbar_wait(&full[s], ph);
        ptx::tc_fence_after();
        const uint32_t a_lo = a_lo0 + s * (A_BYTES >> 4), b_lo = b_lo0 + s * (C::B_BYTES >> 4);
        if (ptx::elect_one()) {
#pragma unroll
          for (int kk = 0; kk < KC / 16; ++kk) {
            if constexpr (KC == 64)
              ptx::umma_f16<1>(tacc, ptx::sw128_desc(a_lo + kk * 2), ptx::sw128_desc(b_lo + kk * 2), idesc, (ks | kk) != 0);
            else
              ptx::umma_f16<1>(tacc, ptx::sw64_desc(a_lo + kk * 2), ptx::sw64_desc(b_lo + kk * 2), idesc, (ks | kk) != 0);
          }
          ptx::umma_commit<1>(&empty[s]);
          if (ks == ksteps - 1) ptx::umma_commit<1>(&t_full[buf]);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      }
    }
  } else if (warp >= 4 && ((warp - 4) >> 2) < WPQ) {
    // ===================================================== epilogue: TMEM -> BN -> act -> global (direct 16 B stores)
    const int e = warp - 4;
    const int quarter = e & 3, chalf = e >> 2;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    constexpr int GROUPS = CW / 32;   // 32-column groups per warp
    uint32_t unit_it = 0;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++unit_it) {
      const int nt = u % g.n_tiles, mt = u / g.n_tiles;
      const int buf = unit_it & 1;
      ptx::mbar_wait(&t_full[buf], (unit_it >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t tacc = tmem_base + lane_base + buf * TBUF + chalf * CW;
      const int tw = mt % g.tiles_w, th = (mt / g.tiles_w) % g.tiles_h, nb = mt / (g.tiles_w * g.tiles_h);
      const int r = quarter * 32 + lane;
      const int wi = r & ((1 << g.lw) - 1), hi = (r >> g.lw) & ((1 << g.lh) - 1), ni = r >> (g.lw + g.lh);
      const int img = (nb << (7 - g.lw - g.lh)) + ni;
      const int ph = (th << g.lh) + hi, pw = (tw << g.lw) + wi;
      // bricks are powers of two and may overhang the image (sizes that are not powers of two): overhanging rows read
      // zeros through TMA and are never stored
      const bool row_ok = img < g.n_images && pw < (MODE == 3 ? g.W >> 1 : g.W) && ph < (MODE == 3 ? g.H >> 1 : g.H);
      // MODE 3: (ph, pw) already are pooled coordinates (the brick tiles the pooled grid)
      const int64_t p = MODE == 2 ? ((int64_t)img * (g.H >> 1) + (ph >> 1)) * (g.W >> 1) + (pw >> 1)
                                  : ((int64_t)img * g.H + ph) * g.W + pw;
      // RH kernels have a single N tile, so the values never change and the four quarter warps of a column group may
      // share (and redundantly write) one copy; otherwise every warp keeps its own
      float* my_ss = smem_ss + (RH ? chalf : e) * 64;
      if constexpr (SS_SMEM) {
        static_assert(!SS_SMEM || CW == 32, "one 32-column group per epilogue warp");
        if (unit_it == 0 || g.n_tiles > 1) {
          __syncwarp();
          my_ss[lane] = __ldg(scale + nt * BN + chalf * CW + lane);
          my_ss[32 + lane] = __ldg(shift + nt * BN + chalf * CW + lane);
          __syncwarp();
        }
      }
      int64_t prow[4];   // MODE 0: pixel index of the 4 rows this lane writes out (8 * j + lane / 4), -1 if past the batch
      if constexpr (MODE == 0 || MODE == 3) {
        const int oH = MODE == 3 ? g.H >> 1 : g.H, oW = MODE == 3 ? g.W >> 1 : g.W;   // grid the brick tiles
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int rr = quarter * 32 + 8 * j + (lane >> 2);
          const int rwi = rr & ((1 << g.lw) - 1), rhi = (rr >> g.lw) & ((1 << g.lh) - 1), rni = rr >> (g.lw + g.lh);
          const int rimg = (nb << (7 - g.lw - g.lh)) + rni;
          const int rph = (th << g.lh) + rhi, rpw = (tw << g.lw) + rwi;
          prow[j] = (rimg < g.n_images && rph < oH && rpw < oW) ? ((int64_t)rimg * oH + rph) * oW + rpw : -1;
        }
      }
#pragma unroll 1
      for (int gi = 0; gi < GROUPS; ++gi) {
        const int co0 = nt * BN + chalf * CW + gi * 32;
        uint32_t v[32];
        if constexpr (MODE != 3) {
          ptx::tmem_ld16(tacc + gi * 32, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
          ptx::tmem_ld16(tacc + gi * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
          ptx::tmem_ld_wait();
        }
        const float4* sc4 = SS_SMEM ? reinterpret_cast<const float4*>(my_ss) : reinterpret_cast<const float4*>(scale + co0);
        const float4* sh4 = SS_SMEM ? reinterpret_cast<const float4*>(my_ss + 32) : reinterpret_cast<const float4*>(shift + co0);
        auto ld4 = [&](const float4* q) -> float4 {
          if constexpr (SS_SMEM) return *q; else return __ldg(q);
        };
        if constexpr (MODE == 2) {
          // max-pool BEFORE shift + activation (both monotone non-decreasing): max_window act(s*a + t) ==
          // act(max_window(s*a) + t); each lane of a 2 x 2 quad then finishes only its own 8 of the 32 channels
          const int wt = 1 << g.lw;
          float m[32];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 sc = ld4(sc4 + i);
            const float s4[4] = {sc.x, sc.y, sc.z, sc.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float t = __uint_as_float(v[4 * i + j]) * s4[j];
              t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, 1));
              t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, wt));
              m[4 * i + j] = t;
            }
          }
          const int sub = (wi & 1) | ((hi & 1) << 1);
          const float4 sha = ld4(sh4 + sub * 2), shb = ld4(sh4 + sub * 2 + 1);
          const float sh8[8] = {sha.x, sha.y, sha.z, sha.w, shb.x, shb.y, shb.z, shb.w};
          float y[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float t = sub == 0 ? m[j] : sub == 1 ? m[8 + j] : sub == 2 ? m[16 + j] : m[24 + j];
            y[j] = act2d<ACT>(t + sh8[j]);
          }
          if (row_ok)
            *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(out) + p * g.cout_pad + co0 + sub * 8) =
                make_uint4(pack2<FMT>(y[0], y[1]), pack2<FMT>(y[2], y[3]), pack2<FMT>(y[4], y[5]), pack2<FMT>(y[6], y[7]));
        } else if constexpr (MODE == 1) {
          float* dst = reinterpret_cast<float*>(out) + p * g.cout_real + co0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 sc = ld4(sc4 + i), sh = ld4(sh4 + i);
            float x[4];
            x[0] = act2d<ACT>(fmaf(__uint_as_float(v[4 * i]), sc.x, sh.x));
            x[1] = act2d<ACT>(fmaf(__uint_as_float(v[4 * i + 1]), sc.y, sh.y));
            x[2] = act2d<ACT>(fmaf(__uint_as_float(v[4 * i + 2]), sc.z, sh.z));
            x[3] = act2d<ACT>(fmaf(__uint_as_float(v[4 * i + 3]), sc.w, sh.w));
            if (row_ok) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (co0 + 4 * i + j < g.cout_real) dst[4 * i + j] = x[j];
            }
          }
        } else {
          uint32_t o[16];
          if constexpr (MODE == 3) {
            // four accumulators (one per pool-window position) in four BN-column blocks of this lane: the pool is a
            // per-thread max of scale * acc, taken before the BN shift and the activation (both monotone)
#pragma unroll
            for (int h16 = 0; h16 < 2; ++h16) {
              uint32_t a4[4][16];
#pragma unroll
              for (int a = 0; a < 4; ++a) ptx::tmem_ld16(tacc + a * BN + gi * 32 + h16 * 16, a4[a]);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 sc = ld4(sc4 + h16 * 4 + i), sh = ld4(sh4 + h16 * 4 + i);
                const float s4[4] = {sc.x, sc.y, sc.z, sc.w}, t4[4] = {sh.x, sh.y, sh.z, sh.w};
                float y[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int c = 4 * i + j;
                  const float m = fmaxf(fmaxf(__uint_as_float(a4[0][c]) * s4[j], __uint_as_float(a4[1][c]) * s4[j]),
                                        fmaxf(__uint_as_float(a4[2][c]) * s4[j], __uint_as_float(a4[3][c]) * s4[j]));
                  y[j] = act2d<ACT>(m + t4[j]);
                }
                o[h16 * 8 + 2 * i] = pack2<FMT>(y[0], y[1]);
                o[h16 * 8 + 2 * i + 1] = pack2<FMT>(y[2], y[3]);
              }
            }
          } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 sc = ld4(sc4 + i), sh = ld4(sh4 + i);
            const float x0 = act2d<ACT>(fmaf(__uint_as_float(v[4 * i]), sc.x, sh.x));
            const float x1 = act2d<ACT>(fmaf(__uint_as_float(v[4 * i + 1]), sc.y, sh.y));
            const float x2 = act2d<ACT>(fmaf(__uint_as_float(v[4 * i + 2]), sc.z, sh.z));
            const float x3 = act2d<ACT>(fmaf(__uint_as_float(v[4 * i + 3]), sc.w, sh.w));
            o[2 * i] = pack2<FMT>(x0, x1);
            o[2 * i + 1] = pack2<FMT>(x2, x3);
          }
          }
          {
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

template <int BN, int KC, int FMT, int MODE, bool RH>
int launch_bn(const CUtensorMap& ta, const CUtensorMap& tw, void* out, const float* scale, const float* shift,
              const Conv2dGeom& g, int act, int grid, cudaStream_t st) {
  constexpr int kSmemBytes = RH ? CfgRH<BN, KC, MODE>::SMEM_BYTES : Cfg<BN, KC>::SMEM_BYTES;
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
  auto launch = [&](auto kern) -> int {
    A3D_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    A3D_CUDA_OK(launch_chain(kern, dim3(grid), dim3(kThreads), kSmemBytes, st, 1, ta, tw, out, scale, shift, g));
    return A3D_OK;
  };
  switch (act) {
    case A3D_ACT_ELU: return launch(conv2d_tc_kernel<BN, KC, FMT, A3D_ACT_ELU, MODE, RH>);
    case A3D_ACT_RELU: return launch(conv2d_tc_kernel<BN, KC, FMT, A3D_ACT_RELU, MODE, RH>);
    case A3D_ACT_LRELU: return launch(conv2d_tc_kernel<BN, KC, FMT, A3D_ACT_LRELU, MODE, RH>);
    case A3D_ACT_LRELU01: return launch(conv2d_tc_kernel<BN, KC, FMT, A3D_ACT_LRELU01, MODE, RH>);
    case A3D_ACT_NONE: return launch(conv2d_tc_kernel<BN, KC, FMT, A3D_ACT_NONE, MODE, RH>);
    default: set_error("conv2d: unsupported activation %d", act); return A3D_ERR_INVALID;
  }
}

template <int BN, int KC, bool RH>
int launch_fmt(const CUtensorMap& ta, const CUtensorMap& tw, void* out, const float* scale, const float* shift,
               const Conv2dGeom& g, int fmt, int act, int mode, int grid, cudaStream_t st) {
  if constexpr (RH) {   // resident-weight variant: 16-bit outputs only (plain or pooled)
    if (mode == 1) { set_error("conv2d: the resident-weight variant has no fp32 output mode"); return A3D_ERR_INVALID; }
    if constexpr (BN == 64) {
      if (mode == 3)
        return fmt == A3D_DTYPE_F16 ? launch_bn<BN, KC, A3D_DTYPE_F16, 3, true>(ta, tw, out, scale, shift, g, act, grid, st)
                                    : launch_bn<BN, KC, A3D_DTYPE_BF16, 3, true>(ta, tw, out, scale, shift, g, act, grid, st);
    }
    if (mode == 3) { set_error("conv2d: the four-accumulator pool is built for 64-wide N tiles only"); return A3D_ERR_INVALID; }
    if (fmt == A3D_DTYPE_F16)
      return mode == 0 ? launch_bn<BN, KC, A3D_DTYPE_F16, 0, true>(ta, tw, out, scale, shift, g, act, grid, st)
                       : launch_bn<BN, KC, A3D_DTYPE_F16, 2, true>(ta, tw, out, scale, shift, g, act, grid, st);
    return mode == 0 ? launch_bn<BN, KC, A3D_DTYPE_BF16, 0, true>(ta, tw, out, scale, shift, g, act, grid, st)
                     : launch_bn<BN, KC, A3D_DTYPE_BF16, 2, true>(ta, tw, out, scale, shift, g, act, grid, st);
  } else {
    if (fmt == A3D_DTYPE_F16) {
      if (mode == 0) return launch_bn<BN, KC, A3D_DTYPE_F16, 0, false>(ta, tw, out, scale, shift, g, act, grid, st);
      if (mode == 1) return launch_bn<BN, KC, A3D_DTYPE_F16, 1, false>(ta, tw, out, scale, shift, g, act, grid, st);
      return launch_bn<BN, KC, A3D_DTYPE_F16, 2, false>(ta, tw, out, scale, shift, g, act, grid, st);
    }
    if (mode == 0) return launch_bn<BN, KC, A3D_DTYPE_BF16, 0, false>(ta, tw, out, scale, shift, g, act, grid, st);
    if (mode == 1) return launch_bn<BN, KC, A3D_DTYPE_BF16, 1, false>(ta, tw, out, scale, shift, g, act, grid, st);
    return launch_bn<BN, KC, A3D_DTYPE_BF16, 2, false>(ta, tw, out, scale, shift, g, act, grid, st);
  }
}

}  // namespace

int conv2d_tc_bn(int cout_pad) { return cout_pad % 256 == 0 ? 256 : (cout_pad % 128 == 0 ? 128 : 64); }

int launch_conv2d_tc(const CUtensorMap& tmap_act, const CUtensorMap& tmap_wgt, void* out, const float* scale,
                     const float* shift, const Conv2dGeom& g, int bn, int fmt, int act, bool pool, bool out_f32,
                     int num_sms, cudaStream_t st, int64_t* launches) {
  if (g.n_images <= 0) return A3D_OK;
  if (pool && g.rh != 2 && (out_f32 || g.lw > 4 || g.lh < 1 || (g.H & 1) || (g.W & 1))) {
    set_error("conv2d: fused pool needs a brick at most 16 wide and at least 2 high on even sizes");
    return A3D_ERR_INVALID;
  }
  const int total = g.m_tiles * g.n_tiles;
  const int grid = total < num_sms ? total : num_sms;
  const int mode = out_f32 ? 1 : (pool ? (g.rh == 2 ? 3 : 2) : 0);
  int rc;
  if (g.rh) {
    // resident weights + haloed activation box: 3 x 3, one Cin chunk, one N tile, 16 x 8 brick inside one image
    if (g.taps != 9 || g.cin_chunks != 1 || g.n_tiles != 1 || g.lw != 4 || g.lh != 3) {
      set_error("conv2d: geometry not eligible for the resident-weight variant");
      return A3D_ERR_INVALID;
    }
    if (g.kc == 32 && bn == 64) rc = launch_fmt<64, 32, true>(tmap_act, tmap_wgt, out, scale, shift, g, fmt, act, mode, grid, st);
    else if (g.kc == 64 && bn == 128) rc = launch_fmt<128, 64, true>(tmap_act, tmap_wgt, out, scale, shift, g, fmt, act, mode, grid, st);
    else { set_error("conv2d: no resident-weight kernel for kc %d / N tile %d", g.kc, bn); return A3D_ERR_INVALID; }
  } else if (g.kc == 32) {
    if (bn != 64) { set_error("conv2d: 32-channel K steps are built for 64-wide N tiles only"); return A3D_ERR_INVALID; }
    rc = launch_fmt<64, 32, false>(tmap_act, tmap_wgt, out, scale, shift, g, fmt, act, mode, grid, st);
  } else if (bn == 256) rc = launch_fmt<256, 64, false>(tmap_act, tmap_wgt, out, scale, shift, g, fmt, act, mode, grid, st);
  else if (bn == 128) rc = launch_fmt<128, 64, false>(tmap_act, tmap_wgt, out, scale, shift, g, fmt, act, mode, grid, st);
  else rc = launch_fmt<64, 64, false>(tmap_act, tmap_wgt, out, scale, shift, g, fmt, act, mode, grid, st);
  if (rc == A3D_OK && launches) ++*launches;
  return rc;
}

}  // namespace a3d
