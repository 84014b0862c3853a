// Stride-2 transposed 3-D convolution (k = 4, padding 'same') + folded BatchNorm + activation as a tcgen05
// implicit GEMM for sm_100a.  Replaces conv3DDec(), /root/reference/src/net_core/autoencoder3D.py:41-54, for the
// stride-2 hidden layers (512->256, 256->128; the 128->64 layer normally runs in convt_l4_sw.cu).
//
// Formulation.  Output voxel o = 2j + p (p = parity per axis) receives input voxels j + delta with tap
// t = p + 1 - 2*delta:   p = 0: delta in {-1, 0} (taps 3, 1);   p = 1: delta in {0, +1} (taps 2, 0).
// Each output-parity class is therefore a 2x2x2 gather conv over the input grid:
//     D[(n, j), co] = sum_{delta, ci} X[n, j + delta, ci] * W[t(p, delta), co, ci]        (zero outside the grid)
//
// Tiling.  One work unit = one input row (fixed d, h; all W positions along w) x NT = 128 / W decodes = 128 GEMM rows,
// for a fixed (pd, ph) and BOTH pw parities (COUT <= 128) or one pw (COUT = 256).  GEMM rows are ordered
// (w major, decode minor), so a shift of the input by delta_w = +-1 is a shift by NT rows = a multiple of the
// 1024-byte swizzle atom: ONE TMA box (64 ch, NT decodes, W + 2 positions incl. zero-filled halo) serves all three
// delta_w via the start address of the UMMA shared-memory descriptor.  The out-of-bounds fill of TMA is the
// 'same' padding.  delta_w = 0 feeds both pw parities in a single MMA of N = 2*COUT.
//
// PAIR = 2 (default for 512->256 and 256->128): two CTAs of a cluster run the same (d, h, parity class) on two adjacent
// decode blocks as ONE tcgen05 cta_group::2 MMA of M = 256.  They share every weight tile: each CTA loads and holds
// only its N-half (contiguous row ranges of the repacked weights), which halves the TMA weight writes and the
// B-operand reads -- these layers are bound by shared-memory bandwidth (MMA operand reads + TMA writes), not by the
// tensor pipe.  PAIR = 1 keeps the single-CTA path (A3D_CONV_PAIR=1, and the 128->64 layer with A3D_L4_IMPL=generic);
// HP pairs two input rows of ONE decode block instead (small calls); the launcher picks the variant per call size.
//
// Roles (384 threads): warps 0..7 = epilogue, warp 8 = TMEM allocator, warp 10 = TMA producer, warp 11 = MMA issuer
// (converged warp; only tcgen05.mma / commit are predicated on an elected lane, so descriptors stay in uniform
// registers).  fp32 accumulators are multi-buffered in TMEM (512 columns) so the epilogue of unit i overlaps the main
// loop of unit i + 1.  Epilogue: tcgen05.ld -> scale/shift -> activation -> 16-bit -> XOR-swizzled 128-byte staging
// rows in shared memory -> every warp-level store writes four complete 128-byte lines.  Persistent CTAs; static unit
// schedule in walk.h (regular workers keep their parity class and rotate through the positions, helper workers use the
// SMs a multiple of the class count leaves over); programmatic dependent launch (the prologue overlaps the previous layer).
#include <cstdlib>

#include "epilogue.cuh"
#include "internal.h"
#include "ptx.cuh"
#include "walk.h"

namespace a3d {

namespace {

template <int CIN_, int COUT_, int WIN_, int PAIR_, bool HP_ = false>
struct Cfg {
  static constexpr int CIN = CIN_, COUT = COUT_, WIN = WIN_, PAIR = PAIR_;
  // HP (PAIR = 2 only): the two CTAs of a pair take the input rows h = 2 hp and 2 hp + 1 of the SAME decode block instead of
  // the same row of two decode blocks -- for calls with an odd / single decode block (32 latents are ONE block of the
  // 512->256 layer), where the decode pairing runs half empty and the single-CTA kernel is bound by streaming every
  // weight tile to every SM.  The two rows see different h borders, so no (sd, sh) step is skipped in h: the row outside
  // the grid arrives as a zero tile (TMA out-of-bounds fill), 1 / (2 W) of the MACs.
  static constexpr bool HP = HP_;
  static_assert(!HP_ || PAIR_ == 2, "HP pairs two CTAs");
  static constexpr int POS_PER_BLOCK = HP_ ? WIN_ * WIN_ / 2 : WIN_ * WIN_;   // schedule positions per decode-block group
  static constexpr int NT = 128 / WIN;                    // decodes per unit
  static constexpr bool PWB = (COUT <= 128);              // both pw parities in one unit
  static constexpr int NPAR = PWB ? 4 : 8;                // parity classes per position
  static constexpr int NACC = PWB ? 2 * COUT : COUT;      // fp32 accumulator columns per unit
  static constexpr int NBUF = 512 / NACC;                 // accumulator buffers in TMEM (4 for Cout = 64, else 2)
  static constexpr int TMEM_COLS = 512;
  static constexpr int CHUNKS = CIN / 64;                 // 64-channel K chunks
  static constexpr int A_BYTES = (WIN + 2) * NT * 128;    // one input row incl. halo, one chunk
  static constexpr int BROWS = PWB ? 4 * COUT : 2 * COUT; // weight rows per (sd, sh, chunk)
  static constexpr int BSLOT_ROWS = 256;                  // weight rows per slot (whole MMA groups)
  static constexpr int BSLOTS = BROWS / BSLOT_ROWS;       // weight slots per input row
  static constexpr int B_BYTES = BSLOT_ROWS * 128 / PAIR; // bytes of a slot held by ONE CTA
  static constexpr int A_STAGES = (PAIR == 2) ? 4 : 3;
  static constexpr int B_STAGES = (PAIR == 2) ? 5 : ((COUT == 256) ? 3 : 4);
  static constexpr int OUT_STAGE_BYTES = 8 * 4096;        // 32 rows x 128 B per epilogue warp
  static constexpr int NUM_BARS = 2 * A_STAGES + 2 * B_STAGES + 2 * NBUF;
  static constexpr int SMEM_BYTES = 1024 /*align slack*/ + A_STAGES * A_BYTES + B_STAGES * B_BYTES + OUT_STAGE_BYTES +
                                    NUM_BARS * 8 + 16 + 2 * NACC * 4;
  static_assert(BROWS % BSLOT_ROWS == 0, "weight rows per input row must fill whole slots");
  static_assert(A_BYTES % 1024 == 0 && (NT * 128) % 1024 == 0, "shifted A views must stay atom aligned");
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB shared-memory limit");
  static_assert(!(PAIR == 2 && COUT == 64), "the 128->64 layer has its own 2-CTA kernel (convt_l4_sw.cu)");
};

constexpr int kEpiWarps = 8;                 // warps 0..7, 2 per scheduler: warps e and e + 4 share a TMEM lane quarter
constexpr int kThreads = 128 + 32 * kEpiWarps;
constexpr int kWarpAlloc = 8, kWarpTma = 10, kWarpMma = 11;

template <class C, int FMT, int ACT>
__global__ void __launch_bounds__(kThreads, 1)
convt_s2_tc_kernel(const __grid_constant__ CUtensorMap tmap_act, const __grid_constant__ CUtensorMap tmap_wgt,
                   uint16_t* __restrict__ out, const float* __restrict__ scale, const float* __restrict__ shift,
                   int n_blocks, int n_alloc) {
  constexpr int COUT = C::COUT, WIN = C::WIN, NT = C::NT, NACC = C::NACC, PAIR = C::PAIR;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128-byte swizzle, computed on the shared-window address so the pointer keeps its
  // __shared__ provenance (LDS/STS instead of generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::A_STAGES * C::A_BYTES;
  uint8_t* smem_o = smem_b + C::B_STAGES * C::B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_o + C::OUT_STAGE_BYTES);
  uint64_t* a_full = bars;                       // PAIR = 2: used on the leader CTA (bytes of both CTAs)
  uint64_t* a_empty = a_full + C::A_STAGES;      // per CTA (multicast commit)
  uint64_t* b_full = a_empty + C::A_STAGES;
  uint64_t* b_empty = b_full + C::B_STAGES;
  uint64_t* t_full = b_empty + C::B_STAGES;      // per CTA (multicast commit)
  uint64_t* t_empty = t_full + C::NBUF;          // PAIR = 2: used on the leader CTA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + C::NBUF);
  float* s_scale = reinterpret_cast<float*>(tmem_slot + 4);
  float* s_shift = s_scale + NACC;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = (PAIR == 2) ? (int)ptx::cluster_ctarank() : 0;
  const int cl_id = blockIdx.x / PAIR;           // cluster (or CTA) index
  const int n_cl = gridDim.x / PAIR;
  const int nb_groups = C::HP ? n_blocks : (n_blocks + PAIR - 1) / PAIR;
  const Walk walk = make_walk(C::NPAR, nb_groups * C::POS_PER_BLOCK, n_cl);
  // position -> (input row h, depth d, decode block) of THIS CTA
  auto decode_pos = [&](int pos, int& h, int& d, int& nb) {
    if constexpr (C::HP) {
      h = 2 * (pos % (WIN / 2)) + rank;
      d = (pos / (WIN / 2)) % WIN;
      nb = pos / C::POS_PER_BLOCK;
    } else {
      h = pos % WIN;
      d = (pos / WIN) % WIN;
      nb = (pos / (WIN * WIN)) * PAIR + rank;
    }
  };
  const int my_units = walk.count(cl_id);

  if (warp == kWarpTma && lane == 0) {
    ptx::prefetch_tmap(&tmap_act);
    ptx::prefetch_tmap(&tmap_wgt);
  }
  if (warp == kWarpMma && lane == 0) {
    // full barriers take ONE arrival: the (leader's) arrive.expect_tx for the bytes of all CTAs of the pair; the
    // peer's TMA only contributes complete_tx bytes (a transiently negative tx-count is legal)
    for (int i = 0; i < C::A_STAGES; ++i) { ptx::mbar_init(&a_full[i], 1); ptx::mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < C::B_STAGES; ++i) { ptx::mbar_init(&b_full[i], 1); ptx::mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < C::NBUF; ++i) { ptx::mbar_init(&t_full[i], 1); ptx::mbar_init(&t_empty[i], PAIR * kEpiWarps); }
    ptx::fence_barrier_init();
  }
  for (int i = threadIdx.x; i < NACC; i += blockDim.x) {
    s_scale[i] = scale[i % COUT];
    s_shift[i] = shift[i % COUT];
  }
  if constexpr (PAIR == 2) ptx::cluster_sync_all();   // barrier inits visible before any remote arrive / multicast
  if (warp == kWarpAlloc) {
    ptx::tmem_alloc<PAIR>(tmem_slot, C::TMEM_COLS);
    ptx::tmem_relinquish<PAIR>();
  }
  ptx::tc_fence_before();
  if constexpr (PAIR == 2) ptx::cluster_sync_all(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_sync();     // the input activations are the previous kernel's output

  if (warp == kWarpTma) {
    // ===================================================== TMA producer (one per CTA)
    if (lane == 0) {
      uint32_t a_it = 0, b_it = 0;
      for (int k = 0; k < my_units; ++k) {
        int par, pos;
        if (!walk.unit(cl_id, k, par, pos)) continue;
        int h, d, nb;
        decode_pos(pos, h, d, nb);
        const int pd = C::PWB ? (par >> 1) : (par >> 2);
        const int ph = C::PWB ? (par & 1) : ((par >> 1) & 1);
        for (int sd = 0; sd < 2; ++sd) {
          const int id = d + sd - 1 + pd;
          if (id < 0 || id >= WIN) continue;
          for (int sh = 0; sh < 2; ++sh) {
            const int ih = h + sh - 1 + ph;
            if (!C::HP && (ih < 0 || ih >= WIN)) continue;     // HP: a row outside the grid is a zero tile (TMA fill)
            for (int c = 0; c < C::CHUNKS; ++c) {
              const int as = a_it % C::A_STAGES;
              ptx::mbar_wait(&a_empty[as], ((a_it / C::A_STAGES) & 1) ^ 1);
              if (rank == 0) ptx::mbar_expect_tx(&a_full[as], PAIR * C::A_BYTES);
              if constexpr (PAIR == 2)
                ptx::tma_load_5d_2sm(smem_a + as * C::A_BYTES, &tmap_act, &a_full[as], c * 64, nb * NT, -1, ih, id);
              else
                ptx::tma_load_5d(smem_a + as * C::A_BYTES, &tmap_act, &a_full[as], c * 64, nb * NT, -1, ih, id);
              ++a_it;
              const int row0 = ((((par * 2 + sd) * 2 + sh) * C::CHUNKS) + c) * C::BROWS;
#pragma unroll
              for (int j = 0; j < C::BSLOTS; ++j) {
                const int bs = b_it % C::B_STAGES;
                uint8_t* dst = smem_b + bs * C::B_BYTES;
                ptx::mbar_wait(&b_empty[bs], ((b_it / C::B_STAGES) & 1) ^ 1);
                if (rank == 0) ptx::mbar_expect_tx(&b_full[bs], PAIR * C::B_BYTES);
                const int r0 = row0 + j * C::BSLOT_ROWS;
                if constexpr (PAIR == 1) {
                  ptx::tma_load_2d(dst, &tmap_wgt, &b_full[bs], 0, r0);
                } else if (C::PWB && j == 1) {
                  // COUT = 128, slot 1: dw=-1 tile rows [0,128), dw=+1 tile rows [128,256): this CTA's 64-row N-halves
                  ptx::tma_load_2d_2sm(dst, &tmap_wgt, &b_full[bs], 0, r0 + rank * 64);
                  ptx::tma_load_2d_2sm(dst + 64 * 128, &tmap_wgt, &b_full[bs], 0, r0 + 128 + rank * 64);
                } else {
                  // one N = 256 MMA: this CTA's N-half = 128 contiguous rows (two 64-row boxes)
                  ptx::tma_load_2d_2sm(dst, &tmap_wgt, &b_full[bs], 0, r0 + rank * 128);
                  ptx::tma_load_2d_2sm(dst + 64 * 128, &tmap_wgt, &b_full[bs], 0, r0 + rank * 128 + 64);
                }
                ++b_it;
              }
            }
          }
        }
      }
    }
  } else if (warp == kWarpMma) {
    // ===================================================== MMA issuer (leader CTA only when PAIR = 2)
    if (rank == 0) {
      uint32_t a_it = 0, b_it = 0, unit_it = 0;
      constexpr int MM = 128 * PAIR;
      constexpr uint32_t idesc_full = ptx::make_idesc_f16(MM, NACC > 256 ? 256 : NACC, FMT);
      constexpr uint32_t idesc_half = ptx::make_idesc_f16(MM, COUT, FMT);
      const uint32_t a_lo0 = ptx::sw128_desc_lo(ptx::smem_u32(smem_a));
      const uint32_t b_lo0 = ptx::sw128_desc_lo(ptx::smem_u32(smem_b));
      for (int k = 0; k < my_units; ++k) {
        int par, pos;
        if (!walk.unit(cl_id, k, par, pos)) continue;
        int h, d, nb_unused;
        decode_pos(pos, h, d, nb_unused);
        const int pd = C::PWB ? (par >> 1) : (par >> 2);
        const int ph = C::PWB ? (par & 1) : ((par >> 1) & 1);
        const int pw = par & 1;  // only meaningful when !PWB
        const int buf = unit_it % C::NBUF;
        ptx::mbar_wait(&t_empty[buf], ((unit_it / C::NBUF) & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t tacc = tmem_base + buf * NACC;
        uint32_t accum = 0;
        for (int sd = 0; sd < 2; ++sd) {
          const int id = d + sd - 1 + pd;
          if (id < 0 || id >= WIN) continue;
          for (int sh = 0; sh < 2; ++sh) {
            const int ih = h + sh - 1 + ph;
            if (!C::HP && (ih < 0 || ih >= WIN)) continue;
            for (int c = 0; c < C::CHUNKS; ++c) {
              const int as = a_it % C::A_STAGES;
              ptx::mbar_wait(&a_full[as], (a_it / C::A_STAGES) & 1);
              const uint32_t a_lo = a_lo0 + as * (C::A_BYTES >> 4);
#pragma unroll
              for (int j = 0; j < C::BSLOTS; ++j) {
                const int bs = b_it % C::B_STAGES;
                ptx::mbar_wait(&b_full[bs], (b_it / C::B_STAGES) & 1);
                ptx::tc_fence_after();
                const uint32_t b_lo = b_lo0 + bs * (C::B_BYTES >> 4);
                constexpr uint32_t W1 = (NT * 128) >> 4;       // one position along w, in 16-byte units
                // offsets (>> 4) of the B tiles inside this CTA's part of the slot
                constexpr uint32_t BB = ((2 * COUT / PAIR) * 128) >> 4;   // COUT = 64: start of the dw = -1 tile
                constexpr uint32_t BC = ((3 * COUT / PAIR) * 128) >> 4;   // COUT = 64: start of the dw = +1 tile
                constexpr uint32_t BH = ((COUT / PAIR) * 128) >> 4;       // COUT = 128, slot 1: start of the dw = +1 tile
                if (ptx::elect_one()) {
#pragma unroll
                  for (int kk = 0; kk < 4; ++kk) {
                    const uint32_t ko = kk * 2;  // 16 elements * 2 bytes inside the 128-byte swizzled row, >> 4
                    if constexpr (C::PWB && C::BSLOTS == 1) {
                      // COUT = 64: (pw0,tw1 | pw1,tw2) dw=0, N = 128;  pw0,tw3 dw=-1;  pw1,tw0 dw=+1
                      ptx::umma_f16<PAIR>(tacc, ptx::sw128_desc(a_lo + W1 + ko), ptx::sw128_desc(b_lo + ko), idesc_full,
                                          (kk == 0) ? accum : 1u);
                      ptx::umma_f16<PAIR>(tacc, ptx::sw128_desc(a_lo + ko), ptx::sw128_desc(b_lo + BB + ko), idesc_half, 1);
                      ptx::umma_f16<PAIR>(tacc + COUT, ptx::sw128_desc(a_lo + 2 * W1 + ko),
                                          ptx::sw128_desc(b_lo + BC + ko), idesc_half, 1);
                    } else if constexpr (C::PWB) {
                      // COUT = 128: slot 0 = (pw0,tw1 | pw1,tw2) dw=0 (N = 256); slot 1 = pw0,tw3 dw=-1 | pw1,tw0 dw=+1
                      if (j == 0) {
                        ptx::umma_f16<PAIR>(tacc, ptx::sw128_desc(a_lo + W1 + ko), ptx::sw128_desc(b_lo + ko), idesc_full,
                                            (kk == 0) ? accum : 1u);
                      } else {
                        ptx::umma_f16<PAIR>(tacc, ptx::sw128_desc(a_lo + ko), ptx::sw128_desc(b_lo + ko), idesc_half, 1);
                        ptx::umma_f16<PAIR>(tacc + COUT, ptx::sw128_desc(a_lo + 2 * W1 + ko),
                                            ptx::sw128_desc(b_lo + BH + ko), idesc_half, 1);
                      }
                    } else {
                      // COUT = 256, one pw per unit: slot 0 = dw=0 tap, slot 1 = dw=+-1 tap
                      const uint32_t a_off = (j == 0) ? W1 : (pw ? 2 * W1 : 0);
                      ptx::umma_f16<PAIR>(tacc, ptx::sw128_desc(a_lo + a_off + ko), ptx::sw128_desc(b_lo + ko), idesc_full,
                                          (kk == 0) ? accum : 1u);
                    }
                  }
                  ptx::umma_commit<PAIR>(&b_empty[bs]);  // slot reusable (in both CTAs) once these MMAs retire
                  if (j == C::BSLOTS - 1) ptx::umma_commit<PAIR>(&a_empty[as]);
                }
                __syncwarp();
                if (j == 0) accum = 1;
                ++b_it;
              }
              ++a_it;
            }
          }
        }
        if (ptx::elect_one()) ptx::umma_commit<PAIR>(&t_full[buf]);  // accumulators complete -> epilogue(s)
        __syncwarp();
        ++unit_it;
      }
    }
  } else if (warp < kEpiWarps) {
    // ===================================================== epilogue: TMEM -> BN/act -> 16-bit -> smem staging -> global
    const int e = warp;
    const int quarter = e & 3;                         // TMEM lane quarter this warp may read (== warp % 4)
    const int chalf = e >> 2;                          // which half of the accumulator columns
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    constexpr int OD = 2 * WIN;
    constexpr int NCOLS = NACC / 2;                    // columns per warp
    uint8_t* stage = smem_o + e * 4096;                // this warp's 32 rows x 128 B staging tile
    uint32_t unit_it = 0;
    for (int k = 0; k < my_units; ++k) {
      int par, pos;
      if (!walk.unit(cl_id, k, par, pos)) continue;
      int h, d, nb;
      decode_pos(pos, h, d, nb);
      const int pd = C::PWB ? (par >> 1) : (par >> 2);
      const int ph = C::PWB ? (par & 1) : ((par >> 1) & 1);
      const int buf = unit_it % C::NBUF;
      ptx::mbar_wait(&t_full[buf], (unit_it / C::NBUF) & 1);
      ptx::tc_fence_after();
      const uint32_t tacc = tmem_base + lane_base + buf * NACC + chalf * NCOLS;
#pragma unroll 1
      for (int ch = 0; ch < NCOLS / 64; ++ch) {         // 64 output channels (128 B per row) at a time
        const int col0 = chalf * NCOLS + ch * 64;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t v[32];
          ptx::tmem_ld16(tacc + ch * 64 + half * 32, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
          ptx::tmem_ld16(tacc + ch * 64 + half * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
          ptx::tmem_ld_wait();
          const int col = col0 + half * 32;
          uint32_t o[16];
          const float4* sc4 = reinterpret_cast<const float4*>(s_scale + col);
          const float4* sh4 = reinterpret_cast<const float4*>(s_shift + col);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 sc = sc4[i], sh = sh4[i];
            const float x0 = activate<ACT>(fmaf(__uint_as_float(v[4 * i]), sc.x, sh.x));
            const float x1 = activate<ACT>(fmaf(__uint_as_float(v[4 * i + 1]), sc.y, sh.y));
            const float x2 = activate<ACT>(fmaf(__uint_as_float(v[4 * i + 2]), sc.z, sh.z));
            const float x3 = activate<ACT>(fmaf(__uint_as_float(v[4 * i + 3]), sc.w, sh.w));
            o[2 * i] = pack2<FMT>(x0, x1);
            o[2 * i + 1] = pack2<FMT>(x2, x3);
          }
          // lane = row: 4 x 16 B into the row's 128-byte line, 16-byte chunks XOR-swizzled by (row & 7)
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            const int chunk = (half * 4 + c4) ^ (lane & 7);
            *reinterpret_cast<uint4*>(stage + lane * 128 + chunk * 16) =
                make_uint4(o[4 * c4], o[4 * c4 + 1], o[4 * c4 + 2], o[4 * c4 + 3]);
          }
        }
        if (ch == NCOLS / 64 - 1) {
          // all TMEM reads of this warp for the unit are done: release the accumulator buffer (one arrive per warp)
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (rank == 0) ptx::mbar_arrive(&t_empty[buf]);
            else ptx::mbar_arrive_cluster(&t_empty[buf], 0);
          }
        }
        __syncwarp();
        // 8 lanes per row: every warp-level store writes four complete 128-byte lines
        const int pw = C::PWB ? (col0 / COUT) : (par & 1);
        const int co = col0 % COUT;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = it * 4 + (lane >> 3);
          const int c16 = lane & 7;
          const uint4 val = *reinterpret_cast<const uint4*>(stage + r * 128 + ((c16 ^ (r & 7)) * 16));
          const int grow = quarter * 32 + r;
          const int wr = grow / NT, nr = nb * NT + grow % NT;
          if (nr < n_alloc) {
            A3D_DEV_CHECK(nr >= 0 && (unsigned)(2 * d + pd) < (unsigned)OD && (unsigned)(2 * h + ph) < (unsigned)OD &&
                          (unsigned)(2 * wr + pw) < (unsigned)OD && co >= 0 && co + c16 * 8 + 8 <= COUT);
            const size_t vox = (((size_t)nr * OD + (2 * d + pd)) * OD + (2 * h + ph)) * OD + (2 * wr + pw);
            __stcs(reinterpret_cast<uint4*>(out + vox * COUT + co + c16 * 8), val);   // streaming: keep L2 for the inputs
          }
        }
        __syncwarp();
      }
      ++unit_it;
    }
  }

  ptx::tc_fence_before();
  if constexpr (PAIR == 2) ptx::cluster_sync_all(); else __syncthreads();
  if (warp == kWarpAlloc) ptx::tmem_dealloc<PAIR>(tmem_base, C::TMEM_COLS);
}

template <class C>
int launch_cfg(const ConvLayer& L, void* out, int64_t n, int64_t n_alloc, int fmt, int act, int num_sms,
               cudaStream_t st) {
  constexpr int PAIR = C::PAIR;
  const int n_blocks = (int)((n + C::NT - 1) / C::NT);
  const int n_pos = (C::HP ? n_blocks : (n_blocks + PAIR - 1) / PAIR) * C::POS_PER_BLOCK;   // >= 8 >= NPAR
  int n_cl = num_sms / PAIR;
  if (n_cl > n_pos * C::NPAR) n_cl = n_pos * C::NPAR;
  // regular workers keep their parity class for the whole launch, the left-over workers help (see Walk); when helping
  // does not pay (tiny launches) they are not launched at all
  {
    const Walk w = make_walk(C::NPAR, n_pos, n_cl);
    n_cl = w.reg + w.helpers;
  }
  const CUtensorMap& tw = (PAIR == 2) ? L.tmap_wgt64 : L.tmap_wgt;
  auto launch = [&](auto kern) -> int {
    A3D_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    A3D_CUDA_OK(launch_chain(kern, dim3(n_cl * PAIR), dim3(kThreads), C::SMEM_BYTES, st, PAIR, L.tmap_act, tw,
                             reinterpret_cast<uint16_t*>(out), (const float*)L.scale, (const float*)L.shift, n_blocks,
                             (int)n_alloc));
    return A3D_OK;
  };
  if (fmt == A3D_DTYPE_F16) {
    switch (act) {
      case A3D_ACT_ELU: return launch(convt_s2_tc_kernel<C, A3D_DTYPE_F16, A3D_ACT_ELU>);
      case A3D_ACT_RELU: return launch(convt_s2_tc_kernel<C, A3D_DTYPE_F16, A3D_ACT_RELU>);
      case A3D_ACT_LRELU: return launch(convt_s2_tc_kernel<C, A3D_DTYPE_F16, A3D_ACT_LRELU>);
      default: return launch(convt_s2_tc_kernel<C, A3D_DTYPE_F16, A3D_ACT_NONE>);
    }
  }
  switch (act) {
    case A3D_ACT_ELU: return launch(convt_s2_tc_kernel<C, A3D_DTYPE_BF16, A3D_ACT_ELU>);
    case A3D_ACT_RELU: return launch(convt_s2_tc_kernel<C, A3D_DTYPE_BF16, A3D_ACT_RELU>);
    case A3D_ACT_LRELU: return launch(convt_s2_tc_kernel<C, A3D_DTYPE_BF16, A3D_ACT_LRELU>);
    default: return launch(convt_s2_tc_kernel<C, A3D_DTYPE_BF16, A3D_ACT_NONE>);
  }
}

}  // namespace

size_t convt_tc_smem_bytes(int cin, int cout, int win) {
  if (cin == 512 && cout == 256 && win == 4) return Cfg<512, 256, 4, 2>::SMEM_BYTES;
  if (cin == 256 && cout == 128 && win == 8) return Cfg<256, 128, 8, 2>::SMEM_BYTES;
  if (cin == 128 && cout == 64 && win == 16) return Cfg<128, 64, 16, 1>::SMEM_BYTES;
  return 0;
}

// Variant choice per call size.  The decode pairing (PAIR = 2) shares every weight tile between two decode blocks: best
// for the large launches, but it needs an even number of blocks and halves the number of schedulable workers.  For small
// calls (the reference's 32- and 72-latent decoder calls) the h pairing (two input rows of ONE block per pair) or the
// single-CTA kernel saves scheduling rounds.  Cost = units of the busiest worker x relative unit time: 1 for the decode
// pairing, 1 + 1 / (2 W) for the h pairing (its zero rows), `single_cost` for the single-CTA kernel (measured: it
// streams every weight tile to every SM and is bound by the L2 -> SM fabric on the 512->256 layer).
using conv_variant::kVarPair;
using conv_variant::kVarHp;
using conv_variant::kVarSingle;
template <class C2>
int pick_variant(int64_t n, int num_sms, float single_cost) {
  return conv_variant::pick(n, C2::NT, C2::WIN, C2::NPAR, num_sms, single_cost);
}

int launch_convt_s2_tc(const ConvLayer& L, void* out, int64_t n, int64_t n_alloc, int fmt, int act, int num_sms,
                       int force_variant, cudaStream_t st, int64_t* launches) {
  // force_variant (A3D_CONV_PAIR, read when the handle is created) = 1 / 2 / 3: single-CTA / decode-pair / h-pair variant
  const int forced = force_variant == 1 ? kVarSingle : force_variant == 2 ? kVarPair : force_variant == 3 ? kVarHp : -1;
  int rc;
  if (L.cin == 512 && L.cout == 256 && L.win == 4) {
    const int v = forced >= 0 ? forced : pick_variant<Cfg<512, 256, 4, 2>>(n, num_sms, 1.6f);
    rc = v == kVarSingle ? launch_cfg<Cfg<512, 256, 4, 1>>(L, out, n, n_alloc, fmt, act, num_sms, st)
         : v == kVarHp   ? launch_cfg<Cfg<512, 256, 4, 2, true>>(L, out, n, n_alloc, fmt, act, num_sms, st)
                         : launch_cfg<Cfg<512, 256, 4, 2>>(L, out, n, n_alloc, fmt, act, num_sms, st);
  } else if (L.cin == 256 && L.cout == 128 && L.win == 8) {
    const int v = forced >= 0 ? forced : pick_variant<Cfg<256, 128, 8, 2>>(n, num_sms, 1.15f);
    rc = v == kVarSingle ? launch_cfg<Cfg<256, 128, 8, 1>>(L, out, n, n_alloc, fmt, act, num_sms, st)
         : v == kVarHp   ? launch_cfg<Cfg<256, 128, 8, 2, true>>(L, out, n, n_alloc, fmt, act, num_sms, st)
                         : launch_cfg<Cfg<256, 128, 8, 2>>(L, out, n, n_alloc, fmt, act, num_sms, st);
  }
  else if (L.cin == 128 && L.cout == 64 && L.win == 16)
    rc = launch_cfg<Cfg<128, 64, 16, 1>>(L, out, n, n_alloc, fmt, act, num_sms, st);
  else {
    set_error("tcgen05 ConvT path supports (512->256,W4), (256->128,W8), (128->64,W16); got %d->%d W%d", L.cin,
              L.cout, L.win);
    return A3D_ERR_INVALID;
  }
  if (rc == A3D_OK && launches) ++*launches;
  return rc;
}

}  // namespace a3d
