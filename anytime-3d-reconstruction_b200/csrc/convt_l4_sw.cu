// 128 -> 64 stride-2 transposed conv (58.6 % of the decoder FLOPs, conv3DDec, autoencoder3D.py:41-54) as a
// weight-stationary 2-CTA tcgen05 kernel that SWEEPS ALONG W and resolves the w taps inside a ring of TMEM accumulators.
//
// Why: in the round-1 kernel (h-sweep, removed; profiles/r01_l4_ws2cta_ncu_full.txt) the delta_w = -1 / +1 taps were
// separate N = 64 MMAs over w-shifted views of the activation tile.  An N = 64 MMA reads 5 KB of shared-memory operands
// per 32 tensor clocks, more than the 128 B/clk the port delivers: that kernel sat at 81 % tensor-active, port bound.  Here every MMA is M = 256 (CTA pair), N = 256:
// 8 KB of operands per 128 tensor clocks per CTA = half the port.
//
// Formulation.  Output voxel w_o = 2 j_o + pw receives input column j = j_o + dw through tap tw = pw + 1 - 2 dw, so
// input column j feeds exactly four (output column, parity) blocks:
//     tap 0 -> (j-1, pw 1)    tap 1 -> (j, pw 0)    tap 2 -> (j, pw 1)    tap 3 -> (j+1, pw 0)
// GEMM rows are the 16 (h) x 8 (decodes) positions of ONE input column j (lane = position, identical for every j), the
// B operand of a K step is [tap0 | tap1 | tap2 | tap3] x 64 output channels (N = 256; cluster rank 0 holds taps 0-1,
// rank 1 taps 2-3), and the accumulator window of step j is four consecutive 64-column blocks of TMEM that OVERLAP the
// window of step j+1 by two blocks: the sum over dw happens in the accumulator, in the same lanes, for free.
//   * delta_h taps: the activation tile of a step holds h' = -1..16 (18 x 8 rows, TMA zero fill = 'same' padding) and
//     the two delta_h views are whole-swizzle-atom shifts of the A descriptor start address.
//   * delta_d taps: two tiles per step (d' = d + delta_d), skipped at the grid border.
//   * TMEM: 4 slots of 128 columns.  Windows may not wrap around column 512 (and the two N-halves of B are pinned to
//     the two CTAs, so a window cannot be split), hence a cycle of three steps A (slots 0,1) -> B (1,2) -> C (2,3) -> A:
//     the step after C starts fresh accumulators in slots 0,1 and the epilogue adds slot 3 (the partial sums C left for
//     the same outputs, same lanes) while draining slot 0.  Every drained slot is re-zeroed with tcgen05.st, so all
//     MMAs accumulate and no MMA ever has to mix fresh and live columns.
//   * Items (decode-block pair, d) are 16 steps; consecutive items of a cluster start alternately in phase A and C, so
//     the first window of an item never touches a slot the previous item's last drain still owns.
// Otherwise: clusters own one (pd, ph) output-parity class for the whole launch (128 KB of weights per CTA resident in
// shared memory, loaded before the programmatic-dependent-launch wait), activation tiles stream through a 4-stage TMA
// ring, the leader CTA's MMA warp drives both SMs, 16 epilogue warps per CTA (a drained slot is handed back first, then
// folded BN + activation + 16-bit pack + staged 32-byte-sector stores), item schedule in l4_sched.h, soft pacing of the
// clusters through per-cluster progress counters (L2 reuse of the activation planes).
#include <cstdlib>

#include "epilogue.cuh"
#include "internal.h"
#include "ptx.cuh"
#include "l4_sched.h"

namespace a3d {
namespace {

using l4::Seg;
using l4::Sched;
using l4::make_sched;
using l4::seg_item;
using l4::item_depth;
constexpr int COUT = 64, WIN = l4::WIN, NT = 8, CHUNKS = 2;
constexpr int A_BYTES = (WIN + 2) * NT * 128;   // 18432: rows (h' = -1..16, decode)
constexpr int A_STAGES = 4;
constexpr int OUT_STAGE_BYTES = 16 * 1024;      // per epilogue warp: 32 rows x 32 B (16 channels), XOR-swizzled
constexpr int SLICE_BYTES = 128 * 128;          // one (sd, sh, chunk) slice of one rank: 2 taps x 64 co rows
constexpr int W_BYTES = 8 * SLICE_BYTES;        // 131072
constexpr int NSLOT = 4, SLOT_COLS = 128;
constexpr int TMEM_COLS = NSLOT * SLOT_COLS;    // 512
constexpr int kEpiWarps = 16;
constexpr int kThreads = 128 + 32 * kEpiWarps;
constexpr int NUM_BARS = 2 * A_STAGES + 2 * NSLOT + 1;
constexpr int SMEM_BYTES = 1024 + W_BYTES + A_STAGES * A_BYTES + OUT_STAGE_BYTES + NUM_BARS * 8 + 16 + 2 * COUT * 4 + 16;

__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
  const uint32_t z = 0;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr),
      "r"(z)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Phase (0 = A, 2 = C) in which the sweep of an item starts.  A function of the item's depth only, so that the split of
// an output's accumulation into two partial sums (slot 3 + slot 0, fp32) depends on WHERE the output voxel is and
// never on its decode's position in the batch: a latent decodes to bit-identical values in any batch / shard.
// Regular clusters step d by 3 from item to item (round k -> k + 1 advances the item by per + 1 = 19 on 148 SMs), the
// helper clusters by 1, so consecutive items alternate A, C: an item that ends in phase C leaves slots 0,1 free for a
// start in A, and one that ends in A has drained slot 2 and frees slot 3 first.  Equal consecutive phases are legal (the
// MMA warp just waits for the drains).
__device__ __forceinline__ int start_phase(int d) { return (d & 1) ? 2 : 0; }

// ELU / ReLU / LeakyReLU / identity on a packed pair of fp32 values, then 16-bit pack.  ELU is branch- and
// predicate-free: max(v, 2^min(v log2e, 0) - 1) equals v for v > 0 (the exponential term is 0) and exp(v) - 1 below
// (exp(v) - 1 >= v everywhere); ex2.approx.ftz has 2^-22 relative error, i.e. <= 2.4e-7 absolute on exp(v) <= 1.
template <int FMT, int ACT>
__device__ __forceinline__ uint32_t act_pack2(uint64_t y) {
  float a, b;
  ptx::f2_unpack(y, a, b);
  if constexpr (ACT == A3D_ACT_ELU) {
    float ta, tb;
    ptx::f2_unpack(ptx::f2_mul(y, ptx::f2_pack(1.4426950408889634f, 1.4426950408889634f)), ta, tb);
    ta = fminf(ta, 0.f);
    tb = fminf(tb, 0.f);
    asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(ta));
    asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(tb));
    ptx::f2_unpack(ptx::f2_add(ptx::f2_pack(ta, tb), ptx::f2_pack(-1.f, -1.f)), ta, tb);
    a = fmaxf(a, ta);
    b = fmaxf(b, tb);
  } else {
    a = activate<ACT>(a);
    b = activate<ACT>(b);
  }
  return pack2<FMT>(a, b);
}

constexpr int kPaceStride = 64;    // ints between two clusters' counters: one 256-byte L2 line pair (own slice) per counter
constexpr long long kPaceTimeout = 300000;   // clocks (~0.2 ms) a producer waits for its peers before it goes on alone for a while
__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_gpu(int* p, int v) {
  asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int FMT, int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
convt_l4_sw_kernel(const __grid_constant__ CUtensorMap tmap_act, const __grid_constant__ CUtensorMap tmap_wgt,
                   uint16_t* __restrict__ out, const float* __restrict__ scale, const float* __restrict__ shift,
                   int n_blocks, int n_alloc, int* __restrict__ progress, int pace_delta) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem + W_BYTES;
  uint8_t* smem_o = smem_a + A_STAGES * A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_o + OUT_STAGE_BYTES);
  uint64_t* a_full = bars;                    // [A_STAGES]  leader: 1 arrival + both CTAs' bytes
  uint64_t* a_empty = a_full + A_STAGES;      // [A_STAGES]  per CTA, multicast commit
  uint64_t* s_full = a_empty + A_STAGES;      // [NSLOT]     per CTA, multicast commit: slot closed
  uint64_t* s_empty = s_full + NSLOT;         // [NSLOT]     leader: 2 x 16 epilogue warps: slot drained and re-zeroed
  uint64_t* w_full = s_empty + NSLOT;         // leader: weights of both CTAs landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
  float* s_scale = reinterpret_cast<float*>(tmem_slot + 2);   // 17 barriers + 8 bytes: 16-byte aligned for the float4 reads
  float* s_shift = s_scale + COUT;
  volatile int* pace_allowed = reinterpret_cast<volatile int*>(s_shift + COUT);   // last step the producer may issue
  volatile int* pace_done = pace_allowed + 1;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int n_items = ((n_blocks + 1) >> 1) * WIN;    // (decode-block pair, d)
  const Sched sched = make_sched(cluster_id, gridDim.x >> 1, n_items);

  // pacing (see the producer): leader CTAs of the regular clusters only
  const int n_reg_per = (int)(gridDim.x >> 3);
  const bool pace_on = progress != nullptr && rank == 0 && cluster_id < 4 * n_reg_per && n_reg_per > 1;
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_act);
    ptx::prefetch_tmap(&tmap_wgt);
    *pace_allowed = (pace_delta & 256) ? -1 : (pace_delta & 255);   // staggered followers wait for the first poll
    *pace_done = 0;
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < A_STAGES; ++i) { ptx::mbar_init(&a_full[i], 1); ptx::mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < NSLOT; ++i) { ptx::mbar_init(&s_full[i], 1); ptx::mbar_init(&s_empty[i], 2 * kEpiWarps); }
    ptx::mbar_init(w_full, 1);
    ptx::fence_barrier_init();
  }
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) {
    s_scale[i] = scale[i];
    s_shift[i] = shift[i];
  }
  ptx::cluster_sync_all();
  if (warp == 2) {
    ptx::tmem_alloc<2>(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish<2>();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp != 0 || lane != 0) ptx::pdl_sync();   // the producer lane first puts the resident weights in flight

  if (warp == 0) {
    // ===================================================== TMA producer (one per CTA)
    if (lane == 0) {
      uint32_t a_it = 0, w_loads = 0;
      // Soft pacing of the regular clusters (L2 reuse): an activation plane is read by the four class clusters of an item
      // and by the four of the next item (the clusters with the neighbouring index cj), 8 TMA loads of the same bytes.
      // They only hit L2 if they happen within its retention time (~20 us at this kernel's 3 TB/s of fills), and nothing
      // keeps 72 free-running clusters that close for 12 ms.  So the leader producer publishes the number of sweep steps
      // it has issued, the otherwise idle warp 3 polls the counters of the class mates and of one cluster of each
      // neighbouring group and keeps `pace_allowed` = slowest peer + pace_delta in shared memory, and the producer does
      // not issue step s before pace_allowed >= s.  The slowest cluster never waits (no deadlock among resident CTAs); a
      // wait longer than kPaceTimeout clocks (a peer that has not started yet, CTAs not co-resident) suspends pacing for 64 steps.
      bool pace = pace_on && !(pace_delta & 512);   // bit 9 (diagnostic): publish and poll, never wait
      int pace_step = 0, pace_resume = 0;
      for (int sg = 0; sg < 2; ++sg) {
        const Seg S = sched.s[sg];
        if (S.count == 0) continue;
        const int pd = S.q >> 1;
        if (w_loads > 0) {
          // class change: every MMA that reads the resident weights has retired once the last activation stage of the
          // previous segment was released
          const uint32_t last = a_it - 1;
          ptx::mbar_wait(&a_empty[last % A_STAGES], (last / A_STAGES) & 1);
        }
        if (rank == 0) ptx::mbar_expect_tx(w_full, 2 * W_BYTES);
        const int wrow0 = (S.q * 2 + (int)rank) * (W_BYTES / 128);
#pragma unroll
        for (int j = 0; j < W_BYTES / 32768; ++j)
          ptx::tma_load_2d_2sm(smem_w + j * 32768, &tmap_wgt, w_full, 0, wrow0 + j * 256);
        // the weights are constants: under programmatic dependent launch the first 128 KB per CTA are in flight while
        // the previous layer is still draining; its output is only touched after pdl_sync()
        if (w_loads == 0) ptx::pdl_sync();
        ++w_loads;
        for (int k = 0; k < S.count; ++k) {
          const int t = seg_item(S, k);
          if (t < 0) continue;
          const int d = item_depth(t, pd);
          const int nb = 2 * (t / WIN) + (int)rank;
          for (int j = 0; j < WIN; ++j, ++pace_step) {
            if (pace && pace_step >= pace_resume && *pace_allowed < pace_step) {
              const long long t0 = clock64();
              while (*pace_allowed < pace_step)
                if (clock64() - t0 > kPaceTimeout) { pace_resume = pace_step + 64; break; }   // peer not running: retry later
            }
            for (int sd = 0; sd < 2; ++sd) {
              const int id = d + sd - 1 + pd;
              if (id < 0 || id >= WIN) continue;
              for (int c = 0; c < CHUNKS; ++c, ++a_it) {
                const int as = a_it % A_STAGES;
                ptx::mbar_wait(&a_empty[as], ((a_it / A_STAGES) & 1) ^ 1);
                if (rank == 0) ptx::mbar_expect_tx(&a_full[as], 2 * A_BYTES);
                // (c, n, h, w, d) view: 64 channels x 8 decodes x h' = -1..16 of input column j, depth id
                ptx::tma_load_5d_2sm(smem_a + as * A_BYTES, &tmap_act, &a_full[as], c * 64, nb * NT, -1, j, id);
              }
            }
            if (pace_on) st_relaxed_gpu(progress + cluster_id * kPaceStride, pace_step + 1);
          }
        }
      }
      if (w_loads == 0) ptx::pdl_sync();   // (a cluster without work still takes part in the launch chain)
      // done: nobody may wait for this cluster any more, and its poller can stop
      if (pace_on) {
        st_relaxed_gpu(progress + cluster_id * kPaceStride, 0x7fffffff);
        *pace_done = 1;
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA's warp drives both SMs)
    if (rank == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(256, 256, FMT);
      const uint32_t a_lo0 = ptx::sw128_desc_lo(ptx::smem_u32(smem_a));
      const uint32_t w_lo0 = ptx::sw128_desc_lo(ptx::smem_u32(smem_w));
      uint32_t a_it = 0, w_loads = 0;
      uint32_t fill_par = 0;    // bit s: parity of the number of fills of slot s so far
      for (int sg = 0; sg < 2; ++sg) {
        const Seg S = sched.s[sg];
        if (S.count == 0) continue;
        const int pd = S.q >> 1, ph = S.q & 1;
        ptx::mbar_wait(w_full, w_loads & 1);
        ++w_loads;
        ptx::tc_fence_after();
        for (int k = 0; k < S.count; ++k) {
          const int t = seg_item(S, k);
          if (t < 0) continue;
          const int d = item_depth(t, pd);
          int phi = start_phase(d);
          for (int j = 0; j < WIN; ++j, phi = (phi == 2) ? 0 : phi + 1) {
            const int s0 = phi, s1 = phi + 1;
            if (j == 0 || phi == 0) { ptx::mbar_wait(&s_empty[s0], (fill_par >> s0) & 1); fill_par ^= 1u << s0; }
            ptx::mbar_wait(&s_empty[s1], (fill_par >> s1) & 1);
            fill_par ^= 1u << s1;
            ptx::tc_fence_after();
            const uint32_t dcol = tmem_base + phi * SLOT_COLS;
            for (int sd = 0; sd < 2; ++sd) {
              const int id = d + sd - 1 + pd;
              if (id < 0 || id >= WIN) continue;
#pragma unroll
              for (int c = 0; c < CHUNKS; ++c, ++a_it) {
                const int as = a_it % A_STAGES;
                ptx::mbar_wait(&a_full[as], (a_it / A_STAGES) & 1);
                ptx::tc_fence_after();
                const uint32_t a_lo = a_lo0 + as * (A_BYTES >> 4);
                if (ptx::elect_one()) {
#pragma unroll
                  for (int sh = 0; sh < 2; ++sh) {
                    // delta_h = sh - 1 + ph: rows (h + delta_h + 1) * 8 + n of the tile = a shift of (sh + ph) atoms
                    const uint32_t av = a_lo + (uint32_t)(sh + ph) * (1024 >> 4);
                    const uint32_t wv = w_lo0 + ((sd * 2 + sh) * CHUNKS + c) * (SLICE_BYTES >> 4);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                      ptx::umma_f16<2>(dcol, ptx::sw128_desc(av + kk * 2), ptx::sw128_desc(wv + kk * 2), idesc, 1u);
                  }
                  ptx::umma_commit<2>(&a_empty[as]);
                }
                __syncwarp();
              }
            }
            if (ptx::elect_one()) {
              ptx::umma_commit<2>(&s_full[s0]);
              if (j == WIN - 1) ptx::umma_commit<2>(&s_full[s1]);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 3) {
    // ===================================================== pacing poller (see the producer)
    if (pace_on) {
      const int my_q = cluster_id & 3, my_cj = cluster_id >> 2;
      const int delta = pace_delta & 255;
      const bool stagger = (pace_delta & 256) != 0;
      int peer = -1, peer_lag = 0;
      if (lane < 3) { const int q = (my_q + 1 + lane) & 3; peer = my_cj * 4 + q; peer_lag = (my_cj & 1) + (q != 0); }
      else if (lane == 3) { const int c = (my_cj + n_reg_per - 1) % n_reg_per; peer = c * 4 + my_q; peer_lag = (c & 1) + (my_q != 0); }
      else if (lane == 4) { const int c = (my_cj + 1) % n_reg_per; peer = c * 4 + my_q; peer_lag = (c & 1) + (my_q != 0); }
      const int my_lag = (my_cj & 1) + (my_q != 0);
      // a peer that leads this cluster in the staggered order must be a full step ahead (its loads have landed in L2);
      // every other peer bounds how far this cluster may run ahead
      const int slack = (stagger && peer_lag < my_lag) ? -2 : delta;
      while (*pace_done == 0) {
        int v = 0x7fffffff;
        if (peer >= 0) {
          const int p = ld_relaxed_gpu(progress + peer * kPaceStride);
          v = p > 0x3fffffff ? 0x7fffffff : p + slack;
        }
        v = __reduce_min_sync(0xffffffffu, v);
        if (lane == 0) *pace_allowed = v;
        __nanosleep(1000);   // ~1 M polls/s per counter line: an L2 slice must not become a hot spot
      }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue (each CTA drains its own TMEM): 16 warps,
    // warp = (lane quarter, 32-column group): groups 0,1 = first half of a slot (pw 1 of column j-1, channels 0-31 /
    // 32-63), groups 2,3 = second half (pw 0 of column j)
    const int e = warp - 4;
    const int quarter = e & 3;
    const int cg = e >> 2;
    const int chalf = cg >> 1;
    const int co0 = (cg & 1) * 32;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    constexpr int OD = 2 * WIN;
    uint8_t* stage = smem_o + e * 1024;      // 32 rows x 32 B (16 channels), XOR-swizzled by row pairs
    auto arrive_empty = [&](int slot) {
      if (lane == 0) {
        if (rank == 0) ptx::mbar_arrive(&s_empty[slot]);
        else ptx::mbar_arrive_cluster(&s_empty[slot], 0);
      }
    };
    // all accumulators start at zero: every MMA of the kernel accumulates
    for (int s = 0; s < NSLOT; ++s) {
      tmem_st16_zero(tmem_base + lane_base + s * SLOT_COLS + cg * 32);
      tmem_st16_zero(tmem_base + lane_base + s * SLOT_COLS + cg * 32 + 16);
    }
    tmem_st_wait();
    ptx::tc_fence_before();
    __syncwarp();
    for (int s = 0; s < NSLOT; ++s) arrive_empty(s);

    // drain this warp's 32 columns of `slot` (+ the same columns of slot 3 when add3), re-zero them, and write 32
    // channels of one output voxel column for the warp's 32 (h, decode) rows
    auto drain = [&](int slot, bool add3, bool valid, int nb, int d, int pd, int ph, int ow) {
      const uint32_t tacc = tmem_base + lane_base + slot * SLOT_COLS + cg * 32;
      uint32_t v[32];
      ptx::tmem_ld16(tacc, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
      ptx::tmem_ld16(tacc + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
      if (add3) {
        const uint32_t tacc3 = tmem_base + lane_base + 3 * SLOT_COLS + cg * 32;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t u[16];
          ptx::tmem_ld16(tacc3 + hh * 16, u);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const uint64_t sum = ptx::f2_add(ptx::f2_pack_bits(v[hh * 16 + i], v[hh * 16 + i + 1]),
                                             ptx::f2_pack_bits(u[i], u[i + 1]));
            v[hh * 16 + i] = (uint32_t)sum;
            v[hh * 16 + i + 1] = (uint32_t)(sum >> 32);
          }
        }
        tmem_st16_zero(tacc3);
        tmem_st16_zero(tacc3 + 16);
      } else {
        ptx::tmem_ld_wait();
      }
      tmem_st16_zero(tacc);
      tmem_st16_zero(tacc + 16);
      // the accumulators are in registers and the columns are zero again: hand the slot back to the MMA warp NOW, before
      // the BN / ELU / staging / global stores below.  (Releasing it after the stores, as the first version did, put the
      // latency of the scattered 32-byte-sector stores on the accumulator ring's critical path: with the stores
      // predicated off the layer ran 15 % faster, with the ELU removed not at all -- profiles/r02_notes.md section 11.)
      tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      arrive_empty(slot);
      if (add3) arrive_empty(3);
      if (valid) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {          // two passes of 16 channels through the 1 KB staging tile
          const int co = co0 + hh * 16;
          uint32_t o[8];
          const ulonglong2* sc2 = reinterpret_cast<const ulonglong2*>(s_scale + co);
          const ulonglong2* sh2 = reinterpret_cast<const ulonglong2*>(s_shift + co);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const ulonglong2 sc = sc2[k], sh = sh2[k];
            const uint64_t y0 = ptx::f2_fma(ptx::f2_pack_bits(v[hh * 16 + 4 * k], v[hh * 16 + 4 * k + 1]), sc.x, sh.x);
            const uint64_t y1 = ptx::f2_fma(ptx::f2_pack_bits(v[hh * 16 + 4 * k + 2], v[hh * 16 + 4 * k + 3]), sc.y, sh.y);
            o[2 * k] = act_pack2<FMT, ACT>(y0);
            o[2 * k + 1] = act_pack2<FMT, ACT>(y1);
          }
          // lane = row: 2 x 16 B into a 32-byte staging row (the two chunks swapped on odd row pairs: conflict-free) ...
#pragma unroll
          for (int c2 = 0; c2 < 2; ++c2)
            *reinterpret_cast<uint4*>(stage + lane * 32 + ((c2 ^ ((lane >> 2) & 1)) * 16)) =
                make_uint4(o[4 * c2], o[4 * c2 + 1], o[4 * c2 + 2], o[4 * c2 + 3]);
          __syncwarp();
          // ... then 2 lanes per row: every warp-level store writes 16 complete 32-byte sectors
#pragma unroll
          for (int it = 0; it < 2; ++it) {
            const int r = it * 16 + (lane >> 1);
            const int c16 = lane & 1;
            const uint4 val = *reinterpret_cast<const uint4*>(stage + r * 32 + ((c16 ^ ((r >> 2) & 1)) * 16));
            const int grow = quarter * 32 + r;          // GEMM row = (h, decode)
            const int h = grow / NT, nr = nb * NT + grow % NT;
            if (nr < n_alloc) {
              A3D_DEV_CHECK(nr >= 0 && (unsigned)(2 * d + pd) < (unsigned)OD && (unsigned)(2 * h + ph) < (unsigned)OD &&
                            (unsigned)ow < (unsigned)OD && co + c16 * 8 + 8 <= COUT);
              const size_t vox = (((size_t)nr * OD + (2 * d + pd)) * OD + (2 * h + ph)) * OD + ow;
              __stcs(reinterpret_cast<uint4*>(out + vox * COUT + co + c16 * 8), val);
            }
          }
          __syncwarp();
        }
      }
    };

    uint32_t full_par = 0;      // bit s: parity of the number of "slot closed" signals consumed for slot s
    for (int sg = 0; sg < 2; ++sg) {
      const Seg S = sched.s[sg];
      if (S.count == 0) continue;
      const int pd = S.q >> 1, ph = S.q & 1;
      for (int k = 0; k < S.count; ++k) {
        const int t = seg_item(S, k);
        if (t < 0) continue;
        const int d = item_depth(t, pd);
        const int nb = 2 * (t / WIN) + (int)rank;
        int phi = start_phase(d);
        for (int j = 0; j < WIN; ++j, phi = (phi == 2) ? 0 : phi + 1) {
          ptx::mbar_wait(&s_full[phi], (full_par >> phi) & 1);
          full_par ^= 1u << phi;
          ptx::tc_fence_after();
          // first half: pw 1 of output column j-1 (w_o = 2j-1); second half: pw 0 of column j (w_o = 2j)
          drain(phi, phi == 0 && j > 0, chalf == 1 || j > 0, nb, d, pd, ph, chalf == 0 ? 2 * j - 1 : 2 * j);
          if (j == WIN - 1) {
            // the sweep ends: pw 1 of column 15 sits in the first half of the next slot (its second half is unused)
            ptx::mbar_wait(&s_full[phi + 1], (full_par >> (phi + 1)) & 1);
            full_par ^= 1u << (phi + 1);
            ptx::tc_fence_after();
            drain(phi + 1, false, chalf == 0, nb, d, pd, ph, 2 * WIN - 1);
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync_all();          // no CTA may exit (or free TMEM) while its peer can still signal it
  if (warp == 2) ptx::tmem_dealloc<2>(tmem_base, TMEM_COLS);
}

}  // namespace

constexpr int kProgressInts = 128 * kPaceStride;
size_t convt_l4_sw_progress_bytes() { return kProgressInts * sizeof(int); }
size_t convt_l4_sw_weight_rows() { return (size_t)4 * 2 * (W_BYTES / 128); }

int launch_convt_l4_sw(const CUtensorMap& tmap_act, const CUtensorMap& tmap_wgt, void* out, const float* scale,
                       const float* shift, int64_t n, int64_t n_alloc, int fmt, int act, int num_sms, int* progress,
                       int pace_delta, cudaStream_t st, int64_t* launches) {
  const int n_blocks = (int)((n + NT - 1) / NT);
  const int n_items = ((n_blocks + 1) / 2) * WIN;
  const int n_clusters = l4::num_clusters(num_sms, n_items);
  // pacing needs every cluster resident at once (true for a grid of <= num_sms CTAs on an otherwise idle GPU) and only
  // pays when the launch is long enough for the clusters to drift apart
  if (n_items < 8 * (n_clusters / 4) || 2 * n_clusters > num_sms || (pace_delta & 255) <= 0) progress = nullptr;
  if (progress) A3D_CUDA_OK(cudaMemsetAsync(progress, 0, kProgressInts * sizeof(int), st));
  auto launch = [&](auto kern) -> int {
    A3D_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    A3D_CUDA_OK(launch_chain(kern, dim3(2 * n_clusters), dim3(kThreads), SMEM_BYTES, st, 1, tmap_act, tmap_wgt,
                             reinterpret_cast<uint16_t*>(out), scale, shift, n_blocks, (int)n_alloc, progress, pace_delta));
    return A3D_OK;
  };
  int rc;
  if (fmt == A3D_DTYPE_F16) {
    switch (act) {
      case A3D_ACT_ELU: rc = launch(convt_l4_sw_kernel<A3D_DTYPE_F16, A3D_ACT_ELU>); break;
      case A3D_ACT_RELU: rc = launch(convt_l4_sw_kernel<A3D_DTYPE_F16, A3D_ACT_RELU>); break;
      case A3D_ACT_LRELU: rc = launch(convt_l4_sw_kernel<A3D_DTYPE_F16, A3D_ACT_LRELU>); break;
      default: rc = launch(convt_l4_sw_kernel<A3D_DTYPE_F16, A3D_ACT_NONE>); break;
    }
  } else {
    switch (act) {
      case A3D_ACT_ELU: rc = launch(convt_l4_sw_kernel<A3D_DTYPE_BF16, A3D_ACT_ELU>); break;
      case A3D_ACT_RELU: rc = launch(convt_l4_sw_kernel<A3D_DTYPE_BF16, A3D_ACT_RELU>); break;
      case A3D_ACT_LRELU: rc = launch(convt_l4_sw_kernel<A3D_DTYPE_BF16, A3D_ACT_LRELU>); break;
      default: rc = launch(convt_l4_sw_kernel<A3D_DTYPE_BF16, A3D_ACT_NONE>); break;
    }
  }
  if (rc == A3D_OK && launches) ++*launches;
  return rc;
}

}  // namespace a3d
