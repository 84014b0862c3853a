// 128 -> 64 stride-2 transposed conv (58.6 % of the decoder FLOPs) as a WEIGHT-STATIONARY 2-CTA tcgen05 kernel.
// Same implicit-GEMM formulation as convt_tc.cu (conv3DDec, autoencoder3D.py:41-54), different data movement:
//
//  * CTA pairs (cluster of 2, tcgen05 cta_group::2, MMA M = 256): the two CTAs run the SAME (d, h, parity class) on two
//    different 8-decode blocks, so they share every B (weight) operand; each CTA supplies its N-half of B.
//  * All weights of one (pd, ph) parity class for one CTA (8 (sd,sh,chunk) slices x 128 rows x 128 B = 128 KB) stay
//    RESIDENT in shared memory for the whole launch; only the activation rows stream through a 5-stage TMA ring
//    (18 KB per stage), and every loaded row feeds TWO consecutive units of an h-sweep (24 MMAs per stage).
//    L2->SM operand traffic drops from ~400 KB to ~72 KB per unit.
//  * One MMA-issuing thread now drives two SMs, halving the per-SM instruction-issue cost of the N = 64 MMAs.
//
// Per (sd, sh, chunk) slice the CTA of cluster rank r holds 128 rows:
//   [ 0, 64)  N-half r of the delta_w = 0 MMA (N = 128):  rank 0 = (pw0, tap_w 1) co 0..63, rank 1 = (pw1, tap_w 2)
//   [64, 96)  N-half r of the delta_w = -1 MMA (N = 64, pw0, tap_w 3): co 32r .. 32r+31
//   [96,128)  N-half r of the delta_w = +1 MMA (N = 64, pw1, tap_w 0): co 32r .. 32r+31
#include <cstdlib>

#include "epilogue.cuh"
#include "internal.h"
#include "ptx.cuh"

namespace a3d {
namespace {

constexpr int COUT = 64, WIN = 16, NT = 8, NACC = 128, CHUNKS = 2;
constexpr int A_BYTES = (WIN + 2) * NT * 128;   // 18432
constexpr int A_STAGES = 4;
constexpr int OUT_STAGE_BYTES = 8 * 2048;        // per epilogue warp: 32 rows x 64 B (32 channels), XOR-swizzled
constexpr int SLICE_BYTES = 128 * 128;          // one (sd, sh, chunk) slice of one rank
constexpr int W_BYTES = 8 * SLICE_BYTES;        // 131072
constexpr int NBUF = 4;                         // two units accumulate while older ones drain
constexpr int TMEM_COLS = NBUF * NACC;          // 512
constexpr int kEpiWarps = 8;
constexpr int kThreads = 128 + 32 * kEpiWarps;
constexpr int NUM_BARS = 2 * A_STAGES + 2 * NBUF + 1;
constexpr int SMEM_BYTES = 1024 + W_BYTES + A_STAGES * A_BYTES + OUT_STAGE_BYTES + NUM_BARS * 8 + 16 + 2 * NACC * 4;

// Schedule.  Cluster c owns parity class q = c % 4 for the whole launch (weights loaded once) and, together with the
// other clusters of its class, walks the items t = (decode-block pair, d).  An item is a sweep over h = 0..15: the
// activation row (d + delta_d, r) is loaded ONCE and feeds two units, h = r - ph (its delta_h = ph tap) and
// h = r - ph + 1 (its delta_h = ph - 1 tap), whose accumulators live in different TMEM buffers.  The four clusters of
// one position group run in lock-step on the four classes, so all but the first touch of a row hit L2.
template <int FMT, int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
convt_l4_ws_kernel(const __grid_constant__ CUtensorMap tmap_act, const __grid_constant__ CUtensorMap tmap_wgt,
                   uint16_t* __restrict__ out, const float* __restrict__ scale, const float* __restrict__ shift,
                   int n_blocks, int n_alloc) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128-byte swizzle, computed on the shared-window address so the pointer keeps its
  // __shared__ provenance (LDS/STS instead of generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem + W_BYTES;
  uint8_t* smem_o = smem_a + A_STAGES * A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_o + OUT_STAGE_BYTES);
  uint64_t* a_full = bars;                    // [A_STAGES]  used on the leader CTA (1 arrival + both CTAs' bytes)
  uint64_t* a_empty = a_full + A_STAGES;      // [A_STAGES]  per CTA, multicast commit
  uint64_t* t_full = a_empty + A_STAGES;      // [NBUF]      per CTA, multicast commit
  uint64_t* t_empty = t_full + NBUF;          // [NBUF]      used on the leader CTA (2 x 8 epilogue warps)
  uint64_t* w_full = t_empty + NBUF;          // leader: weights of both CTAs landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
  float* s_scale = reinterpret_cast<float*>(tmem_slot + 2);
  float* s_shift = s_scale + NACC;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int n_per_class = (gridDim.x >> 1) >> 2;      // clusters per parity class
  const int q = cluster_id & 3;                       // parity class of this cluster
  const int cj = cluster_id >> 2;                     // index of the cluster within its class
  const int pd = q >> 1, ph = q & 1;
  const int n_items = ((n_blocks + 1) >> 1) * WIN;    // (decode-block pair, d)

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_act);
    ptx::prefetch_tmap(&tmap_wgt);
  }
  if (warp == 1 && lane == 0) {
    // a_full / w_full live on the leader: ONE arrival (its own arrive.expect_tx for the bytes of BOTH CTAs); the peer's
    // TMA only contributes complete_tx bytes (a transiently negative tx-count is legal), so no remote arrive is needed
    for (int i = 0; i < A_STAGES; ++i) { ptx::mbar_init(&a_full[i], 1); ptx::mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < NBUF; ++i) { ptx::mbar_init(&t_full[i], 1); ptx::mbar_init(&t_empty[i], 2 * kEpiWarps); }
    ptx::mbar_init(w_full, 1);
    ptx::fence_barrier_init();
  }
  for (int i = threadIdx.x; i < NACC; i += blockDim.x) {
    s_scale[i] = scale[i % COUT];
    s_shift[i] = shift[i % COUT];
  }
  ptx::cluster_sync_all();          // barrier inits visible cluster-wide before any remote arrive / multicast
  if (warp == 2) {
    ptx::tmem_alloc<2>(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish<2>();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================== TMA producer (one per CTA)
    if (lane == 0) {
      // resident weights of this cluster's class, N-half of this rank
      if (rank == 0) ptx::mbar_expect_tx(w_full, 2 * W_BYTES);
      const int wrow0 = (q * 2 + (int)rank) * (W_BYTES / 128);
#pragma unroll
      for (int j = 0; j < W_BYTES / 32768; ++j)
        ptx::tma_load_2d_2sm(smem_w + j * 32768, &tmap_wgt, w_full, 0, wrow0 + j * 256);
      uint32_t a_it = 0;
      for (int t = cj; t < n_items; t += n_per_class) {
        const int d = t % WIN;
        const int nb = 2 * (t / WIN) + (int)rank;
        for (int r = 0; r < WIN; ++r) {
          for (int sd = 0; sd < 2; ++sd) {
            const int id = d + sd - 1 + pd;
            if (id < 0 || id >= WIN) continue;
            for (int c = 0; c < CHUNKS; ++c, ++a_it) {
              const int as = a_it % A_STAGES;
              ptx::mbar_wait(&a_empty[as], ((a_it / A_STAGES) & 1) ^ 1);
              if (rank == 0) ptx::mbar_expect_tx(&a_full[as], 2 * A_BYTES);
              ptx::tma_load_5d_2sm(smem_a + as * A_BYTES, &tmap_act, &a_full[as], c * 64, nb * NT, -1, r, id);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer: the leader CTA's warp drives both SMs; the warp
    // stays converged, only tcgen05.mma / commit are predicated on one elected lane
    if (rank == 0) {
      constexpr uint32_t idesc_full = ptx::make_idesc_f16(256, 128, FMT);
      constexpr uint32_t idesc_half = ptx::make_idesc_f16(256, 64, FMT);
      const uint32_t a_lo0 = ptx::sw128_desc_lo(ptx::smem_u32(smem_a));
      const uint32_t w_lo0 = ptx::sw128_desc_lo(ptx::smem_u32(smem_w));
      constexpr uint32_t W1 = (NT * 128) >> 4;
      ptx::mbar_wait(w_full, 0);
      ptx::tc_fence_after();
      uint32_t a_it = 0, u0 = 0;     // u0: unit counter at the start of the item (unit h of the item is u0 + h)
      for (int t = cj; t < n_items; t += n_per_class, u0 += WIN) {
        const int d = t % WIN;
        for (int r = 0; r < WIN; ++r) {
          const int hA = r - ph;          // unit finishing on this row (its delta_h = ph tap, slice sh = 1)
          const int hB = r - ph + 1;      // unit starting on this row  (its delta_h = ph - 1 tap, slice sh = 0)
          const bool vA = hA >= 0, vB = hB < WIN;
          const uint32_t uA = u0 + hA, uB = u0 + hB;
          const uint32_t tA = tmem_base + (uA % NBUF) * NACC, tB = tmem_base + (uB % NBUF) * NACC;
          // a unit's accumulator buffer must have been drained before its first MMA
          if (vB) ptx::mbar_wait(&t_empty[uB % NBUF], ((uB / NBUF) & 1) ^ 1);
          if (vA && hA == 0 && ph == 0) ptx::mbar_wait(&t_empty[uA % NBUF], ((uA / NBUF) & 1) ^ 1);  // h = 0 has no earlier row
          ptx::tc_fence_after();
          uint32_t accA = (vA && !(hA == 0 && ph == 0)) ? 1u : 0u;   // unit A already holds its sh = 0 contribution
          uint32_t accB = 0;
          for (int sd = 0; sd < 2; ++sd) {
            const int id = d + sd - 1 + pd;
            if (id < 0 || id >= WIN) continue;
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c, ++a_it) {
              const int as = a_it % A_STAGES;
              ptx::mbar_wait(&a_full[as], (a_it / A_STAGES) & 1);
              ptx::tc_fence_after();
              const uint32_t a_lo = a_lo0 + as * (A_BYTES >> 4);
              const uint32_t wA = w_lo0 + ((sd * 2 + 1) * CHUNKS + c) * (SLICE_BYTES >> 4);
              const uint32_t wB = w_lo0 + ((sd * 2 + 0) * CHUNKS + c) * (SLICE_BYTES >> 4);
              if (ptx::elect_one()) {
                if (vA) {
#pragma unroll
                  for (int kk = 0; kk < 4; ++kk) {
                    const uint32_t ko = kk * 2;
                    ptx::umma_f16<2>(tA, ptx::sw128_desc(a_lo + W1 + ko), ptx::sw128_desc(wA + ko), idesc_full,
                                     (kk == 0) ? accA : 1u);
                    ptx::umma_f16<2>(tA, ptx::sw128_desc(a_lo + ko), ptx::sw128_desc(wA + ((64 * 128) >> 4) + ko),
                                     idesc_half, 1);
                    ptx::umma_f16<2>(tA + COUT, ptx::sw128_desc(a_lo + 2 * W1 + ko),
                                     ptx::sw128_desc(wA + ((96 * 128) >> 4) + ko), idesc_half, 1);
                  }
                }
                if (vB) {
#pragma unroll
                  for (int kk = 0; kk < 4; ++kk) {
                    const uint32_t ko = kk * 2;
                    ptx::umma_f16<2>(tB, ptx::sw128_desc(a_lo + W1 + ko), ptx::sw128_desc(wB + ko), idesc_full,
                                     (kk == 0) ? accB : 1u);
                    ptx::umma_f16<2>(tB, ptx::sw128_desc(a_lo + ko), ptx::sw128_desc(wB + ((64 * 128) >> 4) + ko),
                                     idesc_half, 1);
                    ptx::umma_f16<2>(tB + COUT, ptx::sw128_desc(a_lo + 2 * W1 + ko),
                                     ptx::sw128_desc(wB + ((96 * 128) >> 4) + ko), idesc_half, 1);
                  }
                }
                ptx::umma_commit<2>(&a_empty[as]);   // both CTAs' stage `as` reusable once these MMAs retire
              }
              __syncwarp();
              accA = 1;
              accB = 1;
            }
          }
          if (vA) {
            if (ptx::elect_one()) ptx::umma_commit<2>(&t_full[uA % NBUF]);   // unit A complete in both CTAs
            __syncwarp();
          }
          if (vB && r == WIN - 1) {   // ph = 1: the last unit has no row 16, it completes here
            if (ptx::elect_one()) ptx::umma_commit<2>(&t_full[uB % NBUF]);
            __syncwarp();
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue (each CTA drains its own TMEM)
    const int e = warp - 4;
    const int quarter = e & 3;
    const int chalf = e >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    constexpr int OD = 2 * WIN;
    constexpr int NCOLS = NACC / 2;   // 64 columns = one pw parity per warp
    const int pw = chalf;
    uint8_t* stage = smem_o + e * 2048;
    (void)row;
    uint32_t u = 0;
    for (int t = cj; t < n_items; t += n_per_class) {
      const int d = t % WIN;
      const int nb = 2 * (t / WIN) + (int)rank;
      for (int h = 0; h < WIN; ++h, ++u) {
        const int buf = u % NBUF;
        ptx::mbar_wait(&t_full[buf], (u / NBUF) & 1);
        ptx::tc_fence_after();
        const uint32_t tacc = tmem_base + lane_base + buf * NACC + chalf * NCOLS;
#pragma unroll 1
        for (int g = 0; g < NCOLS / 32; ++g) {
          uint32_t v[32];
          ptx::tmem_ld16(tacc + g * 32, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
          ptx::tmem_ld16(tacc + g * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
          ptx::tmem_ld_wait();
          const int co = g * 32;
          uint32_t o[16];
          const float4* sc4 = reinterpret_cast<const float4*>(s_scale + co);
          const float4* sh4 = reinterpret_cast<const float4*>(s_shift + co);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float4 sc = sc4[k], sh = sh4[k];
            const float x0 = activate<ACT>(fmaf(__uint_as_float(v[4 * k]), sc.x, sh.x));
            const float x1 = activate<ACT>(fmaf(__uint_as_float(v[4 * k + 1]), sc.y, sh.y));
            const float x2 = activate<ACT>(fmaf(__uint_as_float(v[4 * k + 2]), sc.z, sh.z));
            const float x3 = activate<ACT>(fmaf(__uint_as_float(v[4 * k + 3]), sc.w, sh.w));
            o[2 * k] = pack2<FMT>(x0, x1);
            o[2 * k + 1] = pack2<FMT>(x2, x3);
          }
          // lane = row: 4 x 16 B into a 64-byte staging row (chunks swizzled by row pairs, conflict-free) ...
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4)
            *reinterpret_cast<uint4*>(stage + lane * 64 + ((c4 ^ ((lane >> 1) & 3)) * 16)) =
                make_uint4(o[4 * c4], o[4 * c4 + 1], o[4 * c4 + 2], o[4 * c4 + 3]);
          __syncwarp();
          // ... then 4 lanes per row: every warp-level store writes 8 contiguous 64-byte half-lines
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int r = it * 8 + (lane >> 2);
            const int c16 = lane & 3;
            const uint4 val = *reinterpret_cast<const uint4*>(stage + r * 64 + ((c16 ^ ((r >> 1) & 3)) * 16));
            const int grow = quarter * 32 + r;
            const int wr = grow / NT, nr = nb * NT + grow % NT;
            if (nr < n_alloc) {
              const size_t vox = (((size_t)nr * OD + (2 * d + pd)) * OD + (2 * h + ph)) * OD + (2 * wr + pw);
              __stcs(reinterpret_cast<uint4*>(out + vox * COUT + co + c16 * 8), val);   // streaming: keep L2 for the inputs
            }
          }
          __syncwarp();
        }
        ptx::tc_fence_before();
        __syncwarp();                      // every lane's tcgen05.ld has completed: one arrive per warp
        if (lane == 0) {
          if (rank == 0) ptx::mbar_arrive(&t_empty[buf]);
          else ptx::mbar_arrive_cluster(&t_empty[buf], 0);
        }
      }
    }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync_all();          // no CTA may exit (or free TMEM) while its peer can still signal it
  if (warp == 2) ptx::tmem_dealloc<2>(tmem_base, TMEM_COLS);
}

}  // namespace

size_t convt_l4_ws_weight_rows() { return (size_t)4 * 2 * (W_BYTES / 128); }

int launch_convt_l4_ws(const CUtensorMap& tmap_act, const CUtensorMap& tmap_wgt, void* out, const float* scale,
                       const float* shift, int64_t n, int64_t n_alloc, int fmt, int act, int num_sms,
                       cudaStream_t st, int64_t* launches) {
  const int n_blocks = (int)((n + NT - 1) / NT);
  const int n_items = ((n_blocks + 1) / 2) * WIN;
  int n_clusters = (num_sms / 2) & ~3;           // a multiple of 4: one cluster per parity class in lock-step
  if (n_clusters > 4 * n_items) n_clusters = 4 * n_items;
  if (n_clusters < 4) n_clusters = 4;
  auto launch = [&](auto kern) -> int {
    A3D_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    kern<<<2 * n_clusters, kThreads, SMEM_BYTES, st>>>(tmap_act, tmap_wgt, reinterpret_cast<uint16_t*>(out), scale,
                                                       shift, n_blocks, (int)n_alloc);
    A3D_CUDA_OK(cudaGetLastError());
    return A3D_OK;
  };
  int rc;
  if (fmt == A3D_DTYPE_F16) {
    switch (act) {
      case A3D_ACT_ELU: rc = launch(convt_l4_ws_kernel<A3D_DTYPE_F16, A3D_ACT_ELU>); break;
      case A3D_ACT_RELU: rc = launch(convt_l4_ws_kernel<A3D_DTYPE_F16, A3D_ACT_RELU>); break;
      case A3D_ACT_LRELU: rc = launch(convt_l4_ws_kernel<A3D_DTYPE_F16, A3D_ACT_LRELU>); break;
      default: rc = launch(convt_l4_ws_kernel<A3D_DTYPE_F16, A3D_ACT_NONE>); break;
    }
  } else {
    switch (act) {
      case A3D_ACT_ELU: rc = launch(convt_l4_ws_kernel<A3D_DTYPE_BF16, A3D_ACT_ELU>); break;
      case A3D_ACT_RELU: rc = launch(convt_l4_ws_kernel<A3D_DTYPE_BF16, A3D_ACT_RELU>); break;
      case A3D_ACT_LRELU: rc = launch(convt_l4_ws_kernel<A3D_DTYPE_BF16, A3D_ACT_LRELU>); break;
      default: rc = launch(convt_l4_ws_kernel<A3D_DTYPE_BF16, A3D_ACT_NONE>); break;
    }
  }
  if (rc == A3D_OK && launches) ++*launches;
  return rc;
}

}  // namespace a3d
