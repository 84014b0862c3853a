// 2-CTA (cta_group::2) variant of the stride-1 'same' Conv2D implicit GEMM of conv2d_tc.cu for the deep 256-wide layers of
// Darknet19 (src/net_core/darknet.py:105-131): these layers are bound by L2 -> SM operand traffic, two thirds of which
// is the [256 x 64] weight tile every 128-pixel brick re-fetches per K step.  A cluster of two CTAs runs two adjacent
// bricks against the same N tile as ONE tcgen05.mma.cta_group::2 (M = 256, N = 256): each CTA loads its own activation
// brick and only HALF of the weight tile (its 128 output channels), so the operand bytes per CTA and K step drop from
// 48 KB to 32 KB while every CTA keeps a double-buffered 2 x 256-column accumulator.
//
// Barrier protocol (as convt_tc.cu): `full` lives on the leader (rank 0) and receives one arrive.expect_tx for the bytes
// of BOTH CTAs; the peer's TMA only contributes complete_tx.  The leader's MMA warp issues, tcgen05.commit multicasts the
// `empty` / `t_full` arrivals to both CTAs; all epilogue warps of both CTAs arrive on the leader's `t_empty`.
// Modes: 16-bit NHWC output (staged, coalesced stores) and 16-bit output with the fused 2 x 2 max-pool (shuffles).
#include <cstdlib>

#include "epilogue.cuh"
#include "internal.h"
#include "ptx.cuh"

namespace a3d {
namespace {

constexpr int BM = 128, BN = 256, KC = 64;
constexpr int kEpiWarps = 16;
constexpr int kThreads = 128 + 32 * kEpiWarps;
constexpr int A_BYTES = BM * KC * 2;          // 16 KB: this CTA's brick
constexpr int BH_BYTES = (BN / 2) * KC * 2;   // 16 KB: this CTA's half of the weight tile
constexpr int STAGES = 6;
constexpr int STAGING_BYTES = kEpiWarps * 32 * 64;
constexpr int NUM_BARS = 2 * STAGES + 4;
constexpr int SMEM_BYTES = 1024 + STAGES * (A_BYTES + BH_BYTES) + STAGING_BYTES + NUM_BARS * 8 + 16;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ void tma_load_4d_2sm(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2,
                                                int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(ptx::smem_u32(bar) & ptx::kPeerBitMask),
      "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

template <int ACT>
__device__ __forceinline__ float act2d(float v) {
  if constexpr (ACT == A3D_ACT_LRELU01) return v > 0.f ? v : 0.1f * v;
  else return activate<ACT>(v);
}

// MODE 0: 16-bit [pixels, cout_pad]; 2: 16-bit with the 2 x 2 max-pool fused, [n, H/2, W/2, cout_pad]
template <int FMT, int ACT, int MODE>
__global__ void __launch_bounds__(kThreads, 1)
conv2d_pair_kernel(const __grid_constant__ CUtensorMap tmap_act, const __grid_constant__ CUtensorMap tmap_wgt,
                   uint16_t* __restrict__ out, const float* __restrict__ scale, const float* __restrict__ shift,
                   Conv2dGeom g) {
  constexpr int CW = BN / 4;   // columns per epilogue warp
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_BYTES;
  uint8_t* smem_stage = smem_b + STAGES * BH_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stage + STAGING_BYTES);
  uint64_t* full = bars;              // used on the leader
  uint64_t* empty = full + STAGES;    // per CTA (multicast commit)
  uint64_t* t_full = empty + STAGES;  // per CTA (multicast commit)
  uint64_t* t_empty = t_full + 2;     // used on the leader
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = (int)ptx::cluster_ctarank();
  const int cl_id = blockIdx.x >> 1, n_cl = gridDim.x >> 1;
  const int total_units = (g.m_tiles >> 1) * g.n_tiles;   // unit = (pair of bricks, N tile)
  const int ksteps = g.taps * g.cin_chunks;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_act);
    ptx::prefetch_tmap(&tmap_wgt);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&t_full[i], 1); ptx::mbar_init(&t_empty[i], 2 * kEpiWarps); }
    ptx::fence_barrier_init();
  }
  ptx::cluster_sync_all();   // barrier inits visible before any remote arrive / multicast
  if (warp == 2) {
    ptx::tmem_alloc<2>(tmem_slot, 512);
    ptx::tmem_relinquish<2>();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_sync();   // programmatic dependent launch: the prologue above overlaps the previous layer (ptx.cuh)

  if (warp == 0 || warp == 3) {
    // ===================================================== TMA producers (even / odd K steps), one pair per CTA
    const int par = warp == 3 ? 1 : 0;
    int s = par;
    uint32_t ph = 0, cnt = 0;
    for (int u = cl_id; u < total_units; u += n_cl) {
      const int nt = u % g.n_tiles, mt = (u / g.n_tiles) * 2 + rank;
      const int tw = mt % g.tiles_w, th = (mt / g.tiles_w) % g.tiles_h, nb = mt / (g.tiles_w * g.tiles_h);
      const int w0 = tw << g.lw, h0 = th << g.lh, n0 = nb << (7 - g.lw - g.lh);
      int dy = g.taps == 9 ? -1 : 0, dx = dy;
      int brow = nt * BN + rank * (BN / 2);     // this CTA's N half of the weight tile
      for (int tap = 0; tap < g.taps; ++tap) {
        for (int kc = 0; kc < g.cin_chunks; ++kc, ++cnt) {
          if ((cnt & 1u) != (uint32_t)par) continue;
          ptx::mbar_wait(&empty[s], ph ^ 1);
          if (ptx::elect_one()) {
            if (rank == 0) ptx::mbar_expect_tx(&full[s], 2 * (A_BYTES + BH_BYTES));
            tma_load_4d_2sm(smem_a + s * A_BYTES, &tmap_act, &full[s], kc * KC, w0 + dx, h0 + dy, n0);
            ptx::tma_load_2d_2sm(smem_b + s * BH_BYTES, &tmap_wgt, &full[s], kc * KC, brow);
          }
          __syncwarp();
          s += 2;
          if (s >= STAGES) { s -= STAGES; ph ^= 1; }
        }
        brow += g.cout_pad;
        if (++dx > 1) { dx = -1; ++dy; }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA only)
    if (rank == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(2 * BM, BN, FMT);
      const uint32_t a_lo0 = ptx::sw128_desc_lo(ptx::smem_u32(smem_a));
      const uint32_t b_lo0 = ptx::sw128_desc_lo(ptx::smem_u32(smem_b));
      uint32_t unit_it = 0, ph = 0;
      int s = 0;
      for (int u = cl_id; u < total_units; u += n_cl, ++unit_it) {
        const int buf = unit_it & 1;
        ptx::mbar_wait(&t_empty[buf], ((unit_it >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t tacc = tmem_base + buf * BN;
        for (int ks = 0; ks < ksteps; ++ks) {
          ptx::mbar_wait(&full[s], ph);
          ptx::tc_fence_after();
          const uint32_t a_lo = a_lo0 + s * (A_BYTES >> 4), b_lo = b_lo0 + s * (BH_BYTES >> 4);
          if (ptx::elect_one()) {
#pragma unroll
            for (int kk = 0; kk < KC / 16; ++kk)
              ptx::umma_f16<2>(tacc, ptx::sw128_desc(a_lo + kk * 2), ptx::sw128_desc(b_lo + kk * 2), idesc, (ks | kk) != 0);
            ptx::umma_commit<2>(&empty[s]);
            if (ks == ksteps - 1) ptx::umma_commit<2>(&t_full[buf]);
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue (both CTAs, each on its own brick)
    const int e = warp - 4;
    const int quarter = e & 3, chalf = e >> 2;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    constexpr int GROUPS = CW / 32;
    uint32_t unit_it = 0;
    for (int u = cl_id; u < total_units; u += n_cl, ++unit_it) {
      const int nt = u % g.n_tiles, mt = (u / g.n_tiles) * 2 + rank;
      const int buf = unit_it & 1;
      ptx::mbar_wait(&t_full[buf], (unit_it >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t tacc = tmem_base + lane_base + buf * BN + chalf * CW;
      const int tw = mt % g.tiles_w, th = (mt / g.tiles_w) % g.tiles_h, nb = mt / (g.tiles_w * g.tiles_h);
      const int r = quarter * 32 + lane;
      const int wi = r & ((1 << g.lw) - 1), hi = (r >> g.lw) & ((1 << g.lh) - 1), ni = r >> (g.lw + g.lh);
      const int img = (nb << (7 - g.lw - g.lh)) + ni;
      const int ph = (th << g.lh) + hi, pw = (tw << g.lw) + wi;
      const bool row_ok = img < g.n_images && pw < g.W && ph < g.H;
      const int64_t p = ((int64_t)img * (g.H >> 1) + (ph >> 1)) * (g.W >> 1) + (pw >> 1);   // pooled index (MODE 2)
      int64_t prow[4];
      if constexpr (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int rr = quarter * 32 + 8 * j + (lane >> 2);
          const int rwi = rr & ((1 << g.lw) - 1), rhi = (rr >> g.lw) & ((1 << g.lh) - 1), rni = rr >> (g.lw + g.lh);
          const int rimg = (nb << (7 - g.lw - g.lh)) + rni;
          const int rph = (th << g.lh) + rhi, rpw = (tw << g.lw) + rwi;
          prow[j] = (rimg < g.n_images && rph < g.H && rpw < g.W) ? ((int64_t)rimg * g.H + rph) * g.W + rpw : -1;
        }
      }
#pragma unroll 1
      for (int gi = 0; gi < GROUPS; ++gi) {
        const int co0 = nt * BN + chalf * CW + gi * 32;
        uint32_t v[32];
        ptx::tmem_ld16(tacc + gi * 32, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
        ptx::tmem_ld16(tacc + gi * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
        ptx::tmem_ld_wait();
        const float4* sc4 = reinterpret_cast<const float4*>(scale + co0);
        const float4* sh4 = reinterpret_cast<const float4*>(shift + co0);
        if constexpr (MODE == 2) {
          const int wt = 1 << g.lw;
          float m[32];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 sc = __ldg(sc4 + i);
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
          const float4 sha = __ldg(sh4 + sub * 2), shb = __ldg(sh4 + sub * 2 + 1);
          const float sh8[8] = {sha.x, sha.y, sha.z, sha.w, shb.x, shb.y, shb.z, shb.w};
          float y[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float t = sub == 0 ? m[j] : sub == 1 ? m[8 + j] : sub == 2 ? m[16 + j] : m[24 + j];
            y[j] = act2d<ACT>(t + sh8[j]);
          }
          if (row_ok)
            *reinterpret_cast<uint4*>(out + p * g.cout_pad + co0 + sub * 8) =
                make_uint4(pack2<FMT>(y[0], y[1]), pack2<FMT>(y[2], y[3]), pack2<FMT>(y[4], y[5]), pack2<FMT>(y[6], y[7]));
        } else {
          uint32_t o[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 sc = __ldg(sc4 + i), sh = __ldg(sh4 + i);
            const float x0 = act2d<ACT>(fmaf(__uint_as_float(v[4 * i]), sc.x, sh.x));
            const float x1 = act2d<ACT>(fmaf(__uint_as_float(v[4 * i + 1]), sc.y, sh.y));
            const float x2 = act2d<ACT>(fmaf(__uint_as_float(v[4 * i + 2]), sc.z, sh.z));
            const float x3 = act2d<ACT>(fmaf(__uint_as_float(v[4 * i + 3]), sc.w, sh.w));
            o[2 * i] = pack2<FMT>(x0, x1);
            o[2 * i + 1] = pack2<FMT>(x2, x3);
          }
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
            if (prow[j] >= 0) *reinterpret_cast<uint4*>(out + prow[j] * g.cout_pad + co0 + ch * 8) = q;
          }
          __syncwarp();
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) ptx::mbar_arrive(&t_empty[buf]);
        else ptx::mbar_arrive_cluster(&t_empty[buf], 0);
      }
    }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  if (warp == 2) ptx::tmem_dealloc<2>(tmem_base, 512);
}

template <int FMT, int MODE>
int launch_act(const CUtensorMap& ta, const CUtensorMap& tw, void* out, const float* scale, const float* shift,
               const Conv2dGeom& g, int act, int n_cl, cudaStream_t st) {
  auto launch = [&](auto kern) -> int {
    A3D_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    A3D_CUDA_OK(launch_chain(kern, dim3(n_cl * 2), dim3(kThreads), SMEM_BYTES, st, 2, ta, tw,
                             reinterpret_cast<uint16_t*>(out), scale, shift, g));
    return A3D_OK;
  };
  switch (act) {
    case A3D_ACT_ELU: return launch(conv2d_pair_kernel<FMT, A3D_ACT_ELU, MODE>);
    case A3D_ACT_RELU: return launch(conv2d_pair_kernel<FMT, A3D_ACT_RELU, MODE>);
    case A3D_ACT_LRELU: return launch(conv2d_pair_kernel<FMT, A3D_ACT_LRELU, MODE>);
    case A3D_ACT_LRELU01: return launch(conv2d_pair_kernel<FMT, A3D_ACT_LRELU01, MODE>);
    case A3D_ACT_NONE: return launch(conv2d_pair_kernel<FMT, A3D_ACT_NONE, MODE>);
    default: set_error("conv2d_pair: unsupported activation %d", act); return A3D_ERR_INVALID;
  }
}

}  // namespace

bool conv2d_pair_eligible(const Conv2dGeom& g, int bn, bool pool, bool out_f32) {
  static const bool off = [] { const char* e = getenv("A3D_ENC_PAIR"); return e && e[0] == '0'; }();
  return !off && !g.rh && g.kc == 64 && bn == 256 && !out_f32 && g.m_tiles >= 2 && (g.m_tiles % 2) == 0 &&
         (!pool || (g.lw <= 4 && g.lh >= 1));
}

// tmap_wgt_half: the layer's weight tensor with a (64, 128)-row box (one N half per CTA)
int launch_conv2d_pair(const CUtensorMap& tmap_act, const CUtensorMap& tmap_wgt_half, void* out, const float* scale,
                       const float* shift, const Conv2dGeom& g, int fmt, int act, bool pool, int num_sms, cudaStream_t st,
                       int64_t* launches) {
  if (g.n_images <= 0) return A3D_OK;
  const int total = (g.m_tiles / 2) * g.n_tiles;
  int n_cl = num_sms / 2;
  if (n_cl > total) n_cl = total;
  int rc;
  if (fmt == A3D_DTYPE_F16)
    rc = pool ? launch_act<A3D_DTYPE_F16, 2>(tmap_act, tmap_wgt_half, out, scale, shift, g, act, n_cl, st)
              : launch_act<A3D_DTYPE_F16, 0>(tmap_act, tmap_wgt_half, out, scale, shift, g, act, n_cl, st);
  else
    rc = pool ? launch_act<A3D_DTYPE_BF16, 2>(tmap_act, tmap_wgt_half, out, scale, shift, g, act, n_cl, st)
              : launch_act<A3D_DTYPE_BF16, 0>(tmap_act, tmap_wgt_half, out, scale, shift, g, act, n_cl, st);
  if (rc == A3D_OK && launches) ++*launches;
  return rc;
}

}  // namespace a3d
