// Stride-1 transposed conv (4^3 x 8 -> 4^3 x 512; conv3DDec with strides = 1, autoencoder3D.py:41-54,127-128) + folded
// BN + activation as ONE dense tcgen05 GEMM:  a1[n, (o, co)] = act(scale[co] * sum_k a0[n, k] * Mt[(o, co), k] + shift[co])
// with k = (input voxel i, ci) and Mt[(o, co), (i, ci)] = W[t = o - i + 1, co, ci] (zero where the tap falls outside
// the 4-tap kernel; 42 % of the entries are non-zero).  The input of a decode is only 512 values, so the layer is a
// plain [n, 512] x [512, 32768] GEMM; issuing the zero taps costs 33.5 MFLOP per decode (0.5 % of the decoder) and
// replaces a CUDA-core kernel that took 6.5 % of the step.
//
// Unit = 128 decodes x 256 output columns; K = 512 in 8 chunks of 64 through a 3-stage TMA ring (A 16 KB + B 32 KB per
// stage); fp32 accumulators double-buffered in TMEM (2 x 256 columns); same warp roles as convt_tc.cu.
#include <cstdlib>

#include "epilogue.cuh"
#include "internal.h"
#include "ptx.cuh"

namespace a3d {
namespace {

constexpr int KDIM = 512, NDIM = 64 * 512, COUT = 512;
constexpr int BM = 128, BN = 256, KCHUNKS = KDIM / 64;
constexpr int A_BYTES = BM * 128, B_BYTES = BN * 128;
constexpr int STAGES = 3;
constexpr int OUT_STAGE_BYTES = 8 * 4096;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 128 + 32 * kEpiWarps;
constexpr int NUM_BARS = 2 * STAGES + 4;
constexpr int SMEM_BYTES = 1024 + STAGES * (A_BYTES + B_BYTES) + OUT_STAGE_BYTES + NUM_BARS * 8 + 16 + 2 * COUT * 4;
constexpr int N_TILES = NDIM / BN;   // 128

template <int FMT, int ACT>
__global__ void __launch_bounds__(kThreads, 1)
gemm_l1_kernel(const __grid_constant__ CUtensorMap tmap_a0, const __grid_constant__ CUtensorMap tmap_mt,
               uint16_t* __restrict__ out, const float* __restrict__ scale, const float* __restrict__ shift,
               int m_tiles, int n_rows) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_BYTES;
  uint8_t* smem_o = smem_b + STAGES * B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_o + OUT_STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = full + STAGES;
  uint64_t* t_full = empty + STAGES;
  uint64_t* t_empty = t_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
  float* s_scale = reinterpret_cast<float*>(tmem_slot + 4);
  float* s_shift = s_scale + COUT;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_units = m_tiles * N_TILES;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a0);
    ptx::prefetch_tmap(&tmap_mt);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&t_full[i], 1); ptx::mbar_init(&t_empty[i], 32 * kEpiWarps); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc<1>(tmem_slot, 512);
    ptx::tmem_relinquish<1>();
  }
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) {
    s_scale[i] = scale[i];
    s_shift[i] = shift[i];
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_sync();     // a0 is the previous kernel's output

  if (warp == 0) {
    if (lane == 0) {   // ===================================================== TMA producer
      uint32_t it = 0;
      for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
        const int nt = u % N_TILES, mt = u / N_TILES;
        for (int kc = 0; kc < KCHUNKS; ++kc, ++it) {
          const int s = it % STAGES;
          ptx::mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
          ptx::mbar_expect_tx(&full[s], A_BYTES + B_BYTES);
          ptx::tma_load_2d(smem_a + s * A_BYTES, &tmap_a0, &full[s], kc * 64, mt * BM);
          ptx::tma_load_2d(smem_b + s * B_BYTES, &tmap_mt, &full[s], kc * 64, nt * BN);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (converged warp, elected-lane issue)
    constexpr uint32_t idesc = ptx::make_idesc_f16(BM, BN, FMT);
    const uint32_t a_lo0 = ptx::sw128_desc_lo(ptx::smem_u32(smem_a));
    const uint32_t b_lo0 = ptx::sw128_desc_lo(ptx::smem_u32(smem_b));
    uint32_t it = 0, unit_it = 0;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++unit_it) {
      const int buf = unit_it & 1;
      ptx::mbar_wait(&t_empty[buf], ((unit_it >> 1) & 1) ^ 1);
      ptx::tc_fence_after();
      const uint32_t tacc = tmem_base + buf * BN;
      for (int kc = 0; kc < KCHUNKS; ++kc, ++it) {
        const int s = it % STAGES;
        ptx::mbar_wait(&full[s], (it / STAGES) & 1);
        ptx::tc_fence_after();
        const uint32_t a_lo = a_lo0 + s * (A_BYTES >> 4), b_lo = b_lo0 + s * (B_BYTES >> 4);
        if (ptx::elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            ptx::umma_f16<1>(tacc, ptx::sw128_desc(a_lo + kk * 2), ptx::sw128_desc(b_lo + kk * 2), idesc,
                             (kc | kk) != 0);
          ptx::umma_commit<1>(&empty[s]);
          if (kc == KCHUNKS - 1) ptx::umma_commit<1>(&t_full[buf]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue
    const int e = warp - 4;
    const int quarter = e & 3, chalf = e >> 2;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    uint8_t* stage = smem_o + e * 4096;
    uint32_t unit_it = 0;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++unit_it) {
      const int nt = u % N_TILES, mt = u / N_TILES;
      const int buf = unit_it & 1;
      ptx::mbar_wait(&t_full[buf], (unit_it >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t tacc = tmem_base + lane_base + buf * BN + chalf * (BN / 2);
#pragma unroll 1
      for (int ch = 0; ch < BN / 2 / 64; ++ch) {
        const int col0 = nt * BN + chalf * (BN / 2) + ch * 64;   // global output column = o * 512 + co
        const int co0 = col0 % COUT;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t v[32];
          ptx::tmem_ld16(tacc + ch * 64 + half * 32, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
          ptx::tmem_ld16(tacc + ch * 64 + half * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
          ptx::tmem_ld_wait();
          const float4* sc4 = reinterpret_cast<const float4*>(s_scale + co0 + half * 32);
          const float4* sh4 = reinterpret_cast<const float4*>(s_shift + co0 + half * 32);
          uint32_t o[16];
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
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            const int chunk = (half * 4 + c4) ^ (lane & 7);
            *reinterpret_cast<uint4*>(stage + lane * 128 + chunk * 16) =
                make_uint4(o[4 * c4], o[4 * c4 + 1], o[4 * c4 + 2], o[4 * c4 + 3]);
          }
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = it * 4 + (lane >> 3);
          const int c16 = lane & 7;
          const uint4 val = *reinterpret_cast<const uint4*>(stage + r * 128 + ((c16 ^ (r & 7)) * 16));
          const int n = mt * BM + quarter * 32 + r;
          A3D_DEV_CHECK(col0 >= 0 && col0 + c16 * 8 + 8 <= NDIM && n >= 0);
          if (n < n_rows) *reinterpret_cast<uint4*>(out + (size_t)n * NDIM + col0 + c16 * 8) = val;
        }
        __syncwarp();
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&t_empty[buf]);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<1>(tmem_base, 512);
}

}  // namespace

int launch_gemm_l1(const CUtensorMap& tmap_a0, const CUtensorMap& tmap_mt, void* a1, const float* scale,
                   const float* shift, int64_t n, int64_t n_alloc, int fmt, int act, int num_sms, cudaStream_t st,
                   int64_t* launches) {
  if (n <= 0) return A3D_OK;
  const int m_tiles = (int)((n + BM - 1) / BM);
  const int total = m_tiles * N_TILES;
  const int grid = total < num_sms ? total : num_sms;
  auto launch = [&](auto kern) -> int {
    A3D_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    A3D_CUDA_OK(launch_chain(kern, dim3(grid), dim3(kThreads), SMEM_BYTES, st, 1, tmap_a0, tmap_mt,
                             reinterpret_cast<uint16_t*>(a1), scale, shift, m_tiles, (int)n_alloc));
    return A3D_OK;
  };
  int rc;
  if (fmt == A3D_DTYPE_F16) {
    switch (act) {
      case A3D_ACT_ELU: rc = launch(gemm_l1_kernel<A3D_DTYPE_F16, A3D_ACT_ELU>); break;
      case A3D_ACT_RELU: rc = launch(gemm_l1_kernel<A3D_DTYPE_F16, A3D_ACT_RELU>); break;
      case A3D_ACT_LRELU: rc = launch(gemm_l1_kernel<A3D_DTYPE_F16, A3D_ACT_LRELU>); break;
      default: rc = launch(gemm_l1_kernel<A3D_DTYPE_F16, A3D_ACT_NONE>); break;
    }
  } else {
    switch (act) {
      case A3D_ACT_ELU: rc = launch(gemm_l1_kernel<A3D_DTYPE_BF16, A3D_ACT_ELU>); break;
      case A3D_ACT_RELU: rc = launch(gemm_l1_kernel<A3D_DTYPE_BF16, A3D_ACT_RELU>); break;
      case A3D_ACT_LRELU: rc = launch(gemm_l1_kernel<A3D_DTYPE_BF16, A3D_ACT_LRELU>); break;
      default: rc = launch(gemm_l1_kernel<A3D_DTYPE_BF16, A3D_ACT_NONE>); break;
    }
  }
  if (rc == A3D_OK && launches) ++*launches;
  return rc;
}

}  // namespace a3d
