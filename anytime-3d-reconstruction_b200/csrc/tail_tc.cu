// Fused tail of the anytime path on tcgen05 (sm_100a):
//   final Conv3DTranspose(64 -> 1, k4, s2, 'same', no bias, no BN) + tf.sigmoid   autoencoder3D.py:129-136
//   mean over the K post-sigmoid grids of an object                                 nolbo_test.py:167-177
//   yPred = (mean >= thr), TP / FP / FN against the bit-packed target               function.py:100-115
//
// Cout = 1 makes the gather form GEMV-like, so the layer is run in SCATTER form instead: a dense tap GEMM
//     Y[j, t] = sum_ci X[j, ci] * W5[t, ci]          (M = input voxels, N = 64 taps, K = 64 channels)
// on the tensor cores, followed by col2im  out[2j + t - 1] += Y[j, t]  done separably on the accumulators:
//   w axis: warp shuffles (a warp holds 4 rows of 8 consecutive w),  h and d axes: two shared-memory exchanges.
// One CTA owns an 8x8x8 block of input voxels (origin -1 + 7*i per axis, TMA zero-fills the halo) = 14^3 complete
// output voxels, loops over the K samples of the object with double-buffered TMA stages / TMEM accumulators, keeps the
// running sum of sigmoids in registers and finally thresholds, compares with the target bits and reduces the counts.
#include <cstdlib>

#include "cvt.cuh"
#include "internal.h"
#include "ptx.cuh"

namespace a3d {
namespace {

constexpr int kBlk = 8;                         // input voxels per axis per CTA block
constexpr int kPos = kBlk * kBlk * kBlk;        // 512 GEMM rows = 4 M-tiles of 128
constexpr int kBlocksAxis = 5;                  // origins -1, 6, 13, 20, 27 cover outputs 0..63
constexpr int kItemsPerObj = kBlocksAxis * kBlocksAxis * kBlocksAxis;
constexpr int kABytes = kPos * 128;             // 64 KB per stage
constexpr int kWBytes = 64 * 128;               // W5 as the B operand: 64 taps x 64 ci
constexpr int kExH = 16 * kPos * 4;             // h-exchange: 16 values per voxel
constexpr int kExD = 8 * kPos * 4;              // d-exchange: 8 values per voxel
constexpr int kEpiWarps = 16;
constexpr int kThreads = 128 + 32 * kEpiWarps;  // 640
constexpr int kSmem = 1024 + 2 * kABytes + kWBytes + kExH + kExD + 16 * 8 + 16;

__device__ __forceinline__ void epi_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(32 * kEpiWarps) : "memory"); }

template <int FMT>
__global__ void __launch_bounds__(kThreads, 1)
tail_tc_kernel(const __grid_constant__ CUtensorMap tmap_a4, const __grid_constant__ CUtensorMap tmap_w5, int64_t B,
               int K, int final_sigmoid, const uint8_t* __restrict__ target_bits, float thr,
               unsigned long long* __restrict__ counts, float* __restrict__ mean_prob) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_w = smem + 2 * kABytes;
  float* exH = reinterpret_cast<float*>(smem_w + kWBytes);
  float* exD = exH + 16 * kPos;
  uint64_t* bars = reinterpret_cast<uint64_t*>(exD + 8 * kPos);
  uint64_t* a_full = bars;          // [2]
  uint64_t* a_empty = bars + 2;     // [2]
  uint64_t* t_full = bars + 4;      // [2]
  uint64_t* t_empty = bars + 6;     // [2]
  uint64_t* w_full = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t total_items = B * kItemsPerObj;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a4);
    ptx::prefetch_tmap(&tmap_w5);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&a_full[i], 1);
      ptx::mbar_init(&a_empty[i], 1);
      ptx::mbar_init(&t_full[i], 1);
      ptx::mbar_init(&t_empty[i], 32 * kEpiWarps);
    }
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

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      ptx::mbar_expect_tx(w_full, kWBytes);
      ptx::tma_load_2d(smem_w, &tmap_w5, w_full, 0, 0);
      uint32_t it = 0;
      for (int64_t item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int64_t b = item / kItemsPerObj;
        const int blk = (int)(item % kItemsPerObj);
        const int ad = -1 + 7 * (blk / 25), ah = -1 + 7 * ((blk / 5) % 5), aw = -1 + 7 * (blk % 5);
        for (int k = 0; k < K; ++k, ++it) {
          const int s = it & 1;
          ptx::mbar_wait(&a_empty[s], ((it >> 1) & 1) ^ 1);
          ptx::mbar_expect_tx(&a_full[s], kABytes);
          ptx::tma_load_5d(smem_a + s * kABytes, &tmap_a4, &a_full[s], 0, aw, ah, ad, (int)(b * K + k));
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (warp converged, issue predicated on an elected lane)
    {
      constexpr uint32_t idesc = ptx::make_idesc_f16(128, 64, FMT);
      ptx::mbar_wait(w_full, 0);
      const uint32_t w_lo = ptx::sw128_desc_lo(ptx::smem_u32(smem_w));
      const uint32_t a_lo0 = ptx::sw128_desc_lo(ptx::smem_u32(smem_a));
      uint32_t it = 0;
      for (int64_t item = blockIdx.x; item < total_items; item += gridDim.x) {
        for (int k = 0; k < K; ++k, ++it) {
          const int s = it & 1;
          ptx::mbar_wait(&t_empty[s], ((it >> 1) & 1) ^ 1);
          ptx::mbar_wait(&a_full[s], (it >> 1) & 1);
          ptx::tc_fence_after();
          const uint32_t a_lo = a_lo0 + s * (kABytes >> 4);
          if (ptx::elect_one()) {
#pragma unroll
          for (int m = 0; m < 4; ++m) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              ptx::umma_f16<1>(tmem_base + s * 256 + m * 64, ptx::sw128_desc(a_lo + m * (16384 >> 4) + kk * 2),
                               ptx::sw128_desc(w_lo + kk * 2), idesc, kk > 0);
            }
          }
          ptx::umma_commit<1>(&a_empty[s]);
          ptx::umma_commit<1>(&t_full[s]);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue: col2im + sigmoid + K-mean + threshold + counts
    const int e = warp - 4;
    const int m = e >> 2;                      // M-tile (two d-slices of the block)
    const int quarter = e & 3;                 // TMEM lane quarter == warp % 4
    const int r = m * 128 + quarter * 32 + lane;   // voxel index in the block: (ld*8 + lh)*8 + lw
    const int ld = r >> 6, lh = (r >> 3) & 7, lw = r & 7;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const float invk = 1.f / (float)K;
    uint32_t it = 0;
    for (int64_t item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int64_t b = item / kItemsPerObj;
      const int blk = (int)(item % kItemsPerObj);
      const int ad = -1 + 7 * (blk / 25), ah = -1 + 7 * ((blk / 5) % 5), aw = -1 + 7 * (blk % 5);
      float psum[8];
#pragma unroll
      for (int p = 0; p < 8; ++p) psum[p] = 0.f;
      for (int k = 0; k < K; ++k, ++it) {
        const int s = it & 1;
        ptx::mbar_wait(&t_full[s], (it >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t tacc = tmem_base + lane_base + s * 256 + m * 64;
        // ---- w axis: out_w[pw=0][j] = Y_j[tw=1] + Y_{j-1}[tw=3];  out_w[pw=1][j] = Y_j[tw=2] + Y_{j+1}[tw=0]
        float zw[4][4][2];  // [td][th][pw]
#pragma unroll
        for (int td = 0; td < 4; ++td) {
          uint32_t y[16];   // [th][tw]
          ptx::tmem_ld16(tacc + td * 16, y);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int th = 0; th < 4; ++th) {
            const float up = __shfl_up_sync(0xffffffffu, __uint_as_float(y[th * 4 + 3]), 1);
            const float dn = __shfl_down_sync(0xffffffffu, __uint_as_float(y[th * 4 + 0]), 1);
            zw[td][th][0] = __uint_as_float(y[th * 4 + 1]) + up;
            zw[td][th][1] = __uint_as_float(y[th * 4 + 2]) + dn;
          }
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(&t_empty[s]);   // accumulator is in registers: release the TMEM buffer
        // ---- h axis through shared memory
#pragma unroll
        for (int td = 0; td < 4; ++td)
#pragma unroll
          for (int pw = 0; pw < 2; ++pw) {
            exH[((td * 2 + 0) * 2 + pw) * kPos + r] = zw[td][3][pw];
            exH[((td * 2 + 1) * 2 + pw) * kPos + r] = zw[td][0][pw];
          }
        epi_sync(1);
        float zh[4][2][2];  // [td][ph][pw]
        const int rm = (r - 8) & (kPos - 1), rp = (r + 8) & (kPos - 1);
#pragma unroll
        for (int td = 0; td < 4; ++td)
#pragma unroll
          for (int pw = 0; pw < 2; ++pw) {
            zh[td][0][pw] = zw[td][1][pw] + exH[((td * 2 + 0) * 2 + pw) * kPos + rm];
            zh[td][1][pw] = zw[td][2][pw] + exH[((td * 2 + 1) * 2 + pw) * kPos + rp];
          }
        // ---- d axis through shared memory
#pragma unroll
        for (int ph = 0; ph < 2; ++ph)
#pragma unroll
          for (int pw = 0; pw < 2; ++pw) {
            exD[((0 * 2 + ph) * 2 + pw) * kPos + r] = zh[3][ph][pw];
            exD[((1 * 2 + ph) * 2 + pw) * kPos + r] = zh[0][ph][pw];
          }
        epi_sync(2);
        const int dm = (r - 64) & (kPos - 1), dp = (r + 64) & (kPos - 1);
#pragma unroll
        for (int ph = 0; ph < 2; ++ph)
#pragma unroll
          for (int pw = 0; pw < 2; ++pw) {
            const float o0 = zh[1][ph][pw] + exD[((0 * 2 + ph) * 2 + pw) * kPos + dm];
            const float o1 = zh[2][ph][pw] + exD[((1 * 2 + ph) * 2 + pw) * kPos + dp];
            psum[(0 * 2 + ph) * 2 + pw] += final_sigmoid ? __fdividef(1.f, 1.f + __expf(-o0)) : o0;
            psum[(1 * 2 + ph) * 2 + pw] += final_sigmoid ? __fdividef(1.f, 1.f + __expf(-o1)) : o1;
          }
      }
      // ---- finalize the block: mean over K, threshold, compare with the target bits
      int tp = 0, fp = 0, fn = 0;
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const int pd = p >> 2, ph = (p >> 1) & 1, pw = p & 1;
        const int od = 2 * (ad + ld) + pd, oh = 2 * (ah + lh) + ph, ow = 2 * (aw + lw) + pw;
        const bool ok = (pd ? ld <= 6 : ld >= 1) && (ph ? lh <= 6 : lh >= 1) && (pw ? lw <= 6 : lw >= 1) &&
                        od >= 0 && od < 64 && oh >= 0 && oh < 64 && ow >= 0 && ow < 64;
        if (!ok) continue;
        const float mval = psum[p] * invk;
        const size_t v = ((size_t)od * 64 + oh) * 64 + ow;
        if (mean_prob) mean_prob[(size_t)b * A3D_VOXELS + v] = mval;
        if (target_bits) {
          const int t = (target_bits[(size_t)b * (A3D_VOXELS / 8) + (v >> 3)] >> (v & 7)) & 1;
          const int yv = mval >= thr;
          tp += t & yv;
          fp += (1 - t) & yv;
          fn += t & (1 - yv);
        }
      }
      if (target_bits) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          tp += __shfl_xor_sync(0xffffffffu, tp, o);
          fp += __shfl_xor_sync(0xffffffffu, fp, o);
          fn += __shfl_xor_sync(0xffffffffu, fn, o);
        }
        if (lane == 0) {
          if (tp) atomicAdd(counts + b * 3 + 0, (unsigned long long)tp);
          if (fp) atomicAdd(counts + b * 3 + 1, (unsigned long long)fp);
          if (fn) atomicAdd(counts + b * 3 + 2, (unsigned long long)fn);
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<1>(tmem_base, 512);
}

}  // namespace

int launch_tail_tc(const CUtensorMap& tmap_a4, const CUtensorMap& tmap_w5, int64_t B, int K, int fmt,
                   int final_sigmoid, const uint8_t* target_bits, float thr, unsigned long long* counts,
                   float* mean_prob, int num_sms, cudaStream_t st, int64_t* launches) {
  if (B <= 0) return A3D_OK;
  const int64_t items = B * kItemsPerObj;
  const int grid = (int)(items < num_sms ? items : num_sms);
  auto launch = [&](auto kern) -> int {
    A3D_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    kern<<<grid, kThreads, kSmem, st>>>(tmap_a4, tmap_w5, B, K, final_sigmoid, target_bits, thr, counts, mean_prob);
    A3D_CUDA_OK(cudaGetLastError());
    return A3D_OK;
  };
  const int rc = fmt == A3D_DTYPE_F16 ? launch(tail_tc_kernel<A3D_DTYPE_F16>) : launch(tail_tc_kernel<A3D_DTYPE_BF16>);
  if (rc == A3D_OK && launches) ++*launches;
  return rc;
}

}  // namespace a3d
