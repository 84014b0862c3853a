// Fused tail of the anytime path on tcgen05 (sm_100a):
//   final Conv3DTranspose(64 -> 1, k4, s2, 'same', no bias, no BN) + tf.sigmoid   autoencoder3D.py:129-136
//   mean over the K post-sigmoid grids of an object                                 nolbo_test.py:167-177
//   yPred = (mean >= thr), TP / FP / FN against the bit-packed target               function.py:100-115
//
// Cout = 1 makes the full gather form GEMV-like, so the layer is split per axis:
//   w axis  -- gathered inside the GEMM:  Zw[j, (td, th, pw)] = X[j]*Wa + X[j - e_w]*Wb0 + X[j + e_w]*Wb1
//              (M = input voxels, N = 32 = 4 x 4 taps of (d, h) x 2 output parities of w, K = 64 channels; Wb0 / Wb1 hold
//              the delta_w = -1 / +1 taps in the pw = 0 / pw = 1 columns and zeros elsewhere).  GEMM rows are ordered
//              (w, d, h) with w slowest, so a shift of +-1 in w is a shift of 64 rows = 8 KB = whole swizzle atoms and is
//              just another start address of the A descriptor over the same TMA-loaded tile.
//   h axis  -- col2im on the accumulators with warp shuffles (a warp holds 4 d-rows of 8 consecutive h).
//   d axis  -- col2im through one shared-memory exchange.
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
constexpr int kABytes = kPos * 128;             // 64 KB tile per stage
constexpr int kPad = 64 * 128;                  // 8 KB of zeros before / between / after the tiles (w-shifted views)
constexpr int kAStride = kABytes + kPad;
constexpr int kWRows = 96;                      // Wa | Wb0 | Wb1, 32 rows each
constexpr int kWBytes = kWRows * 128;
constexpr int kExD = 8 * kPos * 4;              // d-exchange: 8 values per voxel (double buffered)
constexpr int kEpiWarps = 16;
constexpr int kThreads = 128 + 32 * kEpiWarps;  // 640
constexpr int kSmem = 1024 + kPad + 2 * kAStride + kWBytes + 2 * kExD + 16 * 8 + 16;

// The d-axis exchange only couples rows r and r +- 8 of one 64-row (d, h) plane group = the two epilogue warps 2j, 2j + 1:
// a 64-thread named barrier per warp pair (ids 2..9) instead of a 512-thread barrier per sample
__device__ __forceinline__ void pair_sync(int pair) { asm volatile("bar.sync %0, 64;" ::"r"(2 + pair) : "memory"); }

template <int FMT>
__global__ void __launch_bounds__(kThreads, 1)
tail_tc_kernel(const __grid_constant__ CUtensorMap tmap_a4, const __grid_constant__ CUtensorMap tmap_w5, int64_t B,
               int K, int final_sigmoid, const uint8_t* __restrict__ target_bits, float thr,
               unsigned long long* __restrict__ counts, float* __restrict__ mean_prob, float gamma,
               double* __restrict__ loss) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128-byte swizzle, computed on the shared-window address so the pointer keeps its
  // __shared__ provenance (LDS/STS instead of generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem + kPad;                 // tile s at smem_a + s * kAStride, zero pads around
  uint8_t* smem_w = smem + kPad + 2 * kAStride;
  float* exD = reinterpret_cast<float*>(smem_w + kWBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(exD + 2 * 8 * kPos);
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
    ptx::tmem_alloc<1>(tmem_slot, 256);
    ptx::tmem_relinquish<1>();
  }
  // zero the three pads: rows read by the w-shifted A views at the block faces must contribute exactly 0
  for (int i = threadIdx.x; i < 3 * (kPad / 16); i += blockDim.x) {
    const int p = i / (kPad / 16), o = i % (kPad / 16);
    *reinterpret_cast<uint4*>(smem + p * kAStride + o * 16) = make_uint4(0, 0, 0, 0);
  }
  ptx::fence_proxy_async();         // generic-proxy zeros visible to the tensor-core (async proxy) reads
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
          // tensor-map dims are (c, h, d, w, n): rows land as (w, d, h) with h fastest
          ptx::tma_load_5d(smem_a + s * kAStride, &tmap_a4, &a_full[s], 0, ah, ad, aw, (int)(b * K + k));
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (warp converged, issue predicated on an elected lane)
    {
      constexpr uint32_t idesc = ptx::make_idesc_f16(128, 32, FMT);
      ptx::mbar_wait(w_full, 0);
      const uint32_t w_lo = ptx::sw128_desc_lo(ptx::smem_u32(smem_w));
      const uint32_t a_lo0 = ptx::sw128_desc_lo(ptx::smem_u32(smem_a));
      constexpr uint32_t WSH = kPad >> 4;        // one step along w = 64 rows
      uint32_t it = 0;
      for (int64_t item = blockIdx.x; item < total_items; item += gridDim.x) {
        for (int k = 0; k < K; ++k, ++it) {
          const int s = it & 1;
          ptx::mbar_wait(&t_empty[s], ((it >> 1) & 1) ^ 1);
          ptx::mbar_wait(&a_full[s], (it >> 1) & 1);
          ptx::tc_fence_after();
          const uint32_t a_lo = a_lo0 + s * (kAStride >> 4);
          if (ptx::elect_one()) {
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              const uint32_t tacc = tmem_base + s * 128 + m * 32;
              const uint32_t am = a_lo + m * (16384 >> 4);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                ptx::umma_f16<1>(tacc, ptx::sw128_desc(am + kk * 2), ptx::sw128_desc(w_lo + kk * 2), idesc, kk > 0);
                ptx::umma_f16<1>(tacc, ptx::sw128_desc(am - WSH + kk * 2),
                                 ptx::sw128_desc(w_lo + ((32 * 128) >> 4) + kk * 2), idesc, 1);
                ptx::umma_f16<1>(tacc, ptx::sw128_desc(am + WSH + kk * 2),
                                 ptx::sw128_desc(w_lo + ((64 * 128) >> 4) + kk * 2), idesc, 1);
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
    // ===================================================== epilogue: col2im (h, d) + sigmoid + K-mean + threshold + counts
    const int e = warp - 4;
    const int m = e >> 2;                      // M-tile (two w-slices of the block)
    const int quarter = e & 3;                 // TMEM lane quarter == warp % 4
    const int r = m * 128 + quarter * 32 + lane;   // voxel index in the block: (lw*8 + ld)*8 + lh
    const int lw = r >> 6, ld = (r >> 3) & 7, lh = r & 7;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const float invk = 1.f / (float)K;
    const int rm = (r - 8) & (kPos - 1), rp = (r + 8) & (kPos - 1);   // d - 1 / d + 1 neighbours
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
        const uint32_t tacc = tmem_base + lane_base + s * 128 + m * 32;
        uint32_t y[32];   // [td][th][pw]
        ptx::tmem_ld16(tacc, *reinterpret_cast<uint32_t(*)[16]>(&y[0]));
        ptx::tmem_ld16(tacc + 16, *reinterpret_cast<uint32_t(*)[16]>(&y[16]));
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(&t_empty[s]);   // accumulator is in registers: release the TMEM buffer
        // ---- h axis (lanes are 8 consecutive h): out_h[ph=0][j] = Z_j[th=1] + Z_{j-1}[th=3]; [ph=1] = Z_j[th=2] + Z_{j+1}[th=0]
        float zh[4][2][2];  // [td][ph][pw]
#pragma unroll
        for (int td = 0; td < 4; ++td)
#pragma unroll
          for (int pw = 0; pw < 2; ++pw) {
            const float up = __shfl_up_sync(0xffffffffu, __uint_as_float(y[(td * 4 + 3) * 2 + pw]), 1);
            const float dn = __shfl_down_sync(0xffffffffu, __uint_as_float(y[(td * 4 + 0) * 2 + pw]), 1);
            zh[td][0][pw] = __uint_as_float(y[(td * 4 + 1) * 2 + pw]) + up;
            zh[td][1][pw] = __uint_as_float(y[(td * 4 + 2) * 2 + pw]) + dn;
          }
        // ---- d axis through shared memory (double buffered across samples: one block sync per sample)
        float* ex = exD + (it & 1) * (8 * kPos);
#pragma unroll
        for (int ph = 0; ph < 2; ++ph)
#pragma unroll
          for (int pw = 0; pw < 2; ++pw) {
            ex[((0 * 2 + ph) * 2 + pw) * kPos + r] = zh[3][ph][pw];
            ex[((1 * 2 + ph) * 2 + pw) * kPos + r] = zh[0][ph][pw];
          }
        pair_sync(e >> 1);
#pragma unroll
        for (int ph = 0; ph < 2; ++ph)
#pragma unroll
          for (int pw = 0; pw < 2; ++pw) {
            const float o0 = zh[1][ph][pw] + ex[((0 * 2 + ph) * 2 + pw) * kPos + rm];
            const float o1 = zh[2][ph][pw] + ex[((1 * 2 + ph) * 2 + pw) * kPos + rp];
            // final_sigmoid 1 (counts only): sigmoid(x) = 0.5 + 0.5 tanh(x / 2), sum the tanh terms (one MUFU each), the
            // affine part is applied to the mean;  2 (probabilities / loss emitted): 1 / (1 + exp(-x))
            psum[(0 * 2 + ph) * 2 + pw] += final_sigmoid == 1 ? ptx::tanh_approx(0.5f * o0)
                                           : final_sigmoid ? __fdividef(1.f, 1.f + __expf(-o0)) : o0;
            psum[(1 * 2 + ph) * 2 + pw] += final_sigmoid == 1 ? ptx::tanh_approx(0.5f * o1)
                                           : final_sigmoid ? __fdividef(1.f, 1.f + __expf(-o1)) : o1;
          }
      }
      // ---- finalize the block: mean over K, threshold, compare with the target bits
      int tp = 0, fp = 0, fn = 0;
      float lsum = 0.f;
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const int pd = p >> 2, ph = (p >> 1) & 1, pw = p & 1;
        const int od = 2 * (ad + ld) + pd, oh = 2 * (ah + lh) + ph, ow = 2 * (aw + lw) + pw;
        const bool ok = (pd ? ld <= 6 : ld >= 1) && (ph ? lh <= 6 : lh >= 1) && (pw ? lw <= 6 : lw >= 1) &&
                        od >= 0 && od < 64 && oh >= 0 && oh < 64 && ow >= 0 && ow < 64;
        if (!ok) continue;
        const float mval = final_sigmoid == 1 ? fmaf(psum[p], 0.5f * invk, 0.5f) : psum[p] * invk;
        const size_t v = ((size_t)od * 64 + oh) * 64 + ow;
        if (mean_prob) mean_prob[(size_t)b * A3D_VOXELS + v] = mval;
        if (target_bits) {
          const int t = (target_bits[(size_t)b * (A3D_VOXELS / 8) + (v >> 3)] >> (v & 7)) & 1;
          const int yv = mval >= thr;
          tp += t & yv;
          fp += (1 - t) & yv;
          fn += t & (1 - yv);
          if (loss) {   // weighted BCE, function.py:73-82: clip to [1e-7, 1 - 1e-7] in fp32 like tf.clip_by_value
            const float pc = fminf(fmaxf(mval, 1e-7f), 1.f - 1e-7f);
            lsum -= t ? gamma * logf(pc) : (1.f - gamma) * logf(1.f - pc);
          }
        }
      }
      if (target_bits) {
        if (loss) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
          if (lane == 0) atomicAdd(loss + b, (double)lsum);
        }
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
  if (warp == 2) ptx::tmem_dealloc<1>(tmem_base, 256);
}

}  // namespace

int launch_tail_tc(const CUtensorMap& tmap_a4, const CUtensorMap& tmap_w5, int64_t B, int K, int fmt,
                   int final_sigmoid, const uint8_t* target_bits, float thr, unsigned long long* counts,
                   float* mean_prob, float gamma, double* loss, int num_sms, cudaStream_t st, int64_t* launches) {
  if (B <= 0) return A3D_OK;
  const int64_t items = B * kItemsPerObj;
  const int grid = (int)(items < num_sms ? items : num_sms);
  auto launch = [&](auto kern) -> int {
    A3D_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    // sigmoid mode: 1 = tanh form for the counts-only path, 2 = exp form whenever probabilities or the loss are emitted
    const int sig = !final_sigmoid ? 0 : (mean_prob || loss) ? 2 : 1;
    kern<<<grid, kThreads, kSmem, st>>>(tmap_a4, tmap_w5, B, K, sig, target_bits, thr, counts, mean_prob, gamma, loss);
    A3D_CUDA_OK(cudaGetLastError());
    return A3D_OK;
  };
  const int rc = fmt == A3D_DTYPE_F16 ? launch(tail_tc_kernel<A3D_DTYPE_F16>) : launch(tail_tc_kernel<A3D_DTYPE_BF16>);
  if (rc == A3D_OK && launches) ++*launches;
  return rc;
}

}  // namespace a3d
