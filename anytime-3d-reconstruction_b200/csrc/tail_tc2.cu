// Fused tail, "pair" variant (default for even K): same maths as tail_tc.cu --
//   final Conv3DTranspose(64 -> 1, k4, s2, 'same', no bias, no BN) + tf.sigmoid   autoencoder3D.py:129-136
//   mean over the K post-sigmoid grids of an object                                 nolbo_test.py:167-177
//   yPred = (mean >= thr), TP / FP / FN (+ weighted BCE) against the bit-packed target   function.py:100-115,73-82
// -- with 1.75x less shared-memory operand traffic per block, which is what bounds this kernel (tail_tc.cu reads every
// activation tile three times through w-shifted descriptor views with N = 32 MMAs).
//
// A CTA block is 4 (w) x 8 (d) x 8 (h) input voxels of TWO consecutive samples of one object; GEMM rows are ordered
// (w, sample, d, h), so M-tile m (128 rows) is exactly w-slice m of both samples.  ONE MMA per (tile, K step) with N = 64:
//   columns  0..31  Za[(td, th, pw)]  = X . W[td, th, tap_w = pw + 1]        (delta_w = 0)
//   columns 32..47  Zm[(td, th)]      = X . W[td, th, tap_w = 3]             (feeds output parity pw = 0 of slice m + 1)
//   columns 48..63  Zp[(td, th)]      = X . W[td, th, tap_w = 0]             (feeds output parity pw = 1 of slice m - 1)
// The w-axis col2im is then free: the accumulators of slices m - 1, m, m + 1 live in the SAME TMEM lanes, in different
// column blocks, so the epilogue thread of (slice m, row r) simply loads Za from block m, Zm from block m - 1 and Zp
// from block m + 1.  h axis by warp shuffles, d axis by one shared-memory exchange, as in tail_tc.cu.  Blocks advance by
// 3 slices in w (6 complete output columns) and by 7 voxels in d and h (14 complete outputs).
#include <cstdlib>

#include "cvt.cuh"
#include "internal.h"
#include "ptx.cuh"

namespace a3d {
namespace {

constexpr int kBw = 4, kBd = 8, kBh = 8;
constexpr int kRows = kBw * 2 * kBd * kBh;       // 512 GEMM rows = 4 M-tiles of 128
constexpr int kBlocksW = 11, kBlocksDH = 5;      // w origins -1 + 3i (i < 11), d / h origins -1 + 7i (i < 5)
constexpr int kItemsPerObj = kBlocksW * kBlocksDH * kBlocksDH;
constexpr int kABytes = kRows * 128;             // 64 KB per stage
constexpr int kWRows = 64;
constexpr int kWBytes = kWRows * 128;
constexpr int kExD = 8 * kRows * 4;              // d-exchange: 8 floats per row, double buffered
constexpr int kEpiWarps = 16;
constexpr int kThreads = 128 + 32 * kEpiWarps;   // 640
constexpr int kSmem = 1024 + 2 * kABytes + kWBytes + 2 * kExD + 16 * 8 + 16;

__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, %0;" ::"r"(32 * kEpiWarps) : "memory"); }
// The d-axis exchange only couples rows r and r +- 8 of one 64-row (d, h) plane group = the two epilogue warps 2j, 2j + 1:
// a 64-thread named barrier per warp pair (ids 2..9) instead of a 512-thread barrier per sample
__device__ __forceinline__ void pair_sync(int pair) { asm volatile("bar.sync %0, 64;" ::"r"(2 + pair) : "memory"); }

template <int FMT>
__global__ void __launch_bounds__(kThreads, 1)
tail_pair_kernel(const __grid_constant__ CUtensorMap tmap_a4, const __grid_constant__ CUtensorMap tmap_w5, int64_t B,
                 int K, int final_sigmoid, const uint8_t* __restrict__ target_bits, float thr,
                 unsigned long long* __restrict__ counts, float* __restrict__ mean_prob, float gamma,
                 double* __restrict__ loss) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_w = smem + 2 * kABytes;
  float* exD = reinterpret_cast<float*>(smem_w + kWBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(exD + 2 * 8 * kRows);
  uint64_t* a_full = bars;          // [2]
  uint64_t* a_empty = bars + 2;     // [2]
  uint64_t* t_full = bars + 4;      // [2]
  uint64_t* t_empty = bars + 6;     // [2]
  uint64_t* w_full = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t total_items = B * kItemsPerObj;
  const int pairs = K >> 1;         // K is even (the launcher falls back to tail_tc.cu otherwise)

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
    // ===================================================== TMA producer (converged warp, elected-lane issue)
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(w_full, kWBytes);
      ptx::tma_load_2d(smem_w, &tmap_w5, w_full, 0, 0);
    }
    __syncwarp();
    uint32_t it = 0;
    for (int64_t item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int64_t b = item / kItemsPerObj;
      const int blk = (int)(item % kItemsPerObj);
      const int aw = -1 + 3 * (blk % kBlocksW), ah = -1 + 7 * ((blk / kBlocksW) % kBlocksDH);
      const int ad = -1 + 7 * (blk / (kBlocksW * kBlocksDH));
      for (int kp = 0; kp < pairs; ++kp, ++it) {
        const int s = it & 1;
        ptx::mbar_wait(&a_empty[s], ((it >> 1) & 1) ^ 1);
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(&a_full[s], kABytes);
          // tensor-map dims are (c, h, d, n, w): rows land as (w, sample, d, h) with h fastest
          ptx::tma_load_5d(smem_a + s * kABytes, &tmap_a4, &a_full[s], 0, ah, ad, (int)(b * K + 2 * kp), aw);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (converged warp, elected-lane issue)
    constexpr uint32_t idesc = ptx::make_idesc_f16(128, 64, FMT);
    ptx::mbar_wait(w_full, 0);
    const uint32_t w_lo = ptx::sw128_desc_lo(ptx::smem_u32(smem_w));
    const uint32_t a_lo0 = ptx::sw128_desc_lo(ptx::smem_u32(smem_a));
    uint32_t it = 0;
    for (int64_t item = blockIdx.x; item < total_items; item += gridDim.x) {
      for (int kp = 0; kp < pairs; ++kp, ++it) {
        const int s = it & 1;
        ptx::mbar_wait(&t_empty[s], ((it >> 1) & 1) ^ 1);
        ptx::mbar_wait(&a_full[s], (it >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t a_lo = a_lo0 + s * (kABytes >> 4);
        if (ptx::elect_one()) {
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const uint32_t tacc = tmem_base + s * 256 + m * 64;
            const uint32_t am = a_lo + m * (16384 >> 4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              ptx::umma_f16<1>(tacc, ptx::sw128_desc(am + kk * 2), ptx::sw128_desc(w_lo + kk * 2), idesc, kk > 0);
          }
          ptx::umma_commit<1>(&a_empty[s]);
          ptx::umma_commit<1>(&t_full[s]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue: col2im (w, h, d) + sigmoid + K-mean + threshold + counts
    const int e = warp - 4;
    const int m = e >> 2;                      // M-tile = w-slice of the block
    const int quarter = e & 3;                 // TMEM lane quarter == warp % 4
    const int rt = quarter * 32 + lane;        // row in the tile: sample * 64 + ld * 8 + lh
    const int r = m * 128 + rt;                // row in the block
    const int slot = rt >> 6, ld = (rt >> 3) & 7, lh = rt & 7, lw = m;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const float invk = 1.f / (float)K;
    const int rm = r - 8, rp = r + 8;          // d - 1 / d + 1 neighbours (same sample; only used where ld allows)
    uint32_t it = 0;
    for (int64_t item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int64_t b = item / kItemsPerObj;
      const int blk = (int)(item % kItemsPerObj);
      const int aw = -1 + 3 * (blk % kBlocksW), ah = -1 + 7 * ((blk / kBlocksW) % kBlocksDH);
      const int ad = -1 + 7 * (blk / (kBlocksW * kBlocksDH));
      float psum[8];
#pragma unroll
      for (int p = 0; p < 8; ++p) psum[p] = 0.f;
      for (int kp = 0; kp < pairs; ++kp, ++it) {
        const int s = it & 1;
        ptx::mbar_wait(&t_full[s], (it >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t tblk = tmem_base + lane_base + s * 256;
        uint32_t y[32];    // Za [td][th][pw]
        uint32_t zm[16];   // Zm of slice m - 1 [td][th]  -> pw = 0
        uint32_t zp[16];   // Zp of slice m + 1 [td][th]  -> pw = 1
        ptx::tmem_ld16(tblk + m * 64, *reinterpret_cast<uint32_t(*)[16]>(&y[0]));
        ptx::tmem_ld16(tblk + m * 64 + 16, *reinterpret_cast<uint32_t(*)[16]>(&y[16]));
        // slices outside the block: any in-range address (the values only reach outputs that are masked below)
        ptx::tmem_ld16(tblk + (m > 0 ? m - 1 : 0) * 64 + 32, zm);
        ptx::tmem_ld16(tblk + (m < 3 ? m + 1 : 3) * 64 + 48, zp);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(&t_empty[s]);   // accumulators are in registers: release the TMEM buffer
        // ---- w axis: add the neighbours' contributions (register adds only)
        float z[32];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          z[2 * q] = __uint_as_float(y[2 * q]) + __uint_as_float(zm[q]);
          z[2 * q + 1] = __uint_as_float(y[2 * q + 1]) + __uint_as_float(zp[q]);
        }
        // ---- h axis (lanes are 8 consecutive h): out_h[ph=0][j] = Z_j[th=1] + Z_{j-1}[th=3]; [ph=1] = Z_j[th=2] + Z_{j+1}[th=0]
        float zh[4][2][2];  // [td][ph][pw]
#pragma unroll
        for (int td = 0; td < 4; ++td)
#pragma unroll
          for (int pw = 0; pw < 2; ++pw) {
            const float up = __shfl_up_sync(0xffffffffu, z[(td * 4 + 3) * 2 + pw], 1);
            const float dn = __shfl_down_sync(0xffffffffu, z[(td * 4 + 0) * 2 + pw], 1);
            zh[td][0][pw] = z[(td * 4 + 1) * 2 + pw] + up;
            zh[td][1][pw] = z[(td * 4 + 2) * 2 + pw] + dn;
          }
        // ---- d axis through shared memory (double buffered across pairs: one block sync per pair)
        float* ex = exD + (it & 1) * (8 * kRows);
#pragma unroll
        for (int ph = 0; ph < 2; ++ph)
#pragma unroll
          for (int pw = 0; pw < 2; ++pw) {
            ex[((0 * 2 + ph) * 2 + pw) * kRows + r] = zh[3][ph][pw];
            ex[((1 * 2 + ph) * 2 + pw) * kRows + r] = zh[0][ph][pw];
          }
        pair_sync(e >> 1);
#pragma unroll
        for (int ph = 0; ph < 2; ++ph)
#pragma unroll
          for (int pw = 0; pw < 2; ++pw) {
            const float o0 = zh[1][ph][pw] + (ld >= 1 ? ex[((0 * 2 + ph) * 2 + pw) * kRows + rm] : 0.f);
            const float o1 = zh[2][ph][pw] + (ld <= 6 ? ex[((1 * 2 + ph) * 2 + pw) * kRows + rp] : 0.f);
            psum[(0 * 2 + ph) * 2 + pw] += final_sigmoid ? __fdividef(1.f, 1.f + __expf(-o0)) : o0;
            psum[(1 * 2 + ph) * 2 + pw] += final_sigmoid ? __fdividef(1.f, 1.f + __expf(-o1)) : o1;
          }
      }
      // ---- combine the two sample slots (rows r and r + 64 of a tile) through shared memory, then finalize in slot 0
      float* ex = exD + (it & 1) * (8 * kRows);   // the buffer the NEXT pair would use: its last readers finished two syncs ago
      if (slot == 1) {
#pragma unroll
        for (int p = 0; p < 8; ++p) ex[p * kRows + r] = psum[p];
      }
      epi_sync();
      int tp = 0, fp = 0, fn = 0;
      float lsum = 0.f;
      if (slot == 0) {
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          const int pd = p >> 2, ph = (p >> 1) & 1, pw = p & 1;
          const int od = 2 * (ad + ld) + pd, oh = 2 * (ah + lh) + ph, ow = 2 * (aw + lw) + pw;
          const bool ok = (pd ? ld <= 6 : ld >= 1) && (ph ? lh <= 6 : lh >= 1) && (pw ? lw <= 2 : lw >= 1) &&
                          od >= 0 && od < 64 && oh >= 0 && oh < 64 && ow >= 0 && ow < 64;
          if (!ok) continue;
          const float mval = (psum[p] + ex[p * kRows + r + 64]) * invk;
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
      }
      if (target_bits) {
        if (loss) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
          if (lane == 0 && lsum != 0.f) atomicAdd(loss + b, (double)lsum);
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
      epi_sync();   // the slot exchange buffer is reused by the next item's d-exchange
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<1>(tmem_base, 512);
}

}  // namespace

int launch_tail_pair(const CUtensorMap& tmap_a4p, const CUtensorMap& tmap_w5p, int64_t B, int K, int fmt,
                     int final_sigmoid, const uint8_t* target_bits, float thr, unsigned long long* counts,
                     float* mean_prob, float gamma, double* loss, int num_sms, cudaStream_t st, int64_t* launches) {
  if (B <= 0) return A3D_OK;
  if (K < 2 || (K & 1)) { set_error("tail_pair: K must be even"); return A3D_ERR_INVALID; }
  const int64_t items = B * kItemsPerObj;
  const int grid = (int)(items < num_sms ? items : num_sms);
  auto launch = [&](auto kern) -> int {
    A3D_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    kern<<<grid, kThreads, kSmem, st>>>(tmap_a4p, tmap_w5p, B, K, final_sigmoid, target_bits, thr, counts, mean_prob, gamma,
                                        loss);
    A3D_CUDA_OK(cudaGetLastError());
    return A3D_OK;
  };
  const int rc = fmt == A3D_DTYPE_F16 ? launch(tail_pair_kernel<A3D_DTYPE_F16>) : launch(tail_pair_kernel<A3D_DTYPE_BF16>);
  if (rc == A3D_OK && launches) ++*launches;
  return rc;
}

}  // namespace a3d
