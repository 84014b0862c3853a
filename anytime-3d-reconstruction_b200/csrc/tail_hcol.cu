// Fused tail (the tcgen05 tail kernel):
//   final Conv3DTranspose(64 -> 1, k4, s2, 'same', no bias, no BN) + tf.sigmoid   autoencoder3D.py:129-136
//   mean over the K post-sigmoid grids of an object                                 nolbo_test.py:167-177
//   yPred = (mean >= thr), TP / FP / FN (+ weighted BCE) against the bit-packed target   function.py:100-115,73-82
// with every activation tile read ONCE by the tensor core.
//
// A CTA block is 4 (w) x 4 (d) x 32 (h) input voxels of ONE sample = four w-slices of 128 GEMM rows; M-tile m is exactly
// w-slice m and the rows of a tile are (d 4, h 32): a warp (32 TMEM lanes) is one d line with the WHOLE h axis, so there
// is no h halo and any K works.
// ONE MMA per (tile, K step) with N = 64:
//   columns  0..15  Za0[(td, j)]  = X . W[td, th(j), tap_w = 1]        (delta_w = 0, output parity pw = 0)
//   columns 16..31  Za1[(td, j)]  = X . W[td, th(j), tap_w = 2]        (delta_w = 0, output parity pw = 1)
//   columns 32..47  Zm[(td, j)]   = X . W[td, th(j), tap_w = 3]        (feeds output parity pw = 0 of slice m + 1)
//   columns 48..63  Zp[(td, j)]   = X . W[td, th(j), tap_w = 0]        (feeds output parity pw = 1 of slice m - 1)
// with th(j) = (j + 1) & 3, i.e. the h taps in the order 1, 2, 3, 0: (th 1, th 2) are the delta_h = 0 taps of output
// parities ph = 0 / 1 and (th 3, th 0) their delta_h = -1 / +1 partners, so every epilogue add works on an aligned
// register PAIR and is issued as one packed add.rn.f32x2 (FADD2).
// The w-axis col2im is then free: the accumulators of slices m - 1, m, m + 1 live in the SAME TMEM lanes, in different
// column blocks, so the epilogue thread of (slice m, row r) simply loads Za from block m, Zm from block m - 1 and Zp
// from block m + 1.  h axis by warp shuffles (lanes 0 / 31 sit on the grid border: their missing neighbour is the zero
// padding), d axis by one shared-memory exchange between the four warps of a tile (128-thread named barrier).  Blocks
// advance by 3 slices in w (6 complete output columns) and by 3 voxels in d.
// What bounds the kernel is the HBM -> L2 -> SM path (17.2 GB of activations per 4096 decodes, read once from DRAM);
// see profiles/r01_notes.md and r02_notes.md for the ablations (pair blocks, three-view kernel, w-sweep: all measured,
// none kept -- this is the only tcgen05 tail in the library; csrc/tail.cu is the CUDA-core cross-check).
#include <cstdlib>
#include <type_traits>

#include "cvt.cuh"
#include "internal.h"
#include "ptx.cuh"

namespace a3d {
namespace {

constexpr int kRows = 512;                       // 4 (w) x 4 (d) x 32 (h) GEMM rows = 4 M-tiles of 128
constexpr int kBlocksW = 11;                     // w origins -1 + 3i (i < 11)
constexpr int kBlocksD4 = 11;                    // d origins -1 + 3i (i < 11), h complete
constexpr int kItemsPerObj = kBlocksW * kBlocksD4;   // 121 blocks per sample
constexpr int kABytes = kRows * 128;             // 64 KB per stage
constexpr int kWRows = 64;
constexpr int kWBytes = kWRows * 128;
constexpr int kExD = 8 * kRows * 4;              // d-exchange: 8 floats per row, double buffered
constexpr int kEpiWarps = 16;
constexpr int kThreads = 128 + 32 * kEpiWarps;   // 640
constexpr int kSmem = 1024 + 2 * kABytes + kWBytes + 2 * kExD + 16 * 8 + 16;

__device__ __forceinline__ void tile_sync(int tile) { asm volatile("bar.sync %0, 128;" ::"r"(2 + tile) : "memory"); }

// SIG: 0 = linear output, 1 = sigmoid as 0.5 + 0.5 tanh(x / 2) (one MUFU per voxel; the counts-only path), 2 / 3 = sigmoid
// as 1 / (1 + exp(-x)) (full relative accuracy near 0 and 1: used whenever probabilities (2) or the BCE loss (3) are
// emitted; the loss code is compiled out of 2)
template <int FMT, int SIG>
__global__ void __launch_bounds__(kThreads, 1)
tail_hcol_kernel(const __grid_constant__ CUtensorMap tmap_a4, const __grid_constant__ CUtensorMap tmap_w5, int64_t B,
                 int K, const uint8_t* __restrict__ target_bits, float thr,
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

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a4);
    ptx::prefetch_tmap(&tmap_w5);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&a_full[i], 1);
      ptx::mbar_init(&a_empty[i], 1);
      ptx::mbar_init(&t_full[i], 1);
      ptx::mbar_init(&t_empty[i], kEpiWarps);
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
  ptx::pdl_sync();     // a4 is the previous kernel's output; counts / loss were zeroed earlier in the stream

  if (warp == 0) {
    // ===================================================== TMA producer (converged warp, elected-lane issue)
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(w_full, kWBytes);
      ptx::tma_load_2d(smem_w, &tmap_w5, w_full, 0, 0);
    }
    __syncwarp();
    uint32_t it = 0;
    for (int64_t item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int b = (int)item / kItemsPerObj;    // the launcher guarantees total_items < 2^31: 32-bit division by a constant
      const int blk = (int)item - b * kItemsPerObj;
      const int aw = -1 + 3 * (blk % kBlocksW), ad = -1 + 3 * (blk / kBlocksW);
      for (int kp = 0; kp < K; ++kp, ++it) {
        const int s = it & 1;
        ptx::mbar_wait(&a_empty[s], ((it >> 1) & 1) ^ 1);
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(&a_full[s], kABytes);
          // tensor-map dims are (c, h, d, n, w), box 64 x 32 x 4 x 1 x 4: rows land as (w, d, h) with h fastest
          ptx::tma_load_5d(smem_a + s * kABytes, &tmap_a4, &a_full[s], 0, 0, ad, b * K + kp, aw);
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
      for (int kp = 0; kp < K; ++kp, ++it) {
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
    const int rt = quarter * 32 + lane;        // row in the tile: ld * 32 + lh
    const int r = m * 128 + rt;                // row in the block
    const int ld = quarter, lh = lane, lw = m;
    constexpr int kDStep = 32, kDLast = 3;     // rows between d neighbours; last d index of a tile
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const float invk = 1.f / (float)K;
    // d - 1 / d + 1 neighbours (same sample).  Rows at the d border of the block read their own slot instead: those sums
    // only reach outputs that the `ok` mask below drops
    const int rm = ld >= 1 ? r - kDStep : r, rp = ld < kDLast ? r + kDStep : r;
    // lanes 0 / 31 sit on the grid border in h, their missing neighbour is the zero padding
    const uint64_t hmask = ptx::f2_pack(lane >= 1 ? 1.f : 0.f, lane <= 30 ? 1.f : 0.f);
    const uint64_t half2 = ptx::f2_pack(0.5f, 0.5f);
    uint64_t* exq = reinterpret_cast<uint64_t*>(exD);   // exchange buffers as (ph = 0, ph = 1) pairs: [2][4][kRows]
    uint32_t it = 0;
    for (int64_t item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int b = (int)item / kItemsPerObj;    // the launcher guarantees total_items < 2^31: 32-bit division by a constant
      const int blk = (int)item - b * kItemsPerObj;
      const int aw = -1 + 3 * (blk % kBlocksW), ad = -1 + 3 * (blk / kBlocksW);
      // this row's 2 x 2 x 2 outputs: od = 2 * d0 + pd, ...; an output is complete when the neighbour row on that side is
      // inside the block (or the output itself lies outside the grid and is dropped).  The target bytes (one per (pd, ph):
      // the pw = 0 / 1 outputs are adjacent bits) are fetched NOW so that their latency hides behind the K samples
      const int64_t obj = b;
      constexpr bool fin = true;
      const int d0 = ad + ld, h0 = lh, w0 = aw + lw;
      const bool ind = (unsigned)d0 < 32u, inw = (unsigned)w0 < 32u;
      const bool okd[2] = {ind && ld >= 1, ind && ld < kDLast};
      const bool okh[2] = {true, true};
      const bool okw[2] = {inw && lw >= 1, inw && lw <= 2};
      const int vbase = (2 * d0 * 64 + 2 * h0) * 64 + 2 * w0;
      const int bit0 = (2 * w0) & 7;
      uint32_t tbyte[4] = {0u, 0u, 0u, 0u};
      if (target_bits && fin && (okw[0] || okw[1])) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (okd[q >> 1] && okh[q & 1]) {
            A3D_DEV_CHECK(obj >= 0 && obj < B && (unsigned)(vbase + (q >> 1) * 4096 + (q & 1) * 64 + 1) < (unsigned)A3D_VOXELS);
            tbyte[q] = target_bits[(size_t)obj * (A3D_VOXELS / 8) + ((vbase + (q >> 1) * 4096 + (q & 1) * 64) >> 3)];
          }
      }
      // running sums over the K samples, [pd][pw] as (ph = 0, ph = 1) pairs; with the sigmoid: sums of tanh(logit / 2)
      uint64_t psum[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) psum[p] = 0ull;
      for (int kp = 0; kp < K; ++kp, ++it) {
        const int s = it & 1;
        ptx::mbar_wait(&t_full[s], (it >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t tblk = tmem_base + lane_base + s * 256;
        uint32_t ya[2][16];   // Za [pw][td][j]
        uint32_t zn[2][16];   // [0]: Zm of slice m - 1 -> pw = 0;  [1]: Zp of slice m + 1 -> pw = 1   [td][j]
        ptx::tmem_ld16(tblk + m * 64, ya[0]);
        ptx::tmem_ld16(tblk + m * 64 + 16, ya[1]);
        // slices outside the block: any in-range address (the values only reach outputs that are masked below)
        ptx::tmem_ld16(tblk + (m > 0 ? m - 1 : 0) * 64 + 32, zn[0]);
        ptx::tmem_ld16(tblk + (m < 3 ? m + 1 : 3) * 64 + 48, zn[1]);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&t_empty[s]);   // accumulators are in registers: release the TMEM buffer
        // ---- w axis (register pairs) and h axis (lanes are the 32 h positions):
        //      out_h[ph=0] = Z[th=1] + Z_{h-1}[th=3];  out_h[ph=1] = Z[th=2] + Z_{h+1}[th=0]
        uint64_t zh[4][2];   // [td][pw] as (ph = 0, ph = 1)
#pragma unroll
        for (int pw = 0; pw < 2; ++pw)
#pragma unroll
          for (int td = 0; td < 4; ++td) {
            const uint32_t* y = &ya[pw][td * 4];
            const uint32_t* n = &zn[pw][td * 4];
            const uint64_t c = ptx::f2_add(ptx::f2_pack_bits(y[0], y[1]), ptx::f2_pack_bits(n[0], n[1]));   // th 1, 2
            const uint64_t o = ptx::f2_add(ptx::f2_pack_bits(y[2], y[3]), ptx::f2_pack_bits(n[2], n[3]));   // th 3, 0
            float o3, o0;
            ptx::f2_unpack(o, o3, o0);
            const float up = __shfl_up_sync(0xffffffffu, o3, 1);
            const float dn = __shfl_down_sync(0xffffffffu, o0, 1);
            zh[td][pw] = ptx::f2_fma(ptx::f2_pack(up, dn), hmask, c);
          }
        // ---- d axis through shared memory (double buffered across samples: one tile barrier per sample)
        uint64_t* ex = exq + (it & 1) * (4 * kRows);
#pragma unroll
        for (int pw = 0; pw < 2; ++pw) {
          ex[(0 * 2 + pw) * kRows + r] = zh[3][pw];
          ex[(1 * 2 + pw) * kRows + r] = zh[0][pw];
        }
        tile_sync(m);
#pragma unroll
        for (int pw = 0; pw < 2; ++pw) {
          uint64_t o0 = ptx::f2_add(zh[1][pw], ex[(0 * 2 + pw) * kRows + rm]);
          uint64_t o1 = ptx::f2_add(zh[2][pw], ex[(1 * 2 + pw) * kRows + rp]);
          if constexpr (SIG == 1) {
            float a0, a1, b0, b1;
            ptx::f2_unpack(ptx::f2_mul(o0, half2), a0, a1);
            ptx::f2_unpack(ptx::f2_mul(o1, half2), b0, b1);
            o0 = ptx::f2_pack(ptx::tanh_approx(a0), ptx::tanh_approx(a1));
            o1 = ptx::f2_pack(ptx::tanh_approx(b0), ptx::tanh_approx(b1));
          } else if constexpr (SIG >= 2) {
            float a0, a1, b0, b1;
            ptx::f2_unpack(o0, a0, a1);
            ptx::f2_unpack(o1, b0, b1);
            o0 = ptx::f2_pack(__fdividef(1.f, 1.f + __expf(-a0)), __fdividef(1.f, 1.f + __expf(-a1)));
            o1 = ptx::f2_pack(__fdividef(1.f, 1.f + __expf(-b0)), __fdividef(1.f, 1.f + __expf(-b1)));
          }
          psum[0 * 2 + pw] = ptx::f2_add(psum[0 * 2 + pw], o0);
          psum[1 * 2 + pw] = ptx::f2_add(psum[1 * 2 + pw], o1);
        }
      }
      if constexpr (SIG == 1) {
        // counts-only path (no probabilities, no loss: the launcher picks SIG = 1 only then): branch-free bit masks over
        // the 8 outputs, bit p = (pd * 2 + ph) * 2 + pw -- predictions Y, targets T (two adjacent bits per byte), complete
        // outputs M -- and three popcounts.  At K = 1 this runs once per block iteration and the epilogue is issue bound.
        if (fin && target_bits) {
          const float c = 0.5f * invk;
          uint32_t Y = 0u, T = 0u;
#pragma unroll
          for (int q = 0; q < 4; ++q) {            // q = pd * 2 + pw holds (ph = 0, ph = 1)
            float v0, v1;
            ptx::f2_unpack(psum[q], v0, v1);
            const int pd = q >> 1, pw = q & 1;
            Y |= (uint32_t)(fmaf(v0, c, 0.5f) >= thr) << ((pd * 2 + 0) * 2 + pw);
            Y |= (uint32_t)(fmaf(v1, c, 0.5f) >= thr) << ((pd * 2 + 1) * 2 + pw);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) T |= ((tbyte[q] >> bit0) & 3u) << (2 * q);   // q = pd * 2 + ph
          const uint32_t M = ((okd[0] ? 0x0Fu : 0u) | (okd[1] ? 0xF0u : 0u)) & ((okh[0] ? 0x33u : 0u) | (okh[1] ? 0xCCu : 0u)) &
                             ((okw[0] ? 0x55u : 0u) | (okw[1] ? 0xAAu : 0u));
          uint32_t packed = __popc(Y & T & M) | (__popc(Y & ~T & M) << 10) | (__popc(~Y & T & M) << 20);
          packed = __reduce_add_sync(0xffffffffu, packed);
          if (lane < 3) {
            const uint32_t f = (packed >> (10 * lane)) & 1023u;
            A3D_DEV_CHECK(obj >= 0 && obj < B);
            if (f) atomicAdd(counts + obj * 3 + lane, (unsigned long long)f);
          }
        }
      } else if constexpr (SIG == 2) {
        // probabilities (and optionally counts), no loss: decoder(z) and return_grid.  Same bit order p as above.
        if (fin) {
          float mv[8];
#pragma unroll
          for (int q = 0; q < 4; ++q) {            // q = pd * 2 + pw holds (ph = 0, ph = 1)
            float v0, v1;
            ptx::f2_unpack(psum[q], v0, v1);
            const int pd = q >> 1, pw = q & 1;
            mv[(pd * 2 + 0) * 2 + pw] = v0 * invk;
            mv[(pd * 2 + 1) * 2 + pw] = v1 * invk;
          }
          if (mean_prob) {
            float* base = mean_prob + (size_t)obj * A3D_VOXELS + vbase;
#pragma unroll
            for (int q = 0; q < 4; ++q) {          // q = pd * 2 + ph; the pw = 0 / 1 outputs are adjacent floats
              if (!(okd[q >> 1] && okh[q & 1])) continue;
              float* dst = base + (q >> 1) * 4096 + (q & 1) * 64;
              A3D_DEV_CHECK(obj < B && (okw[0] ? vbase + (q >> 1) * 4096 + (q & 1) * 64 >= 0 : true) &&
                            (!(okw[0] || okw[1]) || (unsigned)(vbase + (q >> 1) * 4096 + (q & 1) * 64 + (okw[0] ? 0 : 1)) < (unsigned)A3D_VOXELS) &&
                            (!okw[1] || vbase + (q >> 1) * 4096 + (q & 1) * 64 + 1 < A3D_VOXELS));
              if (okw[0] && okw[1]) *reinterpret_cast<float2*>(dst) = make_float2(mv[2 * q], mv[2 * q + 1]);
              else if (okw[0]) dst[0] = mv[2 * q];
              else if (okw[1]) dst[1] = mv[2 * q + 1];
            }
          }
          if (target_bits) {
            uint32_t Y = 0u, T = 0u;
#pragma unroll
            for (int p = 0; p < 8; ++p) Y |= (uint32_t)(mv[p] >= thr) << p;
#pragma unroll
            for (int q = 0; q < 4; ++q) T |= ((tbyte[q] >> bit0) & 3u) << (2 * q);
            const uint32_t M = ((okd[0] ? 0x0Fu : 0u) | (okd[1] ? 0xF0u : 0u)) & ((okh[0] ? 0x33u : 0u) | (okh[1] ? 0xCCu : 0u)) &
                               ((okw[0] ? 0x55u : 0u) | (okw[1] ? 0xAAu : 0u));
            uint32_t packed = __popc(Y & T & M) | (__popc(Y & ~T & M) << 10) | (__popc(~Y & T & M) << 20);
            packed = __reduce_add_sync(0xffffffffu, packed);
            if (lane < 3) {
              const uint32_t f = (packed >> (10 * lane)) & 1023u;
              A3D_DEV_CHECK(obj >= 0 && obj < B);
            if (f) atomicAdd(counts + obj * 3 + lane, (unsigned long long)f);
            }
          }
        }
      } else if (fin) {
        uint32_t packed = 0;   // tp | fp << 10 | fn << 20 (a warp adds at most 256 per field)
        float lsum = 0.f;
#pragma unroll
        for (int pd = 0; pd < 2; ++pd)
#pragma unroll
          for (int ph = 0; ph < 2; ++ph) {
            if (!(okd[pd] && okh[ph] && (okw[0] || okw[1]))) continue;
            const int v0 = vbase + pd * 4096 + ph * 64;     // voxel index of the pw = 0 output; pw = 1 is the next bit
            A3D_DEV_CHECK(obj < B && (!(okw[0] && okw[1]) || (v0 >= 0 && v0 + 1 < A3D_VOXELS)));
            const uint32_t byte = tbyte[pd * 2 + ph];
            float mpair = 0.f;
#pragma unroll
            for (int pw = 0; pw < 2; ++pw) {
              if (!okw[pw]) continue;
              float v2[2];
              ptx::f2_unpack(psum[pd * 2 + pw], v2[0], v2[1]);
              // mean of sigmoid = 0.5 + 0.5 * mean of tanh(logit / 2)
              const float mval = SIG == 1 ? fmaf(v2[ph], 0.5f * invk, 0.5f) : v2[ph] * invk;
              if (mean_prob) {   // the pw = 0 / 1 outputs are adjacent floats: one 8-byte store when both are complete
                if (okw[0] && okw[1]) {
                  if (pw == 0) mpair = mval;
                  else *reinterpret_cast<float2*>(mean_prob + (size_t)obj * A3D_VOXELS + v0) = make_float2(mpair, mval);
                } else {
                  A3D_DEV_CHECK((unsigned)(v0 + pw) < (unsigned)A3D_VOXELS);
                  mean_prob[(size_t)obj * A3D_VOXELS + v0 + pw] = mval;
                }
              }
              if (target_bits) {
                const uint32_t t = (byte >> (bit0 + pw)) & 1u;
                const uint32_t yv = mval >= thr;
                packed += (t & yv) + (((t ^ 1u) & yv) << 10) + ((t & (yv ^ 1u)) << 20);
                if (SIG != 2 && loss) {   // weighted BCE, function.py:73-82: clip to [1e-7, 1 - 1e-7] in fp32 like tf.clip_by_value
                  const float pc = fminf(fmaxf(mval, 1e-7f), 1.f - 1e-7f);
                  lsum -= t ? gamma * logf(pc) : (1.f - gamma) * logf(1.f - pc);
                }
              }
            }
          }
        if (target_bits) {
          if (SIG != 2 && loss) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
            if (lane == 0 && lsum != 0.f) atomicAdd(loss + obj, (double)lsum);
          }
          packed = __reduce_add_sync(0xffffffffu, packed);
          if (lane < 3) {
            const uint32_t f = (packed >> (10 * lane)) & 1023u;
            A3D_DEV_CHECK(obj >= 0 && obj < B);
            if (f) atomicAdd(counts + obj * 3 + lane, (unsigned long long)f);
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<1>(tmem_base, 512);
}

}  // namespace

int launch_tail_hcol(const CUtensorMap& tmap_a4h, const CUtensorMap& tmap_w5p, int64_t B, int K, int fmt,
                     int final_sigmoid, const uint8_t* target_bits, float thr, unsigned long long* counts,
                     float* mean_prob, float gamma, double* loss, int num_sms, cudaStream_t st, int64_t* launches) {
  if (B <= 0) return A3D_OK;
  const int64_t items = B * kItemsPerObj;
  if (items > 0x7fffffffll || B * (int64_t)K > 0x7fffffffll) { set_error("tail: batch too large for one launch"); return A3D_ERR_INVALID; }
  const int grid = (int)(items < num_sms ? items : num_sms);
  auto launch = [&](auto kern) -> int {
    A3D_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    A3D_CUDA_OK(launch_chain(kern, dim3(grid), dim3(kThreads), kSmem, st, 1, tmap_a4h, tmap_w5p, B, K, target_bits, thr,
                             counts, mean_prob, gamma, loss));
    return A3D_OK;
  };
  // sigmoid mode: 1 = tanh form for the counts-only path, 2 = exp form whenever probabilities or the loss are emitted
  const int sig = !final_sigmoid ? 0 : loss ? 3 : mean_prob ? 2 : 1;
  auto pick = [&](auto fmt_c) -> int {
    constexpr int F = decltype(fmt_c)::value;
    return sig == 0 ? launch(tail_hcol_kernel<F, 0>)
                    : sig == 1 ? launch(tail_hcol_kernel<F, 1>)
                               : sig == 2 ? launch(tail_hcol_kernel<F, 2>) : launch(tail_hcol_kernel<F, 3>);
  };
  const int rc = fmt == A3D_DTYPE_F16 ? pick(std::integral_constant<int, A3D_DTYPE_F16>{})
                                      : pick(std::integral_constant<int, A3D_DTYPE_BF16>{});
  if (rc == A3D_OK && launches) ++*launches;
  return rc;
}

}  // namespace a3d
