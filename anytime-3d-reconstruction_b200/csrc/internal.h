// Internal declarations shared by the translation units of liba3d (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/a3d.h"

namespace a3d {

void set_error(const char* fmt, ...);

#define A3D_CUDA_OK(expr)                                                                     \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::a3d::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return A3D_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

constexpr float kBnEps = 1e-3f;  // Keras BatchNormalization default epsilon

// Checked build (`python anytime-3d-reconstruction_b200/build.py --checked` -> liba3d_checked.so, -DA3D_CHECKED): every
// hand-computed global-memory index of the decoder kernels (epilogue stores, target-bit reads, probability stores, count
// atomics, imputation outputs) and every shared-memory staging offset is range-checked on the device and traps with a
// message.  compute-sanitizer is closed on the GPU pool this was developed on; tests/test_gpu_checked_build.py runs the
// ragged / border cases of the parity suite through this build instead.  TMA loads need no check: the tensor map bounds
// them in hardware (out-of-range elements are zero filled = the 'same' padding).  The release build compiles the
// checks out.
#ifdef A3D_CHECKED
#define A3D_DEV_CHECK(cond)                                                                                        \
  do {                                                                                                             \
    if (!(cond)) {                                                                                                 \
      printf("A3D_DEV_CHECK failed: %s (%s:%d) block %d thread %d\n", #cond, __FILE__, __LINE__, (int)blockIdx.x,   \
             (int)threadIdx.x);                                                                                    \
      __trap();                                                                                                    \
    }                                                                                                              \
  } while (0)
#else
#define A3D_DEV_CHECK(cond) ((void)0)
#endif

// Launch a kernel of the decoder chain with programmatic stream serialization (see ptx::pdl_sync): the kernel's
// prologue overlaps the tail of the previous kernel in the stream.  A3D_PDL=0 turns the attribute off (plain
// stream-ordered launches; griddepcontrol.* are then no-ops).  `cluster` > 1 adds a run-time cluster dimension (kernels
// with a compile-time __cluster_dims__ pass 1).
bool pdl_enabled();
template <class... KArgs, class... Args>
cudaError_t launch_chain(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster,
                         Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  unsigned na = 0;
  if (cluster > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)cluster;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// host helpers shared by handle.cu (decoder) and enc2d.cu (image encoder)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn();
uint16_t cvt16(float v, int fmt);
int upload(const void* src, size_t bytes, void** dst);   // cudaMalloc on first use + synchronous H2D copy
void fold_bn(const std::vector<float>& g, const std::vector<float>& b, const std::vector<float>& m,
             const std::vector<float>& v, std::vector<float>& scale, std::vector<float>& shift);

// One stride-2 transposed-conv layer (k=4, 'same'): in [N, W,W,W, CIN] -> out [N, 2W,2W,2W, COUT], NDHWC, 16-bit.
struct ConvLayer {
  int cin = 0, cout = 0, win = 0;
  // tcgen05 path
  CUtensorMap tmap_act;   // 5-D view (c, n, w, h, d) of the input activations, box (64, 128/W, W+2, 1, 1), SW128
  CUtensorMap tmap_wgt;   // 2-D view (64 ci, rows) of the per-(parity,tap,chunk) repacked weights, box (64, 256), SW128
  CUtensorMap tmap_wgt64; // same tensor, box (64, 64): N-half loads of the 2-CTA path
  void* wgt_packed = nullptr;   // device, 16-bit
  CUtensorMap tmap_act_sw;      // w-sweep kernel (convt_l4_sw.cu): (c, n, h, w, d) view, box (64, 8, W+2, 1, 1)
  CUtensorMap tmap_wgt_sw;      // ... its resident weights: [class][rank][sd][sh][chunk][2 taps x 64 co] rows
  void* wgt_sw = nullptr;
  // SIMT path: [tap 64][ci][co] 16-bit
  void* wgt_tco = nullptr;
  float* scale = nullptr;  // folded BN: y = scale*conv + shift
  float* shift = nullptr;
};

// kernels (defined in the .cu files); all asynchronous on `st`
int launch_dense_l1(const float* z, int64_t n, int D, const float* wd, const float* bd, const float* s0,
                    const float* h0, void* a0, const void* w1_tco, const float* s1, const float* h1, void* a1,
                    int fmt, int act, bool do_s1, cudaStream_t st, int64_t* launches);
int launch_gemm_l1(const CUtensorMap& tmap_a0, const CUtensorMap& tmap_mt, void* a1, const float* scale,
                   const float* shift, int64_t n, int64_t n_alloc, int fmt, int act, int num_sms, cudaStream_t st,
                   int64_t* launches);
int launch_convt_s2_tc(const ConvLayer& L, void* out, int64_t n, int64_t n_alloc, int fmt, int act, int num_sms,
                       int force_variant, cudaStream_t st, int64_t* launches);
int launch_convt_l4_sw(const CUtensorMap& tmap_act, const CUtensorMap& tmap_wgt, void* out, const float* scale,
                       const float* shift, int64_t n, int64_t n_alloc, int fmt, int act, int num_sms, int* progress,
                       int pace_delta, cudaStream_t st, int64_t* launches);
size_t convt_l4_sw_progress_bytes();
int launch_convt_s2_simt(const ConvLayer& L, const void* in, void* out, int64_t n, int fmt, int act,
                         cudaStream_t st, int64_t* launches);
// Final ConvT(64->1, s2) + sigmoid + mean over K + threshold + TP/FP/FN (+ optional mean grid).
int launch_tail(const void* a4, const float* w5, int64_t B, int K, int fmt, int final_sigmoid,
                const uint8_t* target_bits, float thr, unsigned long long* counts, float* mean_prob, float gamma,
                double* loss, cudaStream_t st, int64_t* launches);
// tcgen05 tail (tail_hcol.cu): tmap_a4h = (c,h,d,n,w) view of act[4] with box (64,32,4,1,4); tmap_w5h = 64-row weight tile
int launch_tail_hcol(const CUtensorMap& tmap_a4h, const CUtensorMap& tmap_w5h, int64_t B, int K, int fmt,
                     int final_sigmoid, const uint8_t* target_bits, float thr, unsigned long long* counts,
                     float* mean_prob, float gamma, double* loss, int num_sms, cudaStream_t st, int64_t* launches);
int launch_binary_loss(const float* pred, const float* target, int64_t B, int64_t V, float gamma, double* loss,
                       cudaStream_t st, int64_t* launches);
int launch_counts_sweep(const float* target, const float* pred, int64_t B, int64_t V, const float* thr, int T, int strict,
                        unsigned long long* counts, cudaStream_t st, int64_t* launches);
int launch_impute(const float* z, const float* mask, const float* mu, int C, int64_t B, int K, int D, uint64_t seed,
                  uint64_t obj_offset, int fill, float* z_out, int32_t* cstar, cudaStream_t st, int64_t* launches);
int launch_sampling(const float* mu, const float* logvar, int64_t n, int D, uint64_t seed, uint64_t obj_offset, float* z,
                    cudaStream_t st, int64_t* launches);
int launch_nearest_prior(const float* z, int64_t z_stride, const float* mu, int C, int D, const float* labels, int64_t B,
                         int32_t* idx, int32_t* hits, cudaStream_t st, int64_t* launches);
int launch_grid_convert(const float* in, int64_t voxels, int out_dtype, float thr, void* out, cudaStream_t st,
                        int64_t* launches);
int launch_counts(const float* target, const float* pred, int64_t B, int64_t V, float thr,
                  unsigned long long* counts, cudaStream_t st, int64_t* launches);
int launch_pack(const float* target, int64_t B, int64_t V, uint8_t* bits, cudaStream_t st, int64_t* launches);
int launch_to_f32(const void* src, float* dst, int64_t n, int fmt, cudaStream_t st);

// ---------------------------------------------------------------------------------------------------------------
// 2-D image encoder (Darknet19 + head2D, src/net_core/darknet.py:83-168); kernels in conv2d_tc.cu / enc2d_kernels.cu
// Geometry of one stride-1 'same' Conv2D launch.  The GEMM M tile is a brick of wt x ht pixels x nt images
// (wt * ht * nt = 128, all powers of two) so that one 4-D TMA box per (tap, 64-channel chunk) fetches the A operand.
struct Conv2dGeom {
  int H = 0, W = 0;            // spatial size of the conv input (= output, stride 1)
  int lw = 0, lh = 0;          // log2(wt), log2(ht); nt = 128 >> (lw + lh)
  int tiles_w = 0, tiles_h = 0;
  int m_tiles = 0, n_tiles = 0;
  int taps = 0, cin_chunks = 0, cout_pad = 0, cout_real = 0;
  int kc = 64;                 // channels per K step (64: SWIZZLE_128B rows, 32: SWIZZLE_64B rows)
  int rh = 0;                  // 1: resident weights + haloed activation box (tmap_act box = (kc, 16, 10, 1));
                               // 2: + fused pool through four accumulators: the brick tiles the POOLED grid and tmap_act
                               //    is element-strided (box (kc, 32, 18, 1), strides (1, 2, 2, 1))
  int n_images = 0;
};
int conv2d_tc_bn(int cout_pad);
// pool: fuse the MaxPool2D(2, 2) that follows (out = [n, H/2, W/2, cout_pad] 16-bit); out_f32: out = [n*H*W, cout_real] fp32
int launch_conv2d_tc(const CUtensorMap& tmap_act, const CUtensorMap& tmap_wgt, void* out, const float* scale,
                     const float* shift, const Conv2dGeom& g, int bn, int fmt, int act, bool pool, bool out_f32,
                     int num_sms, cudaStream_t st, int64_t* launches);
// 2-CTA (cta_group::2) variant for 256-wide N tiles (conv2d_pair.cu): two bricks per cluster share every weight tile.
// tmap_wgt_half = the weight tensor with a (64, 128)-row box.
bool conv2d_pair_eligible(const Conv2dGeom& g, int bn, bool pool, bool out_f32);
int launch_conv2d_pair(const CUtensorMap& tmap_act, const CUtensorMap& tmap_wgt_half, void* out, const float* scale,
                       const float* shift, const Conv2dGeom& g, int fmt, int act, bool pool, int num_sms, cudaStream_t st,
                       int64_t* launches);
// first layer: Conv2D(3 -> 32, k3) + BN + act + MaxPool2D(2,2); in fp32 [n,H,W,3], out 16-bit [n,H/2,W/2,cout_pad]
int launch_conv2d_first_pool(const float* in, const float* w27x32, const float* scale, const float* shift, void* out,
                             int64_t n, int H, int W, int cout_pad, int fmt, int act, cudaStream_t st, int64_t* launches);
// same layer on the tensor cores (conv2d_first_tc.cu): w32x32 = 16-bit [co][k = tap*3+ci], zero-padded to K = 32
int launch_conv2d_first_tc(const float* in, const void* w32x32, const float* scale, const float* shift, void* out,
                           int64_t n, int H, int W, int cout_pad, int fmt, int act, int num_sms, cudaStream_t st,
                           int64_t* launches);
int launch_maxpool2d(const void* in, void* out, int64_t n, int H, int W, int C, int fmt, cudaStream_t st,
                     int64_t* launches);
int launch_global_pool(const float* in, float* out, int64_t n, int HW, int C, int is_max, cudaStream_t st,
                       int64_t* launches);
// NHWC channel-(un)padding copies between user buffers (fp32 or 16-bit) and the 16-bit arena
int launch_import_nhwc(const void* in, int in_is_f32, void* out, int64_t pixels, int C, int C_pad, int fmt,
                       cudaStream_t st, int64_t* launches);
int launch_u8_to_f32(const uint8_t* in, float* out, int64_t total, float scale, cudaStream_t st, int64_t* launches);
int launch_export_nhwc(const void* in, void* out, int out_is_f32, int64_t pixels, int C, int C_pad, int fmt,
                       cudaStream_t st, int64_t* launches);
int launch_split_sample(const float* enc_out, int64_t n, int D, int out_stride, float clip, int seed_enable,
                        uint64_t seed, uint64_t obj_offset, float* mean, float* logvar, float* z, cudaStream_t st,
                        int64_t* launches);

// ---------------------------------------------------------------------------------------------------------------
// Voxel encoder (encoder3D, src/net_core/autoencoder3D.py:72-102); kernels in conv3d_tc.cu
struct Conv3dGeom {
  int G = 0;                    // OUTPUT grid (cubic)
  int stride = 1;               // 1 or 2; 'same' with k = 4 -> pad_before = 1
  int lw = 0, lh = 0, ld = 0;   // log2 of the output brick; nt = 128 >> (lw + lh + ld) objects per tile
  int tiles_w = 0, tiles_h = 0, tiles_d = 0;
  int m_tiles = 0, n_tiles = 0;
  int cin_chunks = 0, cout_pad = 0, cout_real = 0;
  int n_objects = 0;
};
int launch_conv3d_tc(const CUtensorMap& tmap_act, const CUtensorMap& tmap_wgt, void* out, const float* scale,
                     const float* shift, const Conv3dGeom& g, int bn, int fmt, int act, bool out_f32, int num_sms,
                     cudaStream_t st, int64_t* launches);
// first layer: fp32 grid [n, Gin^3] -> 16-bit [n, (Gin/2)^3, 64]; w64x64 = 16-bit [co][tap]
int launch_conv3d_first_tc(const float* in, const void* w64x64, const float* scale, const float* shift, void* out,
                           int64_t n, int Gin, int fmt, int act, int num_sms, cudaStream_t st, int64_t* launches);
int launch_sigmoid_inplace(float* x, int64_t n, cudaStream_t st, int64_t* launches);

size_t convt_tc_smem_bytes(int cin, int cout, int win);

}  // namespace a3d
