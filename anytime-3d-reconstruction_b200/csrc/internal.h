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

// One stride-2 transposed-conv layer (k=4, 'same'): in [N, W,W,W, CIN] -> out [N, 2W,2W,2W, COUT], NDHWC, 16-bit.
struct ConvLayer {
  int cin = 0, cout = 0, win = 0;
  // tcgen05 path
  CUtensorMap tmap_act;   // 5-D view (c, n, w, h, d) of the input activations, box (64, 128/W, W+2, 1, 1), SW128
  CUtensorMap tmap_wgt;   // 2-D view (64 ci, rows) of the per-(parity,tap,chunk) repacked weights, box (64, 256), SW128
  CUtensorMap tmap_wgt64; // same tensor, box (64, 64): N-half loads of the 2-CTA path
  void* wgt_packed = nullptr;   // device, 16-bit
  CUtensorMap tmap_wgt_ws;      // weight-stationary 2-CTA layout (128->64 layer only)
  void* wgt_ws = nullptr;
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
                       cudaStream_t st, int64_t* launches);
int launch_convt_l4_ws(const CUtensorMap& tmap_act, const CUtensorMap& tmap_wgt, void* out, const float* scale,
                       const float* shift, int64_t n, int64_t n_alloc, int fmt, int act, int num_sms,
                       cudaStream_t st, int64_t* launches);
int launch_convt_s2_simt(const ConvLayer& L, const void* in, void* out, int64_t n, int fmt, int act,
                         cudaStream_t st, int64_t* launches);
// Final ConvT(64->1, s2) + sigmoid + mean over K + threshold + TP/FP/FN (+ optional mean grid).
int launch_tail(const void* a4, const float* w5, int64_t B, int K, int fmt, int final_sigmoid,
                const uint8_t* target_bits, float thr, unsigned long long* counts, float* mean_prob, float gamma,
                double* loss, cudaStream_t st, int64_t* launches);
int launch_tail_tc(const CUtensorMap& tmap_a4, const CUtensorMap& tmap_w5, int64_t B, int K, int fmt,
                   int final_sigmoid, const uint8_t* target_bits, float thr, unsigned long long* counts,
                   float* mean_prob, float gamma, double* loss, int num_sms, cudaStream_t st, int64_t* launches);
int launch_binary_loss(const float* pred, const float* target, int64_t B, int64_t V, float gamma, double* loss,
                       cudaStream_t st, int64_t* launches);
int launch_counts_sweep(const float* target, const float* pred, int64_t B, int64_t V, const float* thr, int T, int strict,
                        unsigned long long* counts, cudaStream_t st, int64_t* launches);
int launch_impute(const float* z, const float* mask, const float* mu, int C, int64_t B, int K, int D, uint64_t seed,
                  uint64_t obj_offset, int fill, float* z_out, int32_t* cstar, cudaStream_t st, int64_t* launches);
int launch_counts(const float* target, const float* pred, int64_t B, int64_t V, float thr,
                  unsigned long long* counts, cudaStream_t st, int64_t* launches);
int launch_pack(const float* target, int64_t B, int64_t V, uint8_t* bits, cudaStream_t st, int64_t* launches);
int launch_to_f32(const void* src, float* dst, int64_t n, int fmt, cudaStream_t st);

size_t convt_tc_smem_bytes(int cin, int cout, int win);

}  // namespace a3d
