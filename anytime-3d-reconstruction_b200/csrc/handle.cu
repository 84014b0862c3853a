// Host side of liba3d: the C ABI declared in include/a3d.h.  Owns weights (Keras order/layouts), folds BatchNorm,
// repacks the transposed-conv kernels per (output parity, tap, 64-channel chunk) for the tcgen05 kernel, builds the
// TMA tensor maps and drives the per-chunk kernel pipeline on the caller's stream.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <mutex>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "internal.h"

namespace a3d {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

}  // namespace a3d

using namespace a3d;

namespace a3d {

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

bool pdl_enabled() {
  static const bool on = !(getenv("A3D_PDL") && atoi(getenv("A3D_PDL")) == 0);
  return on;
}

uint16_t cvt16(float v, int fmt) {
  if (fmt == A3D_DTYPE_F16) {
    __half h = __float2half_rn(v);
    uint16_t r;
    memcpy(&r, &h, 2);
    return r;
  }
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  uint16_t r;
  memcpy(&r, &h, 2);
  return r;
}

int upload(const void* src, size_t bytes, void** dst) {
  if (!*dst) A3D_CUDA_OK(cudaMalloc(dst, bytes));
  A3D_CUDA_OK(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
  return A3D_OK;
}

void fold_bn(const std::vector<float>& g, const std::vector<float>& b, const std::vector<float>& m,
             const std::vector<float>& v, std::vector<float>& scale, std::vector<float>& shift) {
  const size_t n = g.size();
  scale.resize(n);
  shift.resize(n);
  for (size_t i = 0; i < n; ++i) {
    const float s = g[i] / std::sqrt(v[i] + kBnEps);
    scale[i] = s;
    shift[i] = b[i] - m[i] * s;
  }
}

}  // namespace a3d

struct a3d_handle {
  a3d_desc desc{};
  int num_sms = 0;
  int n_weights = 0;
  std::vector<std::vector<float>> w;   // Keras order, fp32, Keras layouts
  std::vector<int64_t> w_numel;
  bool dirty = true;
  bool all_set = false;
  std::vector<bool> w_set;
  // layer geometry
  int grid0 = 0, ch0 = 0, dense_units = 0;
  // device weights
  float *d_wd = nullptr, *d_bd = nullptr, *d_s0 = nullptr, *d_h0 = nullptr;   // dense + folded BN0
  void* d_w1_tco = nullptr; float *d_s1 = nullptr, *d_h1 = nullptr;          // stride-1 layer
  void* d_mt = nullptr; CUtensorMap tmap_a0, tmap_mt;                         // ... as a dense GEMM (tcgen05 path)
  ConvLayer conv[3];                                                           // stride-2 hidden layers
  float* d_w5 = nullptr;                                                       // final kernel [tap][ci] fp32
  void* d_w5_hcol = nullptr;                                                   // tcgen05 tail (tail_hcol.cu): [Za0 16 | Za1 16 | Zm 16 | Zp 16] rows
  CUtensorMap tmap_a4h, tmap_w5h;                                              // (c,h,d,n,w) view of act[4], 64 x 32 x 4 x 1 x 4 boxes; its W5
  bool tail_simt = false;                                                      // A3D_TAIL_IMPL=simt: CUDA-core tail on the tcgen05 activations
  // arena
  int64_t max_chunk = 0;
  void* act[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // a0..a4
  size_t act_elems[5] = {0, 0, 0, 0, 0};                         // per decode
  size_t arena_bytes = 0;
  int64_t last_chunk_n = 0;
  // host-call staging
  cudaStream_t own_stream = nullptr;
  float *st_z = nullptr, *st_mask = nullptr, *st_mu = nullptr, *st_zout = nullptr, *st_mean = nullptr;
  uint8_t* st_bits = nullptr;
  unsigned long long* st_counts = nullptr;
  int64_t st_B = 0, st_BK = 0, st_C = 0;
  bool st_has_mean = false;
  // a3d_decode_host staging: latents, two fp32 grid buffers, two converted (fp16 / bit) buffers, copy stream, events
  cudaStream_t copy_stream = nullptr;
  float* dh_z = nullptr; int64_t dh_z_cap = 0;
  float* dh_grid[2] = {nullptr, nullptr}; void* dh_out[2] = {nullptr, nullptr}; int64_t dh_sub_cap = 0;
  cudaEvent_t dh_done[2] = {nullptr, nullptr}, dh_free[2] = {nullptr, nullptr};
  // bookkeeping
  int64_t launches = 0;
  bool profiling = false;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  float stage_ms[5] = {0, 0, 0, 0, 0};
  int sticky = 0;
  int* d_progress = nullptr; // pacing counters of the w-sweep kernel (one int per cluster)
  int l4_pace = 8;           // sweep steps a cluster may run ahead of its peers (soft pacing, convt_l4_sw.cu);
                             // A3D_L4_PACE=<n> overrides, 0 switches pacing off
  int conv_variant = 0;      // A3D_CONV_PAIR = 1 / 2 / 3 forces the single-CTA / decode-pair / h-pair variant of the row-unit
                             // kernel (0: chosen per call size)
  int l4_impl = 0;           // 128->64 layer: 0 = w-sweep 2-CTA kernel (convt_l4_sw.cu, default); A3D_L4_IMPL=generic: the
                             // 1-CTA kernel of the other stride-2 layers (cross-check; same values up to summation order)
};

namespace {

int check_handle(const a3d_handle* h) {
  // The launch wrappers test cudaGetLastError() after every kernel launch.  That slot is process-wide and other runtime
  // users (torch; the driver's staging path of a pageable cudaMemcpyAsync) can leave a benign, non-sticky error in it:
  // drop it here so that it is not mistaken for a failure of the first launch of this call.  Sticky errors (an illegal
  // access, a trap) come back on every later call anyway.
  cudaGetLastError();
  if (!h) { set_error("null handle"); return A3D_ERR_INVALID; }
  if (h->sticky) { set_error("handle is in a sticky CUDA error state (%d)", h->sticky); return h->sticky; }
  return A3D_OK;
}

// the stand-alone scoring helpers accept a NULL handle (no sticky state, no launch counting)
int check_opt(const a3d_handle* h) {
  if (h) return check_handle(h);
  cudaGetLastError();
  return A3D_OK;
}
int sticky_opt(a3d_handle* h, int rc) {
  if (h && rc == A3D_ERR_CUDA) h->sticky = rc;
  return rc;
}

// Keras variable table ------------------------------------------------------------------------------------------
void build_weight_table(a3d_handle* h) {
  const a3d_desc& d = h->desc;
  h->w_numel.clear();
  h->w_numel.push_back((int64_t)d.latent_dim * h->dense_units);  // dense/kernel
  h->w_numel.push_back(h->dense_units);                          // dense/bias
  for (int i = 0; i < 4; ++i) h->w_numel.push_back(h->dense_units);
  int cin = h->ch0;
  for (int l = 0; l < d.num_layers; ++l) {
    const int64_t k = d.ksizes[l];
    h->w_numel.push_back(k * k * k * d.filters[l] * cin);
    if (l < d.num_layers - 1)
      for (int i = 0; i < 4; ++i) h->w_numel.push_back(d.filters[l]);
    cin = d.filters[l];
  }
  h->n_weights = (int)h->w_numel.size();
  h->w.assign(h->n_weights, {});
  h->w_set.assign(h->n_weights, false);
}

inline int tap_of(int p, int s) { return 3 - p - 2 * s; }  // tap = p + 1 - 2*delta, delta = s - 1 + p

// Keras kernel [4][4][4][cout][cin]  ->  rows of 64 ci, ordered [par][sd][sh][chunk][block][co]
void pack_tc_weights(const std::vector<float>& wk, int cin, int cout, int fmt, std::vector<uint16_t>& out) {
  const bool pwb = cout <= 128;
  const int npar = pwb ? 4 : 8;
  const int chunks = cin / 64;
  const int nblk = pwb ? 4 : 2;
  out.resize((size_t)npar * 4 * chunks * nblk * cout * 64);
  size_t row = 0;
  for (int par = 0; par < npar; ++par) {
    const int pd = pwb ? (par >> 1) : (par >> 2);
    const int ph = pwb ? (par & 1) : ((par >> 1) & 1);
    const int pw1 = par & 1;
    for (int sd = 0; sd < 2; ++sd)
      for (int sh = 0; sh < 2; ++sh) {
        const int td = tap_of(pd, sd), th = tap_of(ph, sh);
        for (int c = 0; c < chunks; ++c)
          for (int blk = 0; blk < nblk; ++blk) {
            int tw;
            if (pwb) {
              // blk 0: pw0 dw=0 (tw 1); 1: pw1 dw=0 (tw 2); 2: pw0 dw=-1 (tw 3); 3: pw1 dw=+1 (tw 0)
              static const int tws[4] = {1, 2, 3, 0};
              tw = tws[blk];
            } else {
              tw = (blk == 0) ? (pw1 + 1) : (pw1 ? 0 : 3);
            }
            const size_t tap = ((size_t)td * 4 + th) * 4 + tw;
            for (int co = 0; co < cout; ++co, ++row) {
              const float* src = &wk[(tap * cout + co) * cin + (size_t)c * 64];
              uint16_t* dst = &out[row * 64];
              for (int i = 0; i < 64; ++i) dst[i] = cvt16(src[i], fmt);
            }
          }
      }
  }
}

// w-sweep 2-CTA layout of the 128->64 layer (convt_l4_sw.cu): [class q = pd*2+ph][rank][sd][sh][chunk][128 rows];
// rank 0 holds the w taps 0 and 1, rank 1 the taps 2 and 3 (64 output channels each): the N = 256 B operand of a K step
// is [tap0 | tap1 | tap2 | tap3], i.e. the accumulator blocks (j-1, pw 1), (j, pw 0), (j, pw 1), (j+1, pw 0).
void pack_sw_weights(const std::vector<float>& wk, int cin, int cout, int fmt, std::vector<uint16_t>& out) {
  const int chunks = cin / 64;
  out.resize((size_t)4 * 2 * 4 * chunks * 128 * 64);
  size_t row = 0;
  for (int q = 0; q < 4; ++q) {
    const int pd = q >> 1, ph = q & 1;
    for (int rank = 0; rank < 2; ++rank)
      for (int sd = 0; sd < 2; ++sd)
        for (int sh = 0; sh < 2; ++sh) {
          const int td = tap_of(pd, sd), th = tap_of(ph, sh);
          for (int c = 0; c < chunks; ++c)
            for (int r = 0; r < 128; ++r, ++row) {
              const int tw = 2 * rank + (r >> 6), co = r & 63;
              const size_t tap = ((size_t)td * 4 + th) * 4 + tw;
              const float* src = &wk[(tap * cout + co) * cin + (size_t)c * 64];
              uint16_t* dst = &out[row * 64];
              for (int i = 0; i < 64; ++i) dst[i] = cvt16(src[i], fmt);
            }
        }
  }
}

void pack_tco_weights(const std::vector<float>& wk, int ntap, int cin, int cout, int fmt, std::vector<uint16_t>& out) {
  out.resize((size_t)ntap * cin * cout);
  for (int t = 0; t < ntap; ++t)
    for (int co = 0; co < cout; ++co)
      for (int ci = 0; ci < cin; ++ci)
        out[((size_t)t * cin + ci) * cout + co] = cvt16(wk[((size_t)t * cout + co) * cin + ci], fmt);
}

int make_tmaps(a3d_handle* h, int li) {
  ConvLayer& L = h->conv[li];
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return A3D_ERR_CUDA; }
  const CUtensorMapDataType dt =
      h->desc.operand_dtype == A3D_DTYPE_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const uint64_t W = L.win, C = L.cin;
  {
    cuuint64_t dims[5] = {C, (cuuint64_t)h->max_chunk, W, W, W};  // (c, n, w, h, d)
    cuuint64_t strides[4] = {W * W * W * C * 2, C * 2, W * C * 2, W * W * C * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)(128 / L.win), (cuuint32_t)(L.win + 2), 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&L.tmap_act, dt, 5, h->act[li + 1], dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(activations, layer %d) failed: %d", li, (int)r); return A3D_ERR_CUDA; }
  }
  {
    const bool pwb = L.cout <= 128;
    const uint64_t rows = (uint64_t)(pwb ? 4 : 8) * 4 * (L.cin / 64) * (pwb ? 4 : 2) * L.cout;
    cuuint64_t dims[2] = {64, rows};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, 256};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&L.tmap_wgt, dt, 2, L.wgt_packed, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(weights, layer %d) failed: %d", li, (int)r); return A3D_ERR_CUDA; }
    cuuint32_t box64[2] = {64, 64};
    r = enc(&L.tmap_wgt64, dt, 2, L.wgt_packed, dims, strides, box64, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(weights64, layer %d) failed: %d", li, (int)r); return A3D_ERR_CUDA; }
  }
  return A3D_OK;
}

int finalize_weights(a3d_handle* h) {
  if (!h->dirty) return A3D_OK;
  for (int i = 0; i < h->n_weights; ++i)
    if (!h->w_set[i]) { set_error("weight %d of %d was never set", i, h->n_weights); return A3D_ERR_WEIGHTS; }
  const int fmt = h->desc.operand_dtype;
  int rc;
  std::vector<float> sc, sf;
  // the uploads below overwrite live device buffers with synchronous copies: no earlier decode (on any caller stream) may
  // still be reading them
  if (h->d_wd) A3D_CUDA_OK(cudaDeviceSynchronize());
  // dense + BN0
  if ((rc = upload(h->w[0].data(), h->w[0].size() * 4, (void**)&h->d_wd))) return rc;
  if ((rc = upload(h->w[1].data(), h->w[1].size() * 4, (void**)&h->d_bd))) return rc;
  fold_bn(h->w[2], h->w[3], h->w[4], h->w[5], sc, sf);
  if ((rc = upload(sc.data(), sc.size() * 4, (void**)&h->d_s0))) return rc;
  if ((rc = upload(sf.data(), sf.size() * 4, (void**)&h->d_h0))) return rc;
  // stride-1 layer
  std::vector<uint16_t> p16;
  pack_tco_weights(h->w[6], 64, h->ch0, h->desc.filters[0], fmt, p16);
  if ((rc = upload(p16.data(), p16.size() * 2, &h->d_w1_tco))) return rc;
  fold_bn(h->w[7], h->w[8], h->w[9], h->w[10], sc, sf);
  if ((rc = upload(sc.data(), sc.size() * 4, (void**)&h->d_s1))) return rc;
  if ((rc = upload(sf.data(), sf.size() * 4, (void**)&h->d_h1))) return rc;
  if (h->desc.impl == A3D_IMPL_TCGEN05) {
    // dense [512 x 32768] matrix of the stride-1 layer: Mt[(o, co)][(i, ci)] = W[t = o - i + 1][co][ci]
    const int cout = h->desc.filters[0], cin = h->ch0;
    p16.assign((size_t)64 * cout * 512, cvt16(0.f, fmt));
    for (int o = 0; o < 64; ++o)
      for (int i = 0; i < 64; ++i) {
        const int td = (o >> 4) - (i >> 4) + 1, th = ((o >> 2) & 3) - ((i >> 2) & 3) + 1, tw = (o & 3) - (i & 3) + 1;
        if (td < 0 || td > 3 || th < 0 || th > 3 || tw < 0 || tw > 3) continue;
        const float* src = &h->w[6][(size_t)((td * 4 + th) * 4 + tw) * cout * cin];
        for (int co = 0; co < cout; ++co)
          for (int ci = 0; ci < cin; ++ci)
            p16[((size_t)o * cout + co) * 512 + (size_t)i * cin + ci] = cvt16(src[(size_t)co * cin + ci], fmt);
      }
    if ((rc = upload(p16.data(), p16.size() * 2, &h->d_mt))) return rc;
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return A3D_ERR_CUDA; }
    const CUtensorMapDataType dt = fmt == A3D_DTYPE_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    cuuint32_t es[2] = {1, 1};
    cuuint64_t st1[1] = {1024};
    cuuint64_t da[2] = {512, (cuuint64_t)h->max_chunk};
    cuuint32_t ba[2] = {64, 128};
    CUresult r = enc(&h->tmap_a0, dt, 2, h->act[0], da, st1, ba, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(a0) failed: %d", (int)r); return A3D_ERR_CUDA; }
    cuuint64_t dm[2] = {512, (cuuint64_t)64 * cout};
    cuuint32_t bm[2] = {64, 256};
    r = enc(&h->tmap_mt, dt, 2, h->d_mt, dm, st1, bm, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(Mt) failed: %d", (int)r); return A3D_ERR_CUDA; }
  }
  // stride-2 hidden layers
  for (int li = 0; li < 3; ++li) {
    ConvLayer& L = h->conv[li];
    const int base = 11 + 5 * li;
    pack_tc_weights(h->w[base], L.cin, L.cout, fmt, p16);
    if ((rc = upload(p16.data(), p16.size() * 2, &L.wgt_packed))) return rc;
    if (h->desc.impl == A3D_IMPL_SIMT) {
      pack_tco_weights(h->w[base], 64, L.cin, L.cout, fmt, p16);
      if ((rc = upload(p16.data(), p16.size() * 2, &L.wgt_tco))) return rc;
    }
    fold_bn(h->w[base + 1], h->w[base + 2], h->w[base + 3], h->w[base + 4], sc, sf);
    if ((rc = upload(sc.data(), sc.size() * 4, (void**)&L.scale))) return rc;
    if ((rc = upload(sf.data(), sf.size() * 4, (void**)&L.shift))) return rc;
    if ((rc = make_tmaps(h, li))) return rc;
    if (li == 2) {
      EncodeTiledFn enc = get_encode_fn();
      const CUtensorMapDataType dt = fmt == A3D_DTYPE_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
      cuuint64_t strides[1] = {128};
      cuuint32_t box[2] = {64, 256};
      cuuint32_t es[2] = {1, 1};
      pack_sw_weights(h->w[base], L.cin, L.cout, fmt, p16);
      if ((rc = upload(p16.data(), p16.size() * 2, &L.wgt_sw))) return rc;
      cuuint64_t dims[2] = {64, (cuuint64_t)(p16.size() / 64)};
      CUresult r = enc(&L.tmap_wgt_sw, dt, 2, L.wgt_sw, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(sw weights) failed: %d", (int)r); return A3D_ERR_CUDA; }
      const uint64_t W = L.win, C = L.cin;
      cuuint64_t ad[5] = {C, (cuuint64_t)h->max_chunk, W, W, W};   // (c, n, h, w, d)
      cuuint64_t as[4] = {W * W * W * C * 2, W * C * 2, C * 2, W * W * C * 2};
      cuuint32_t ab[5] = {64, 8, (cuuint32_t)(L.win + 2), 1, 1};
      cuuint32_t ae[5] = {1, 1, 1, 1, 1};
      r = enc(&L.tmap_act_sw, dt, 5, h->act[li + 1], ad, as, ab, ae, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(sw activations) failed: %d", (int)r); return A3D_ERR_CUDA; }
    }
  }
  // final kernel [4,4,4,1,64] is already [tap][ci]
  if ((rc = upload(h->w[26].data(), h->w[26].size() * 4, (void**)&h->d_w5))) return rc;
  {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return A3D_ERR_CUDA; }
    const CUtensorMapDataType dt = fmt == A3D_DTYPE_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    cuuint64_t ws[1] = {128};
    CUresult r;
    // B operand of the tail GEMM (tail_hcol.cu): rows n = blk * 16 + td * 4 + j with th = (j + 1) & 3 (h taps in the order 1, 2, 3, 0):
    // blk 0 = tap_w 1 (pw = 0, delta_w = 0), blk 1 = tap_w 2 (pw = 1, delta_w = 0), blk 2 = tap_w 3, blk 3 = tap_w 0
    p16.assign((size_t)64 * 64, cvt16(0.f, fmt));
    for (int td = 0; td < 4; ++td)
      for (int j = 0; j < 4; ++j) {
        const int n = td * 4 + j, q = td * 4 + ((j + 1) & 3);
        for (int ci = 0; ci < 64; ++ci) {
          p16[(size_t)(0 + n) * 64 + ci] = cvt16(h->w[26][(size_t)(q * 4 + 1) * 64 + ci], fmt);
          p16[(size_t)(16 + n) * 64 + ci] = cvt16(h->w[26][(size_t)(q * 4 + 2) * 64 + ci], fmt);
          p16[(size_t)(32 + n) * 64 + ci] = cvt16(h->w[26][(size_t)(q * 4 + 3) * 64 + ci], fmt);
          p16[(size_t)(48 + n) * 64 + ci] = cvt16(h->w[26][(size_t)(q * 4 + 0) * 64 + ci], fmt);
        }
      }
    if ((rc = upload(p16.data(), p16.size() * 2, &h->d_w5_hcol))) return rc;
    cuuint64_t pd[5] = {64, 32, 32, (cuuint64_t)h->max_chunk, 32};   // (c, h, d, n, w): GEMM rows come out (w, d, h)
    cuuint64_t ps[4] = {128 * 32, 128 * 32 * 32, 128ull * 32 * 32 * 32, 128};
    cuuint32_t hb[5] = {64, 32, 4, 1, 4};   // rows (w, d, h) of one sample, the whole h axis per box
    r = enc(&h->tmap_a4h, dt, 5, h->act[4], pd, ps, hb, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(hcol-tail activations) failed: %d", (int)r); return A3D_ERR_CUDA; }
    cuuint64_t wpd[2] = {64, 64};
    cuuint32_t wpb[2] = {64, 64};
    r = enc(&h->tmap_w5h, dt, 2, h->d_w5_hcol, wpd, ws, wpb, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(tail weights) failed: %d", (int)r); return A3D_ERR_CUDA; }
  }
  h->dirty = false;
  return A3D_OK;
}

int sticky(a3d_handle* h, int rc) {
  if (rc == A3D_ERR_CUDA) h->sticky = rc;
  return rc;
}

// Run Dense .. L4 for n (<= max_chunk) latents; leaves the 32^3 x 64 activations in act[4].
int run_hidden(a3d_handle* h, const float* z_dev, int64_t n, cudaStream_t st) {
  const int fmt = h->desc.operand_dtype, act = h->desc.activation;
  int rc;
  if (h->profiling) cudaEventRecord(h->ev[0], st);
  const bool simt_l1 = h->desc.impl == A3D_IMPL_SIMT;
  rc = launch_dense_l1(z_dev, n, h->desc.latent_dim, h->d_wd, h->d_bd, h->d_s0, h->d_h0, h->act[0], h->d_w1_tco,
                       h->d_s1, h->d_h1, h->act[1], fmt, act, simt_l1, st, &h->launches);
  if (rc) return rc;
  if (!simt_l1) {
    rc = launch_gemm_l1(h->tmap_a0, h->tmap_mt, h->act[1], h->d_s1, h->d_h1, n, h->max_chunk, fmt, act, h->num_sms, st,
                        &h->launches);
    if (rc) return rc;
  }
  if (h->profiling) cudaEventRecord(h->ev[1], st);
  for (int li = 0; li < 3; ++li) {
    if (h->desc.impl == A3D_IMPL_SIMT)
      rc = launch_convt_s2_simt(h->conv[li], h->act[li + 1], h->act[li + 2], n, fmt, act, st, &h->launches);
    else if (li == 2 && h->l4_impl == 0)
      rc = launch_convt_l4_sw(h->conv[li].tmap_act_sw, h->conv[li].tmap_wgt_sw, h->act[li + 2], h->conv[li].scale,
                              h->conv[li].shift, n, h->max_chunk, fmt, act, h->num_sms, h->d_progress, h->l4_pace, st, &h->launches);
    else
      rc = launch_convt_s2_tc(h->conv[li], h->act[li + 2], n, h->max_chunk, fmt, act, h->num_sms, h->conv_variant, st,
                              &h->launches);
    if (rc) return rc;
    if (h->profiling) cudaEventRecord(h->ev[2 + li], st);
  }
  h->last_chunk_n = n;
  return A3D_OK;
}

int run_tail(a3d_handle* h, int64_t B, int K, const uint8_t* bits, float thr, unsigned long long* counts, float* mean,
             float gamma, double* loss, cudaStream_t st) {
  const int sig = h->desc.final_activation == A3D_FINAL_SIGMOID;
  // A3D_TAIL_IMPL=simt: the CUDA-core tail on the activations of the tcgen05 layers (on-device cross-check of the
  // tensor-core tail in isolation, tests/test_gpu_parity.py)
  if (h->desc.impl == A3D_IMPL_SIMT || h->tail_simt)
    return launch_tail(h->act[4], h->d_w5, B, K, h->desc.operand_dtype, sig, bits, thr, counts, mean, gamma, loss, st,
                       &h->launches);
  return launch_tail_hcol(h->tmap_a4h, h->tmap_w5h, B, K, h->desc.operand_dtype, sig, bits, thr, counts, mean, gamma,
                          loss, h->num_sms, st, &h->launches);
}

void collect_profile(a3d_handle* h, cudaStream_t st, bool first) {
  if (!h->profiling) return;
  cudaEventRecord(h->ev[5], st);
  cudaEventSynchronize(h->ev[5]);
  for (int i = 0; i < 5; ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]);
    h->stage_ms[i] = (first ? 0.f : h->stage_ms[i]) + ms;
  }
}

}  // namespace

extern "C" {

int a3d_abi_version(void) { return A3D_ABI_VERSION; }

uint32_t a3d_crc32c(const void* data, size_t n, uint32_t crc) {
  static uint32_t tbl[8][256];
  static std::once_flag once;     // loader workers may verify checkpoints concurrently
  std::call_once(once, [] {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      tbl[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int t = 1; t < 8; ++t) tbl[t][i] = (tbl[t - 1][i] >> 8) ^ tbl[0][tbl[t - 1][i] & 0xFF];
  });
  const uint8_t* p = static_cast<const uint8_t*>(data);
  uint32_t c = crc ^ 0xFFFFFFFFu;
  while (n >= 8) {   // slice-by-8
    uint32_t lo, hi;
    memcpy(&lo, p, 4);
    memcpy(&hi, p + 4, 4);
    lo ^= c;
    c = tbl[7][lo & 0xFF] ^ tbl[6][(lo >> 8) & 0xFF] ^ tbl[5][(lo >> 16) & 0xFF] ^ tbl[4][lo >> 24] ^
        tbl[3][hi & 0xFF] ^ tbl[2][(hi >> 8) & 0xFF] ^ tbl[1][(hi >> 16) & 0xFF] ^ tbl[0][hi >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) c = tbl[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}
const char* a3d_last_error(void) { return g_err; }

#ifdef A3D_CHECKED
namespace { __global__ void checked_selftest_kernel(int v) { A3D_DEV_CHECK(v < 0); } }
#endif

int a3d_create(const a3d_desc* d, a3d_handle** out) {
  if (!d || !out) { set_error("null argument"); return A3D_ERR_INVALID; }
  *out = nullptr;
#ifdef A3D_CHECKED
  // proof that the range checks of this build are live: A3D_CHECK_SELFTEST=1 runs a kernel whose check must fail
  if (getenv("A3D_CHECK_SELFTEST") && atoi(getenv("A3D_CHECK_SELFTEST")) == 1) {
    checked_selftest_kernel<<<1, 1>>>(1);
    A3D_CUDA_OK(cudaDeviceSynchronize());
    set_error("checked build: the self-test check did not trap");
    return A3D_ERR_INVALID;
  }
#endif
  if (d->abi_version != A3D_ABI_VERSION) { set_error("ABI version mismatch: %d vs %d", d->abi_version, A3D_ABI_VERSION); return A3D_ERR_INVALID; }
  // supported structure: the decoder every reference script builds (autoencoder3D.py:15-24; test_*_VAE*.py configs)
  static const int kF[5] = {512, 256, 128, 64, 1}, kS[5] = {1, 2, 2, 2, 2};
  bool ok = d->num_layers == 5 && d->out_grid == 64 && d->latent_dim >= 1 && d->latent_dim <= 4096;
  for (int i = 0; ok && i < 5; ++i) ok = d->filters[i] == kF[i] && d->strides[i] == kS[i] && d->ksizes[i] == 4;
  if (!ok) {
    set_error("unsupported decoder structure: this build implements filter_num_list [512,256,128,64,1], "
              "filter_size_list [4]*5, strides_list [1,2,2,2,2], output_shape [64,64,64,1], any input_dim");
    return A3D_ERR_INVALID;
  }
  if (d->activation < 0 || d->activation > 3 || d->final_activation < 0 || d->final_activation > 1 ||
      (d->operand_dtype != A3D_DTYPE_F16 && d->operand_dtype != A3D_DTYPE_BF16) ||
      (d->impl != A3D_IMPL_TCGEN05 && d->impl != A3D_IMPL_SIMT)) {
    set_error("invalid activation / dtype / impl field");
    return A3D_ERR_INVALID;
  }
  if (d->max_chunk < 32 || d->max_chunk % 32 != 0) { set_error("max_chunk must be a positive multiple of 32"); return A3D_ERR_INVALID; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || d->device >= ndev) {
    cudaGetLastError();
    set_error("no CUDA device %d available; liba3d has no CPU path", d->device);
    return A3D_ERR_NO_DEVICE;
  }
  cudaDeviceProp prop;
  A3D_CUDA_OK(cudaGetDeviceProperties(&prop, d->device));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; liba3d is built for sm_100a only", d->device, prop.major, prop.minor);
    return A3D_ERR_NO_DEVICE;
  }
  A3D_CUDA_OK(cudaSetDevice(d->device));
  a3d_handle* h = new a3d_handle();
  h->desc = *d;
  h->num_sms = prop.multiProcessorCount;
  h->grid0 = 4;       // output_shape[:-1] / prod(strides)            autoencoder3D.py:115
  h->ch0 = 8;         // max(filter_num_list[0] / 64, 8)              autoencoder3D.py:116-118
  h->dense_units = h->grid0 * h->grid0 * h->grid0 * h->ch0;  // :120
  build_weight_table(h);
  h->max_chunk = d->max_chunk;
  { const char* e = getenv("A3D_L4_IMPL"); h->l4_impl = (e && std::string(e) == "generic") ? 2 : 0; }
  { const char* e = getenv("A3D_L4_PACE"); if (e) h->l4_pace = atoi(e); }
  { const char* e = getenv("A3D_CONV_PAIR"); if (e) h->conv_variant = atoi(e); }
  if (cudaMalloc(&h->d_progress, convt_l4_sw_progress_bytes()) != cudaSuccess) {
    set_error("allocation of the pacing counters failed");
    a3d_destroy(h);
    return A3D_ERR_CUDA;
  }
  { const char* e = getenv("A3D_TAIL_IMPL"); h->tail_simt = e && std::string(e) == "simt"; }
  const int geo[3][3] = {{512, 256, 4}, {256, 128, 8}, {128, 64, 16}};
  for (int i = 0; i < 3; ++i) { h->conv[i].cin = geo[i][0]; h->conv[i].cout = geo[i][1]; h->conv[i].win = geo[i][2]; }
  h->act_elems[0] = 512; h->act_elems[1] = 64 * 512; h->act_elems[2] = 512 * 256; h->act_elems[3] = 4096 * 128;
  h->act_elems[4] = 32768 * 64;
  for (int i = 0; i < 5; ++i) {
    const size_t bytes = h->act_elems[i] * 2 * (size_t)h->max_chunk;
    cudaError_t e = cudaMalloc(&h->act[i], bytes);
    if (e != cudaSuccess) {
      set_error("arena allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
      a3d_destroy(h);
      return A3D_ERR_CUDA;
    }
    cudaMemset(h->act[i], 0, bytes);
    h->arena_bytes += bytes;
  }
  cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
  for (int i = 0; i < 6; ++i) cudaEventCreate(&h->ev[i]);
  *out = h;
  return A3D_OK;
}

void a3d_destroy(a3d_handle* h) {
  if (!h) return;
  cudaDeviceSynchronize();
  for (int i = 0; i < 5; ++i) cudaFree(h->act[i]);
  cudaFree(h->d_wd); cudaFree(h->d_bd); cudaFree(h->d_s0); cudaFree(h->d_h0);
  cudaFree(h->d_mt); cudaFree(h->d_w1_tco); cudaFree(h->d_s1); cudaFree(h->d_h1); cudaFree(h->d_w5); cudaFree(h->d_w5_hcol);
  for (auto& L : h->conv) { cudaFree(L.wgt_packed); cudaFree(L.wgt_sw); cudaFree(L.wgt_tco); cudaFree(L.scale); cudaFree(L.shift); }
  cudaFree(h->st_z); cudaFree(h->st_mask); cudaFree(h->st_mu); cudaFree(h->st_zout); cudaFree(h->st_mean);
  cudaFree(h->st_bits); cudaFree(h->st_counts);
  cudaFree(h->d_progress);
  cudaFree(h->dh_z);
  for (int i = 0; i < 2; ++i) {
    cudaFree(h->dh_grid[i]); cudaFree(h->dh_out[i]);
    if (h->dh_done[i]) cudaEventDestroy(h->dh_done[i]);
    if (h->dh_free[i]) cudaEventDestroy(h->dh_free[i]);
  }
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  for (int i = 0; i < 6; ++i) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  delete h;
}

int a3d_num_weights(const a3d_handle* h) { return h ? h->n_weights : 0; }
int64_t a3d_weight_numel(const a3d_handle* h, int index) {
  return (h && index >= 0 && index < h->n_weights) ? h->w_numel[index] : -1;
}

int a3d_set_weight(a3d_handle* h, int index, const float* host, size_t nbytes) {
  int rc = check_handle(h);
  if (rc) return rc;
  if (index < 0 || index >= h->n_weights || !host) { set_error("bad weight index %d", index); return A3D_ERR_INVALID; }
  if (nbytes != (size_t)h->w_numel[index] * 4) {
    set_error("weight %d: expected %lld fp32 values, got %zu bytes", index, (long long)h->w_numel[index], nbytes);
    return A3D_ERR_WEIGHTS;
  }
  h->w[index].assign(host, host + h->w_numel[index]);
  h->w_set[index] = true;
  h->dirty = true;
  return A3D_OK;
}

int a3d_get_weight(const a3d_handle* h, int index, float* host, size_t nbytes) {
  if (!h || index < 0 || index >= h->n_weights || !host) { set_error("bad weight index %d", index); return A3D_ERR_INVALID; }
  if (!h->w_set[index]) { set_error("weight %d was never set", index); return A3D_ERR_WEIGHTS; }
  if (nbytes != (size_t)h->w_numel[index] * 4) { set_error("weight %d: size mismatch", index); return A3D_ERR_WEIGHTS; }
  memcpy(host, h->w[index].data(), nbytes);
  return A3D_OK;
}

size_t a3d_workspace_bytes(const a3d_handle* h, int64_t n) {
  if (!h) return 0;
  if (n <= 0) return h->arena_bytes;
  size_t per = 0;
  for (int i = 0; i < 5; ++i) per += h->act_elems[i] * 2;
  return per * (size_t)((n + 31) / 32 * 32);
}

int64_t a3d_launch_count(const a3d_handle* h) { return h ? h->launches : 0; }
int a3d_set_profiling(a3d_handle* h, int enable) { if (!h) return A3D_ERR_INVALID; h->profiling = enable != 0; return A3D_OK; }
int a3d_stage_times_ms(a3d_handle* h, float* out, int max_stages) {
  if (!h || !out) return 0;
  int n = max_stages < 5 ? max_stages : 5;
  for (int i = 0; i < n; ++i) out[i] = h->stage_ms[i];
  return n;
}

int a3d_decode(a3d_handle* h, const float* z_dev, int64_t n, float* prob_dev, void* stream) {
  int rc = check_handle(h);
  if (rc) return rc;
  if (n < 0 || (n > 0 && (!z_dev || !prob_dev))) { set_error("a3d_decode: bad arguments"); return A3D_ERR_INVALID; }
  if ((rc = sticky(h, finalize_weights(h)))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int D = h->desc.latent_dim;
  for (int64_t off = 0; off < n; off += h->max_chunk) {
    const int64_t nc = (n - off < h->max_chunk) ? n - off : h->max_chunk;
    if ((rc = sticky(h, run_hidden(h, z_dev + off * D, nc, st)))) return rc;
    rc = run_tail(h, nc, 1, nullptr, 0.5f, nullptr, prob_dev + off * (int64_t)A3D_VOXELS, 0.f, nullptr, st);
    if ((rc = sticky(h, rc))) return rc;
    collect_profile(h, st, off == 0);
  }
  return A3D_OK;
}

int a3d_impute(a3d_handle* h, const float* z_dev, const float* mask_dev, const float* mu_table_dev, int C, int64_t B,
               int K, uint64_t seed, uint64_t obj_offset, int fill_mode, float* z_out_dev, int32_t* cstar_out_dev,
               void* stream) {
  int rc = check_handle(h);
  if (rc) return rc;
  if (B < 0 || K < 1 || fill_mode < 0 || fill_mode > 3 || (B > 0 && (!z_dev || !mask_dev || !z_out_dev)) ||
      (fill_mode != A3D_FILL_NORMAL && fill_mode != A3D_FILL_NONE && (C < 1 || !mu_table_dev))) {
    set_error("a3d_impute: bad arguments");
    return A3D_ERR_INVALID;
  }
  return sticky(h, launch_impute(z_dev, mask_dev, mu_table_dev, C, B, K, h->desc.latent_dim, seed, obj_offset, fill_mode,
                                 z_out_dev, cstar_out_dev, (cudaStream_t)stream, &h->launches));
}

static int anytime_eval_impl(a3d_handle* h, const float* z_bkd_dev, int64_t B, int K, const uint8_t* target_bits_dev,
                             float thr, float gamma, int64_t* counts_dev, double* loss_dev, float* mean_prob_dev,
                             void* stream) {
  int rc = check_handle(h);
  if (rc) return rc;
  if (B < 0 || K < 1 || (B > 0 && !z_bkd_dev) || (target_bits_dev && !counts_dev) || (loss_dev && !target_bits_dev)) {
    set_error("a3d_anytime_eval: bad arguments");
    return A3D_ERR_INVALID;
  }
  if (K > h->max_chunk) {
    set_error("K = %d exceeds the arena (max_chunk = %lld decodes); create the handle with a larger max_chunk", K,
              (long long)h->max_chunk);
    return A3D_ERR_WORKSPACE;
  }
  if ((rc = sticky(h, finalize_weights(h)))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int D = h->desc.latent_dim;
  if (counts_dev && B > 0) A3D_CUDA_OK(cudaMemsetAsync(counts_dev, 0, (size_t)B * 3 * sizeof(int64_t), st));
  if (loss_dev && B > 0) A3D_CUDA_OK(cudaMemsetAsync(loss_dev, 0, (size_t)B * sizeof(double), st));
  const int64_t obj_per_chunk = h->max_chunk / K;
  for (int64_t b0 = 0; b0 < B; b0 += obj_per_chunk) {
    const int64_t nb = (B - b0 < obj_per_chunk) ? B - b0 : obj_per_chunk;
    if ((rc = sticky(h, run_hidden(h, z_bkd_dev + b0 * K * D, nb * K, st)))) return rc;
    rc = run_tail(h, nb, K, target_bits_dev ? target_bits_dev + b0 * (A3D_VOXELS / 8) : nullptr, thr,
                  counts_dev ? reinterpret_cast<unsigned long long*>(counts_dev) + b0 * 3 : nullptr,
                  mean_prob_dev ? mean_prob_dev + b0 * (int64_t)A3D_VOXELS : nullptr, gamma,
                  loss_dev ? loss_dev + b0 : nullptr, st);
    if ((rc = sticky(h, rc))) return rc;
    collect_profile(h, st, b0 == 0);
  }
  return A3D_OK;
}

int a3d_anytime_eval(a3d_handle* h, const float* z_bkd_dev, int64_t B, int K, const uint8_t* target_bits_dev, float thr,
                     int64_t* counts_dev, float* mean_prob_dev, void* stream) {
  return anytime_eval_impl(h, z_bkd_dev, B, K, target_bits_dev, thr, 0.f, counts_dev, nullptr, mean_prob_dev, stream);
}

int a3d_anytime_eval_loss(a3d_handle* h, const float* z_bkd_dev, int64_t B, int K, const uint8_t* target_bits_dev,
                          float thr, float gamma, int64_t* counts_dev, double* loss_dev, float* mean_prob_dev,
                          void* stream) {
  return anytime_eval_impl(h, z_bkd_dev, B, K, target_bits_dev, thr, gamma, counts_dev, loss_dev, mean_prob_dev, stream);
}

int a3d_binary_loss(a3d_handle* h, const float* pred_dev, const float* target_dev, int64_t B, int64_t V, float gamma,
                    double* loss_dev, void* stream) {
  int rc = check_opt(h);
  if (rc) return rc;
  if (B < 0 || V <= 0 || V % 4 != 0 || (B > 0 && (!pred_dev || !target_dev || !loss_dev))) {
    set_error("a3d_binary_loss: bad arguments (V must be a positive multiple of 4)");
    return A3D_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (B > 0) A3D_CUDA_OK(cudaMemsetAsync(loss_dev, 0, (size_t)B * sizeof(double), st));
  return sticky_opt(h, launch_binary_loss(pred_dev, target_dev, B, V, gamma, loss_dev, st, h ? &h->launches : nullptr));
}

int a3d_counts_sweep(a3d_handle* h, const float* target_dev, const float* pred_dev, int64_t B, int64_t V,
                     const float* thresholds, int T, int strict, int64_t* counts_dev, void* stream) {
  int rc = check_opt(h);
  if (rc) return rc;
  if (B < 0 || V <= 0 || V % 4 != 0 || T < 1 || T > 32 || !thresholds ||
      (B > 0 && (!target_dev || !pred_dev || !counts_dev))) {
    set_error("a3d_counts_sweep: bad arguments (1 <= T <= 32, V a positive multiple of 4)");
    return A3D_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (B > 0) A3D_CUDA_OK(cudaMemsetAsync(counts_dev, 0, (size_t)B * T * 3 * sizeof(int64_t), st));
  return sticky_opt(h, launch_counts_sweep(target_dev, pred_dev, B, V, thresholds, T, strict,
                                       reinterpret_cast<unsigned long long*>(counts_dev), st, h ? &h->launches : nullptr));
}

int a3d_counts(a3d_handle* h, const float* target_dev, const float* pred_dev, int64_t B, int64_t V, float thr,
               int64_t* counts_dev, void* stream) {
  int rc = check_opt(h);
  if (rc) return rc;
  if (B < 0 || V <= 0 || V % 8 != 0 || (B > 0 && (!target_dev || !pred_dev || !counts_dev))) {
    set_error("a3d_counts: bad arguments (V must be a positive multiple of 8)");
    return A3D_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (B > 0) A3D_CUDA_OK(cudaMemsetAsync(counts_dev, 0, (size_t)B * 3 * sizeof(int64_t), st));
  return sticky_opt(h, launch_counts(target_dev, pred_dev, B, V, thr, reinterpret_cast<unsigned long long*>(counts_dev), st,
                                 h ? &h->launches : nullptr));
}

int a3d_pack_targets(a3d_handle* h, const float* target_dev, int64_t B, int64_t V, uint8_t* bits_dev, void* stream) {
  int rc = check_opt(h);
  if (rc) return rc;
  if (B < 0 || V <= 0 || V % 8 != 0 || (B > 0 && (!target_dev || !bits_dev))) {
    set_error("a3d_pack_targets: bad arguments");
    return A3D_ERR_INVALID;
  }
  return sticky_opt(h, launch_pack(target_dev, B, V, bits_dev, (cudaStream_t)stream, h ? &h->launches : nullptr));
}

int a3d_anytime_eval_host(a3d_handle* h, const float* z, const float* mask, const float* mu_table, int C, int64_t B,
                          int K, uint64_t seed, uint64_t obj_offset, int fill_mode, const uint8_t* target_bits, float thr,
                          int64_t* counts, float* mean_prob_or_null) {
  int rc = check_handle(h);
  if (rc) return rc;
  if (B <= 0 || K < 1 || !z || !mask || !target_bits || !counts || (fill_mode != A3D_FILL_NORMAL && (!mu_table || C < 1))) {
    set_error("a3d_anytime_eval_host: bad arguments");
    return A3D_ERR_INVALID;
  }
  const int D = h->desc.latent_dim;
  // (re)size the staging buffers; steady-state calls with the same shapes allocate nothing
  if (B > h->st_B || B * K > h->st_BK || C > h->st_C || (mean_prob_or_null && !h->st_has_mean)) {
    cudaFree(h->st_z); cudaFree(h->st_mask); cudaFree(h->st_mu); cudaFree(h->st_zout); cudaFree(h->st_bits);
    cudaFree(h->st_counts); cudaFree(h->st_mean);
    h->st_z = h->st_mask = h->st_mu = h->st_zout = h->st_mean = nullptr; h->st_bits = nullptr; h->st_counts = nullptr;
    // the new capacities are committed only after every allocation succeeded: a failed cudaMalloc leaves the sizes at
    // zero, so the next call allocates again instead of running on null staging pointers
    const int64_t nB = B > h->st_B ? B : h->st_B, nBK = B * K > h->st_BK ? B * K : h->st_BK;
    const int64_t nC = C > h->st_C ? C : (h->st_C > 0 ? h->st_C : 1);
    const bool want_mean = h->st_has_mean || mean_prob_or_null != nullptr;
    h->st_B = h->st_BK = h->st_C = 0;
    h->st_has_mean = false;
    A3D_CUDA_OK(cudaMalloc(&h->st_z, (size_t)nB * D * 4));
    A3D_CUDA_OK(cudaMalloc(&h->st_mask, (size_t)nB * D * 4));
    A3D_CUDA_OK(cudaMalloc(&h->st_mu, (size_t)nC * D * 4));
    A3D_CUDA_OK(cudaMalloc(&h->st_zout, (size_t)nBK * D * 4));
    A3D_CUDA_OK(cudaMalloc(&h->st_bits, (size_t)nB * (A3D_VOXELS / 8)));
    A3D_CUDA_OK(cudaMalloc(&h->st_counts, (size_t)nB * 3 * 8));
    if (want_mean) A3D_CUDA_OK(cudaMalloc(&h->st_mean, (size_t)nB * A3D_VOXELS * 4));
    h->st_B = nB; h->st_BK = nBK; h->st_C = nC; h->st_has_mean = want_mean;
  }
  cudaStream_t st = h->own_stream;
  A3D_CUDA_OK(cudaMemcpyAsync(h->st_z, z, (size_t)B * D * 4, cudaMemcpyHostToDevice, st));
  A3D_CUDA_OK(cudaMemcpyAsync(h->st_mask, mask, (size_t)B * D * 4, cudaMemcpyHostToDevice, st));
  if (mu_table) A3D_CUDA_OK(cudaMemcpyAsync(h->st_mu, mu_table, (size_t)C * D * 4, cudaMemcpyHostToDevice, st));
  A3D_CUDA_OK(cudaMemcpyAsync(h->st_bits, target_bits, (size_t)B * (A3D_VOXELS / 8), cudaMemcpyHostToDevice, st));
  cudaGetLastError();   // see a3d_decode_host
  if ((rc = a3d_impute(h, h->st_z, h->st_mask, h->st_mu, C, B, K, seed, obj_offset, fill_mode, h->st_zout, nullptr, st)))
    return rc;
  if ((rc = a3d_anytime_eval(h, h->st_zout, B, K, h->st_bits, thr, reinterpret_cast<int64_t*>(h->st_counts),
                             mean_prob_or_null ? h->st_mean : nullptr, st)))
    return rc;
  A3D_CUDA_OK(cudaMemcpyAsync(counts, h->st_counts, (size_t)B * 3 * 8, cudaMemcpyDeviceToHost, st));
  if (mean_prob_or_null)
    A3D_CUDA_OK(cudaMemcpyAsync(mean_prob_or_null, h->st_mean, (size_t)B * A3D_VOXELS * 4, cudaMemcpyDeviceToHost, st));
  cudaError_t e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    set_error("a3d_anytime_eval_host: %s", cudaGetErrorString(e));
    h->sticky = A3D_ERR_CUDA;
    return A3D_ERR_CUDA;
  }
  return A3D_OK;
}

int a3d_sampling(const float* mu_dev, const float* logvar_dev, int64_t n, int D, uint64_t seed, uint64_t obj_offset,
                 float* z_dev, void* stream) {
  if (n < 0 || D < 1 || (n > 0 && (!mu_dev || !logvar_dev || !z_dev))) { set_error("a3d_sampling: bad arguments"); return A3D_ERR_INVALID; }
  cudaGetLastError();
  return launch_sampling(mu_dev, logvar_dev, n, D, seed, obj_offset, z_dev, (cudaStream_t)stream, nullptr);
}

int a3d_nearest_prior(const float* z_dev, int64_t z_stride, const float* mu_table_dev, int C, int D,
                      const float* labels_dev, int64_t B, int32_t* idx_out_dev, int32_t* hits_dev, void* stream) {
  if (B < 0 || C < 1 || D < 1 || z_stride < D || (B > 0 && (!z_dev || !mu_table_dev)) || (hits_dev && !labels_dev)) {
    set_error("a3d_nearest_prior: bad arguments");
    return A3D_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  cudaGetLastError();
  if (hits_dev) A3D_CUDA_OK(cudaMemsetAsync(hits_dev, 0, sizeof(int32_t), st));
  return launch_nearest_prior(z_dev, z_stride, mu_table_dev, C, D, labels_dev, B, idx_out_dev, hits_dev, st, nullptr);
}

int a3d_decode_host(a3d_handle* h, const float* z_host, int64_t n, void* out_host, int out_dtype, float thr) {
  int rc = check_handle(h);
  if (rc) return rc;
  if (n < 0 || out_dtype < A3D_OUT_F32 || out_dtype > A3D_OUT_BITS || (n > 0 && (!z_host || !out_host))) {
    set_error("a3d_decode_host: bad arguments");
    return A3D_ERR_INVALID;
  }
  if (n == 0) return A3D_OK;
  if ((rc = sticky(h, finalize_weights(h)))) return rc;
  const int D = h->desc.latent_dim;
  // sub-chunk: small enough that the first (un-overlapped) decode is short against the copies, large enough to keep the
  // decoder kernels efficient; a sub-chunk's D2H copy (1 MiB per fp32 grid) takes ~3x its decode time
  int64_t sub = n <= 512 ? 16 : (n <= 2048 ? 64 : 256);
  if (sub > h->max_chunk) sub = h->max_chunk;
  if (!h->copy_stream) {
    A3D_CUDA_OK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      A3D_CUDA_OK(cudaEventCreateWithFlags(&h->dh_done[i], cudaEventDisableTiming));
      A3D_CUDA_OK(cudaEventCreateWithFlags(&h->dh_free[i], cudaEventDisableTiming));
    }
  }
  if (n > h->dh_z_cap) {
    cudaFree(h->dh_z); h->dh_z = nullptr; h->dh_z_cap = 0;
    A3D_CUDA_OK(cudaMalloc(&h->dh_z, (size_t)n * D * 4));
    h->dh_z_cap = n;
  }
  if (sub > h->dh_sub_cap) {
    for (int i = 0; i < 2; ++i) { cudaFree(h->dh_grid[i]); cudaFree(h->dh_out[i]); h->dh_grid[i] = nullptr; h->dh_out[i] = nullptr; }
    h->dh_sub_cap = 0;
    for (int i = 0; i < 2; ++i) {
      A3D_CUDA_OK(cudaMalloc(&h->dh_grid[i], (size_t)sub * A3D_VOXELS * 4));
      A3D_CUDA_OK(cudaMalloc(&h->dh_out[i], (size_t)sub * A3D_VOXELS * 2));
    }
    h->dh_sub_cap = sub;
  }
  cudaStream_t cs = h->own_stream, ps = h->copy_stream;
  const size_t per = out_dtype == A3D_OUT_F32 ? (size_t)A3D_VOXELS * 4 : out_dtype == A3D_OUT_F16 ? (size_t)A3D_VOXELS * 2
                                                                                               : (size_t)A3D_VOXELS / 8;
  A3D_CUDA_OK(cudaMemcpyAsync(h->dh_z, z_host, (size_t)n * D * 4, cudaMemcpyHostToDevice, cs));
  cudaGetLastError();   // a copy from pageable memory can leave a benign error behind (see check_handle): not a launch failure
  const int64_t nsub = (n + sub - 1) / sub;
  auto compute = [&](int64_t i) -> int {
    const int b = (int)(i & 1);
    const int64_t off = i * sub, nc = (n - off < sub) ? n - off : sub;
    if (i >= 2) A3D_CUDA_OK(cudaStreamWaitEvent(cs, h->dh_free[b], 0));
    int r = sticky(h, run_hidden(h, h->dh_z + off * D, nc, cs));
    if (r) return r;
    if ((r = sticky(h, run_tail(h, nc, 1, nullptr, thr, nullptr, h->dh_grid[b], 0.f, nullptr, cs)))) return r;
    if (out_dtype != A3D_OUT_F32 &&
        (r = sticky(h, launch_grid_convert(h->dh_grid[b], nc * (int64_t)A3D_VOXELS, out_dtype, thr, h->dh_out[b], cs,
                                           &h->launches))))
      return r;
    A3D_CUDA_OK(cudaEventRecord(h->dh_done[b], cs));
    return A3D_OK;
  };
  if ((rc = compute(0))) return rc;
  for (int64_t i = 0; i < nsub; ++i) {
    // queue the next decode BEFORE this sub-chunk's copy: with a pageable out_host the copy blocks the host, and the
    // GPU should be busy meanwhile
    if (i + 1 < nsub && (rc = compute(i + 1))) return rc;
    const int b = (int)(i & 1);
    const int64_t off = i * sub, nc = (n - off < sub) ? n - off : sub;
    A3D_CUDA_OK(cudaStreamWaitEvent(ps, h->dh_done[b], 0));
    const void* src = out_dtype == A3D_OUT_F32 ? (const void*)h->dh_grid[b] : (const void*)h->dh_out[b];
    A3D_CUDA_OK(cudaMemcpyAsync(static_cast<uint8_t*>(out_host) + (size_t)off * per, src, (size_t)nc * per,
                                cudaMemcpyDeviceToHost, ps));
    A3D_CUDA_OK(cudaEventRecord(h->dh_free[b], ps));
    cudaGetLastError();
  }
  cudaError_t e = cudaStreamSynchronize(ps);
  if (e == cudaSuccess) e = cudaStreamSynchronize(cs);
  cudaGetLastError();   // the driver's staging path of a pageable D2H copy can leave a benign error in the last-error slot
  if (e != cudaSuccess) {
    set_error("a3d_decode_host: %s", cudaGetErrorString(e));
    h->sticky = A3D_ERR_CUDA;
    return A3D_ERR_CUDA;
  }
  return A3D_OK;
}

int a3d_debug_time_tail(a3d_handle* h, int64_t B, int K, const uint8_t* target_bits_dev, int64_t* counts_dev, int reps,
                        float* ms_per_launch, void* stream) {
  int rc = check_handle(h);
  if (rc) return rc;
  if (B < 1 || K < 1 || reps < 1 || !target_bits_dev || !counts_dev || !ms_per_launch || B * K > h->last_chunk_n) {
    set_error("a3d_debug_time_tail: bad arguments (B * K must not exceed the %lld decodes of the last chunk)",
              (long long)h->last_chunk_n);
    return A3D_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  A3D_CUDA_OK(cudaMemsetAsync(counts_dev, 0, (size_t)B * 3 * sizeof(int64_t), st));
  for (int i = 0; i < 2; ++i)   // warm-up
    if ((rc = sticky(h, run_tail(h, B, K, target_bits_dev, 0.5f, reinterpret_cast<unsigned long long*>(counts_dev), nullptr,
                                 0.f, nullptr, st)))) return rc;
  A3D_CUDA_OK(cudaEventRecord(h->ev[0], st));
  for (int i = 0; i < reps; ++i)
    if ((rc = sticky(h, run_tail(h, B, K, target_bits_dev, 0.5f, reinterpret_cast<unsigned long long*>(counts_dev), nullptr,
                                 0.f, nullptr, st)))) return rc;
  A3D_CUDA_OK(cudaEventRecord(h->ev[1], st));
  A3D_CUDA_OK(cudaEventSynchronize(h->ev[1]));
  float ms = 0.f;
  A3D_CUDA_OK(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
  *ms_per_launch = ms / (float)reps;
  return A3D_OK;
}

int a3d_debug_read_layer(a3d_handle* h, int layer, int64_t n, float* host, size_t nbytes) {
  int rc = check_handle(h);
  if (rc) return rc;
  if (layer < 0 || layer > 4 || n <= 0 || n > h->max_chunk || !host || nbytes != h->act_elems[layer] * (size_t)n * 4) {
    set_error("a3d_debug_read_layer: bad arguments");
    return A3D_ERR_INVALID;
  }
  float* tmp = nullptr;
  A3D_CUDA_OK(cudaMalloc(&tmp, nbytes));
  A3D_CUDA_OK(cudaDeviceSynchronize());
  rc = launch_to_f32(h->act[layer], tmp, (int64_t)(nbytes / 4), h->desc.operand_dtype, 0);
  if (rc == A3D_OK) {
    cudaError_t e = cudaMemcpy(host, tmp, nbytes, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_error("debug copy failed: %s", cudaGetErrorString(e)); rc = A3D_ERR_CUDA; }
  }
  cudaFree(tmp);
  return rc;
}

}  // extern "C"
