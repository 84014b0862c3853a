// Host side of the image encoder (a3d_enc2d_* entry points of include/a3d.h): Darknet19 backbone + head2D of the
// Pascal3D path, src/net_core/darknet.py:83-168, called as head(backbone(images)) at src/module/nolbo.py:869.
// Owns the Keras-order weights, folds BatchNorm, repacks the Conv2D kernels to [tap][co][ci] 16-bit rows, builds the TMA
// tensor maps once and drives the per-chunk launch sequence on the caller's stream.
//
// Execution plan: a conv followed by a max-pool becomes ONE launch (pool fused into the epilogue); the 3-channel first
// layer runs on CUDA cores (K = 27); every other conv is a tcgen05 implicit GEMM; a final conv followed by a global
// pool writes fp32 and a small reduction kernel produces [n, C].  All hidden activations are 16-bit NHWC with the
// channel count padded to a multiple of 64 (pad channels are zero: zero weight rows / scale 0 / shift 0); the 32-channel
// output of the image layer stays 32 wide and the conv that reads it runs 32-channel K steps (SWIZZLE_64B operands).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>

#include "internal.h"

using namespace a3d;

namespace {

enum { OP_FIRST = 0, OP_CONV = 1, OP_POOL = 2, OP_GPOOL = 3 };

struct Op2d {
  int kind = 0;
  int layer = 0;              // index into the layer list of the layer whose output this op produces
  int H = 0, W = 0;           // input spatial size
  int cin = 0, cin_pad = 0, cout = 0, cout_pad = 0, ksize = 0, act = 0;
  bool bn = false, pool = false, out_f32 = false, is_max = false;
  int in_buf = -1, out_buf = -1;
  int w_index = -1;           // index of the kernel variable in the Keras weight list
  int bn_tile = 0;
  Conv2dGeom g;
  CUtensorMap tmap_act, tmap_wgt, tmap_wgt_half;   // _half: (64, 128)-row box for the 2-CTA kernel (256-wide N tiles)
  void* wgt = nullptr;        // device: 16-bit [tap][cout_pad][cin_pad]; first layer: fp32 [27][32] (CUDA-core kernel)
  void* wgt_first16 = nullptr; // first layer, tensor-core kernel: 16-bit [32 co][32 k], k = tap*3 + ci, 5 zero columns
  float *scale = nullptr, *shift = nullptr;
};

struct Buf2d {
  int H = 0, W = 0, C = 0, C_pad = 0;
  bool f32 = false;
  int64_t alloc_n = 0;
  void* ptr = nullptr;
};

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

}  // namespace

struct a3d_enc2d {
  a3d_enc2d_desc desc{};
  int num_sms = 0;
  std::vector<Op2d> ops;
  std::vector<Buf2d> bufs;          // bufs[0] = imported input (feature-map models only)
  std::vector<int> layer_buf;       // layer index -> buffer holding its output (pooled for conv+pool pairs)
  std::vector<std::vector<float>> w;
  std::vector<int64_t> w_numel;
  std::vector<bool> w_set;
  bool dirty = true;
  int out_h = 0, out_w = 0, out_c = 0;
  bool out_is_pooled_f32 = false;
  int final_buf = -1;
  size_t arena_bytes = 0;
  int64_t launches = 0;
  int64_t last_n = 0;
  int sticky = 0;
  // host-call pipeline (a3d_enc2d_forward_host): two device input stages, a copy stream and a compute stream
  cudaStream_t hs_compute = nullptr, hs_copy = nullptr;
  cudaEvent_t hs_copied[2] = {nullptr, nullptr}, hs_consumed[2] = {nullptr, nullptr};
  float* hs_in[2] = {nullptr, nullptr};
  uint8_t* hs_in_u8[2] = {nullptr, nullptr};   // a3d_enc2d_forward_host_u8: byte stages in front of hs_in
  float* hs_out = nullptr;
  int64_t hs_out_n = 0;
  bool no_p4 = false;         // A3D_ENC_P4=0: keep the shuffle pool in the resident-weight variant (cross-check)
  bool no_rh = false;         // A3D_ENC_RH=0: disable the resident-weight conv variant (cross-check)
  bool first_simt = false;    // A3D_ENC_FIRST=simt: CUDA-core image layer (diagnostic cross-check of the tensor-core one)
};

namespace {

int check(const a3d_enc2d* h) {
  cudaGetLastError();   // drop a stale non-sticky error left by another runtime user (see check_handle in handle.cu)
  if (!h) { set_error("null encoder handle"); return A3D_ERR_INVALID; }
  if (h->sticky) { set_error("encoder handle is in a sticky CUDA error state (%d)", h->sticky); return h->sticky; }
  return A3D_OK;
}
int sticky(a3d_enc2d* h, int rc) {
  if (rc == A3D_ERR_CUDA) h->sticky = rc;
  return rc;
}

// brick of the conv M tile: contiguous pixels for plain convs, at most 16 wide (and >= 2 high) when a pool is fused
void choose_brick(Op2d& op, bool rh_ok, bool p4_ok) {
  const int W = op.W, H = op.H;
  // power-of-two brick (largest that fits, <= 16 wide when a pool is fused); it may overhang images whose size is not a
  // power of two: overhanging rows read zeros through TMA and are masked in the epilogue
  const int wcap = op.pool ? 16 : 128;
  int lw = 0, lh = 0;
  while ((2 << lw) <= W && (2 << lw) <= wcap) ++lw;
  while (lw + lh < 7 && (2 << lh) <= H) ++lh;
  const int wt = 1 << lw, ht = 1 << lh;
  Conv2dGeom& g = op.g;
  g.H = H; g.W = W;
  g.lw = lw; g.lh = lh;
  g.tiles_w = (W + wt - 1) / wt; g.tiles_h = (H + ht - 1) / ht;
  g.taps = op.ksize * op.ksize;
  g.kc = op.cin_pad % 64 == 0 ? 64 : 32;
  g.cin_chunks = op.cin_pad / g.kc;
  g.cout_pad = op.cout_pad;
  g.cout_real = op.cout;
  g.n_tiles = op.cout_pad / op.bn_tile;
  // resident-weight variant (conv2d_tc.cu, CfgRH): shallow 3 x 3 layers whose nine weight tiles fit in shared memory
  const bool rh_shape = (g.kc == 32 && op.bn_tile == 64 && op.cout_pad == 64) || (g.kc == 64 && op.bn_tile == 128 && op.cout_pad == 128);
  if (rh_ok && op.ksize == 3 && g.cin_chunks == 1 && g.n_tiles == 1 && rh_shape && !op.out_f32 && W % 16 == 0 && H % 8 == 0) {
    g.rh = 1;
    g.lw = 4; g.lh = 3;                       // 16 x 8 brick (also for the plain 16-bit mode)
    g.tiles_w = W / 16; g.tiles_h = H / 8;
    if (op.pool && op.bn_tile == 64 && p4_ok && W % 32 == 0 && H % 16 == 0) {
      g.rh = 2;                               // pool through four accumulators: the 16 x 8 brick tiles the pooled grid
      g.tiles_w = (W / 2) / 16; g.tiles_h = (H / 2) / 8;
    }
  }
}

int build_plan(a3d_enc2d* h) {
  const a3d_enc2d_desc& d = h->desc;
  int H = d.in_h, W = d.in_w, C = d.in_ch;
  int cur = -1;     // buffer holding the current activation; -1 = the user's fp32 image
  int C_pad = C;
  h->layer_buf.assign(d.num_layers, -1);
  if (C % 64 == 0) {
    Buf2d b; b.H = H; b.W = W; b.C = C; b.C_pad = C;
    h->bufs.push_back(b);
    cur = 0;
  } else if (C != 3) {
    set_error("in_ch must be 3 (images) or a multiple of 64 (feature maps), got %d", C);
    return A3D_ERR_INVALID;
  }
  int w_index = 0;
  for (int li = 0; li < d.num_layers; ++li) {
    const a3d_layer2d& L = d.layers[li];
    const bool next_pool = li + 1 < d.num_layers && d.layers[li + 1].kind == A3D_L2D_MAXPOOL;
    const bool next_gpool = li + 1 < d.num_layers && (d.layers[li + 1].kind == A3D_L2D_GLOBAL_MAX ||
                                                      d.layers[li + 1].kind == A3D_L2D_GLOBAL_AVG);
    if (L.kind == A3D_L2D_CONV) {
      if ((L.ksize != 1 && L.ksize != 3) || L.filters < 1 || L.activation < 0 || L.activation > A3D_ACT_LRELU01) {
        set_error("layer %d: Conv2D needs ksize 1 or 3, filters >= 1 and a known activation", li);
        return A3D_ERR_INVALID;
      }
      Op2d op;
      op.layer = li; op.H = H; op.W = W; op.cin = C; op.cin_pad = C_pad; op.cout = L.filters;
      op.cout_pad = cur < 0 ? 32 : round_up(L.filters, 64); op.ksize = L.ksize; op.act = L.activation; op.bn = L.batch_norm != 0;
      op.in_buf = cur; op.w_index = w_index;
      h->w_numel.push_back((int64_t)L.ksize * L.ksize * C * L.filters);
      if (op.bn) for (int i = 0; i < 4; ++i) h->w_numel.push_back(L.filters);
      w_index += op.bn ? 5 : 1;
      if (cur < 0) {
        // the 3-channel image layer: CUDA-core kernel with the following pool fused (darknet.py:99-100)
        if (L.ksize != 3 || L.filters != 32 || !next_pool || (H & 1) || (W & 1)) {
          set_error("layer %d: the image layer must be Conv2D(32, 3) followed by MaxPool2D(2,2) on an even size %d x %d (Darknet19)", li, H, W);
          return A3D_ERR_INVALID;
        }
        op.kind = OP_FIRST; op.pool = true;
      } else {
        op.kind = OP_CONV;
        op.pool = next_pool;
        op.out_f32 = next_gpool;
        if (op.pool && ((H & 1) || (W & 1))) {
          set_error("layer %d: MaxPool2D(2,2) on an odd size %d x %d (this build needs even sizes at every pool: image sizes that are multiples of 32 for Darknet19)", li, H, W);
          return A3D_ERR_INVALID;
        }
        op.bn_tile = op.cin_pad % 64 == 0 ? conv2d_tc_bn(op.cout_pad) : 64;   // 32-channel K steps: N tile 64 only
        choose_brick(op, !h->no_rh, !h->no_p4);
        if (op.pool && op.g.lh < 1) { set_error("layer %d: fused pool needs H >= 2", li); return A3D_ERR_INVALID; }
      }
      Buf2d b;
      b.H = op.pool ? H / 2 : H; b.W = op.pool ? W / 2 : W; b.C = L.filters;
      b.C_pad = op.out_f32 ? L.filters : op.cout_pad; b.f32 = op.out_f32;
      h->bufs.push_back(b);
      op.out_buf = (int)h->bufs.size() - 1;
      h->ops.push_back(op);
      cur = op.out_buf;
      h->layer_buf[li] = cur;
      H = b.H; W = b.W; C = L.filters; C_pad = b.C_pad;
      if (op.pool) { h->layer_buf[li + 1] = cur; ++li; }
    } else if (L.kind == A3D_L2D_MAXPOOL) {
      if (cur < 0 || (H & 1) || (W & 1) || h->bufs[cur].f32) { set_error("layer %d: unsupported MaxPool2D placement", li); return A3D_ERR_INVALID; }
      Op2d op;
      op.kind = OP_POOL; op.layer = li; op.H = H; op.W = W; op.cin = C; op.cin_pad = C_pad; op.cout = C; op.cout_pad = C_pad;
      op.in_buf = cur;
      Buf2d b; b.H = H / 2; b.W = W / 2; b.C = C; b.C_pad = C_pad;
      h->bufs.push_back(b);
      op.out_buf = (int)h->bufs.size() - 1;
      h->ops.push_back(op);
      cur = op.out_buf; h->layer_buf[li] = cur; H /= 2; W /= 2;
    } else if (L.kind == A3D_L2D_GLOBAL_MAX || L.kind == A3D_L2D_GLOBAL_AVG) {
      if (cur < 0 || !h->bufs[cur].f32 || li != d.num_layers - 1) {
        set_error("layer %d: a global pool must be the last layer and directly follow a Conv2D (head2D)", li);
        return A3D_ERR_INVALID;
      }
      Op2d op;
      op.kind = OP_GPOOL; op.layer = li; op.H = H; op.W = W; op.cin = C; op.cout = C; op.is_max = L.kind == A3D_L2D_GLOBAL_MAX;
      op.in_buf = cur; op.out_buf = -1;   // writes the caller's buffer
      h->ops.push_back(op);
      h->layer_buf[li] = -1;
      H = 1; W = 1;
      h->out_is_pooled_f32 = true;
    } else {
      set_error("layer %d: unknown kind %d", li, L.kind);
      return A3D_ERR_INVALID;
    }
  }
  if (h->ops.empty()) { set_error("empty layer list"); return A3D_ERR_INVALID; }
  h->out_h = H; h->out_w = W; h->out_c = C;
  h->final_buf = cur;
  if (!h->out_is_pooled_f32 && h->bufs[cur].f32) { set_error("internal: fp32 final buffer without a global pool"); return A3D_ERR_INVALID; }
  h->w.assign(h->w_numel.size(), {});
  h->w_set.assign(h->w_numel.size(), false);
  return A3D_OK;
}

int alloc_arena(a3d_enc2d* h) {
  // each buffer holds max_batch images rounded up to the image count of its consumer's GEMM brick (TMA boxes stay
  // inside the allocation; rows past the real batch are never stored)
  for (auto& b : h->bufs) b.alloc_n = h->desc.max_batch;
  for (auto& op : h->ops)
    if (op.kind == OP_CONV) {
      const int nt = 128 >> (op.g.lw + op.g.lh);
      Buf2d& b = h->bufs[op.in_buf];
      const int64_t need = (int64_t)round_up(h->desc.max_batch, nt);
      if (need > b.alloc_n) b.alloc_n = need;
    }
  for (auto& b : h->bufs) {
    const size_t bytes = (size_t)b.alloc_n * b.H * b.W * b.C_pad * (b.f32 ? 4 : 2);
    cudaError_t e = cudaMalloc(&b.ptr, bytes);
    if (e != cudaSuccess) { set_error("encoder arena allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e)); return A3D_ERR_CUDA; }
    A3D_CUDA_OK(cudaMemset(b.ptr, 0, bytes));
    h->arena_bytes += bytes;
  }
  return A3D_OK;
}

int make_maps(a3d_enc2d* h, Op2d& op) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return A3D_ERR_CUDA; }
  const CUtensorMapDataType dt =
      h->desc.operand_dtype == A3D_DTYPE_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const Buf2d& b = h->bufs[op.in_buf];
  const uint64_t C = op.cin_pad, W = op.W, H = op.H;
  cuuint64_t dims[4] = {C, W, H, (cuuint64_t)b.alloc_n};
  cuuint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
  const cuuint32_t kc = (cuuint32_t)op.g.kc;
  const CUtensorMapSwizzle swz = kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  cuuint32_t box[4] = {kc, 1u << op.g.lw, 1u << op.g.lh, 128u >> (op.g.lw + op.g.lh)};
  cuuint32_t es[4] = {1, 1, 1, 1};
  if (op.g.rh == 1) box[2] = 10;   // haloed box: 16 x (8 + 2) pixels, one load per dx serves the three dy taps
  if (op.g.rh == 2) {              // element-strided box: every second pixel, 16 x 9 pooled positions
    box[1] = 32; box[2] = 18; box[3] = 1;
    es[1] = 2; es[2] = 2;
  }
  CUresult r = enc(&op.tmap_act, dt, 4, b.ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(encoder activations, layer %d) failed: %d", op.layer, (int)r); return A3D_ERR_CUDA; }
  cuuint64_t wd[2] = {C, (cuuint64_t)op.ksize * op.ksize * op.cout_pad};
  cuuint64_t ws[1] = {C * 2};
  cuuint32_t wb[2] = {kc, (cuuint32_t)op.bn_tile};
  cuuint32_t es1[2] = {1, 1};
  r = enc(&op.tmap_wgt, dt, 2, op.wgt, wd, ws, wb, es1, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(encoder weights, layer %d) failed: %d", op.layer, (int)r); return A3D_ERR_CUDA; }
  if (op.bn_tile == 256 && kc == 64) {
    cuuint32_t wh[2] = {64, 128};
    r = enc(&op.tmap_wgt_half, dt, 2, op.wgt, wd, ws, wh, es1, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(encoder half weights, layer %d) failed: %d", op.layer, (int)r); return A3D_ERR_CUDA; }
  }
  return A3D_OK;
}

int finalize(a3d_enc2d* h) {
  if (!h->dirty) return A3D_OK;
  for (size_t i = 0; i < h->w_set.size(); ++i)
    if (!h->w_set[i]) { set_error("encoder weight %zu of %zu was never set", i, h->w_set.size()); return A3D_ERR_WEIGHTS; }
  const int fmt = h->desc.operand_dtype;
  int rc;
  std::vector<float> sc, sf;
  std::vector<uint16_t> p16;
  for (auto& op : h->ops) {
    if (op.kind != OP_FIRST && op.kind != OP_CONV) continue;
    const std::vector<float>& k = h->w[op.w_index];     // Keras Conv2D kernel [kh][kw][cin][cout]
    if (op.bn) fold_bn(h->w[op.w_index + 1], h->w[op.w_index + 2], h->w[op.w_index + 3], h->w[op.w_index + 4], sc, sf);
    else { sc.assign(op.cout, 1.f); sf.assign(op.cout, 0.f); }
    sc.resize(op.cout_pad, 0.f);     // pad channels: scale 0, shift 0 -> act(0) = 0 for every supported activation
    sf.resize(op.cout_pad, 0.f);
    if ((rc = upload(sc.data(), sc.size() * 4, (void**)&op.scale))) return rc;
    if ((rc = upload(sf.data(), sf.size() * 4, (void**)&op.shift))) return rc;
    const int taps = op.ksize * op.ksize;
    if (op.kind == OP_FIRST) {
      if ((rc = upload(k.data(), k.size() * 4, &op.wgt))) return rc;   // already [tap][ci][co] = [27][32]
      p16.assign(32 * 32, cvt16(0.f, fmt));
      for (int kk = 0; kk < 27; ++kk)
        for (int co = 0; co < 32; ++co) p16[(size_t)co * 32 + kk] = cvt16(k[(size_t)kk * 32 + co], fmt);
      if ((rc = upload(p16.data(), p16.size() * 2, &op.wgt_first16))) return rc;
    } else {
      p16.assign((size_t)taps * op.cout_pad * op.cin_pad, cvt16(0.f, fmt));
      for (int t = 0; t < taps; ++t)
        for (int ci = 0; ci < op.cin; ++ci) {
          const float* src = &k[((size_t)t * op.cin + ci) * op.cout];
          for (int co = 0; co < op.cout; ++co)
            p16[((size_t)t * op.cout_pad + co) * op.cin_pad + ci] = cvt16(src[co], fmt);
        }
      if ((rc = upload(p16.data(), p16.size() * 2, &op.wgt))) return rc;
      if ((rc = make_maps(h, op))) return rc;
    }
  }
  h->dirty = false;
  return A3D_OK;
}

int run_chunk(a3d_enc2d* h, const void* in_dev, int in_dtype, int64_t n, void* out_dev, int out_dtype, cudaStream_t st) {
  const int fmt = h->desc.operand_dtype;
  int rc;
  if (h->desc.in_ch % 64 == 0) {
    const Buf2d& b = h->bufs[0];
    if ((rc = launch_import_nhwc(in_dev, in_dtype == A3D_IO_F32, b.ptr, n * b.H * b.W, b.C, b.C_pad, fmt, st, &h->launches)))
      return rc;
  }
  for (auto& op : h->ops) {
    void* out = op.out_buf >= 0 ? h->bufs[op.out_buf].ptr : out_dev;
    switch (op.kind) {
      case OP_FIRST:
        if (!h->first_simt)
          rc = launch_conv2d_first_tc(reinterpret_cast<const float*>(in_dev), op.wgt_first16, op.scale, op.shift, out, n,
                                      op.H, op.W, op.cout_pad, fmt, op.act, h->num_sms, st, &h->launches);
        else
        rc = launch_conv2d_first_pool(reinterpret_cast<const float*>(in_dev), reinterpret_cast<const float*>(op.wgt),
                                      op.scale, op.shift, out, n, op.H, op.W, op.cout_pad, fmt, op.act, st, &h->launches);
        break;
      case OP_CONV: {
        Conv2dGeom g = op.g;
        const int nt = 128 >> (g.lw + g.lh);
        g.n_images = (int)n;
        g.m_tiles = (int)((n + nt - 1) / nt) * g.tiles_w * g.tiles_h;
        if (conv2d_pair_eligible(g, op.bn_tile, op.pool, op.out_f32))
          rc = launch_conv2d_pair(op.tmap_act, op.tmap_wgt_half, out, op.scale, op.shift, g, fmt, op.act, op.pool,
                                  h->num_sms, st, &h->launches);
        else
          rc = launch_conv2d_tc(op.tmap_act, op.tmap_wgt, out, op.scale, op.shift, g, op.bn_tile, fmt, op.act, op.pool,
                                op.out_f32, h->num_sms, st, &h->launches);
        break;
      }
      case OP_POOL:
        rc = launch_maxpool2d(h->bufs[op.in_buf].ptr, out, n, op.H, op.W, op.cin_pad, fmt, st, &h->launches);
        break;
      case OP_GPOOL:
        rc = launch_global_pool(reinterpret_cast<const float*>(h->bufs[op.in_buf].ptr), reinterpret_cast<float*>(out_dev),
                                n, op.H * op.W, op.cin, op.is_max ? 1 : 0, st, &h->launches);
        break;
      default: rc = A3D_ERR_INVALID;
    }
    if (rc) return rc;
  }
  if (!h->out_is_pooled_f32) {
    const Buf2d& b = h->bufs[h->final_buf];
    if ((rc = launch_export_nhwc(b.ptr, out_dev, out_dtype == A3D_IO_F32, n * b.H * b.W, b.C, b.C_pad, fmt, st, &h->launches)))
      return rc;
  }
  h->last_n = n;
  return A3D_OK;
}

}  // namespace

extern "C" {

int a3d_enc2d_create(const a3d_enc2d_desc* d, a3d_enc2d** out) {
  if (!d || !out) { set_error("null argument"); return A3D_ERR_INVALID; }
  *out = nullptr;
  if (d->abi_version != A3D_ABI_VERSION) { set_error("ABI version mismatch: %d vs %d", d->abi_version, A3D_ABI_VERSION); return A3D_ERR_INVALID; }
  if (d->num_layers < 1 || d->num_layers > A3D_ENC_MAX_LAYERS) { set_error("num_layers must be in [1, %d]", A3D_ENC_MAX_LAYERS); return A3D_ERR_INVALID; }
  if (d->in_h < 1 || d->in_w < 1 || d->in_h > 8192 || d->in_w > 8192) {
    set_error("in_h / in_w must be in [1, 8192] (got %d x %d)", d->in_h, d->in_w);
    return A3D_ERR_INVALID;
  }
  if (d->operand_dtype != A3D_DTYPE_F16 && d->operand_dtype != A3D_DTYPE_BF16) { set_error("invalid operand dtype"); return A3D_ERR_INVALID; }
  if (d->max_batch < 1) { set_error("max_batch must be >= 1"); return A3D_ERR_INVALID; }
  a3d_enc2d* h = new a3d_enc2d();
  h->desc = *d;
  { const char* e = getenv("A3D_ENC_RH"); h->no_rh = e && std::string(e) == "0"; }
  { const char* e = getenv("A3D_ENC_P4"); h->no_p4 = e && std::string(e) == "0"; }
  int rc = build_plan(h);   // validates the structure before any device work (usable without a GPU for error paths)
  if (rc) { delete h; return rc; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || d->device >= ndev) {
    cudaGetLastError();
    set_error("no CUDA device %d available; liba3d has no CPU path", d->device);
    delete h;
    return A3D_ERR_NO_DEVICE;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, d->device) != cudaSuccess || prop.major != 10) {
    set_error("device %d is not sm_100; liba3d is built for sm_100a only", d->device);
    delete h;
    return A3D_ERR_NO_DEVICE;
  }
  cudaSetDevice(d->device);
  h->num_sms = prop.multiProcessorCount;
  { const char* e = getenv("A3D_ENC_FIRST"); h->first_simt = e && std::string(e) == "simt"; }
  if ((rc = alloc_arena(h))) { a3d_enc2d_destroy(h); return rc; }
  *out = h;
  return A3D_OK;
}

void a3d_enc2d_destroy(a3d_enc2d* h) {
  if (!h) return;
  cudaDeviceSynchronize();
  for (auto& b : h->bufs) cudaFree(b.ptr);
  for (auto& op : h->ops) { cudaFree(op.wgt); cudaFree(op.wgt_first16); cudaFree(op.scale); cudaFree(op.shift); }
  for (int i = 0; i < 2; ++i) {
    cudaFree(h->hs_in[i]);
    cudaFree(h->hs_in_u8[i]);
    if (h->hs_copied[i]) cudaEventDestroy(h->hs_copied[i]);
    if (h->hs_consumed[i]) cudaEventDestroy(h->hs_consumed[i]);
  }
  cudaFree(h->hs_out);
  if (h->hs_compute) cudaStreamDestroy(h->hs_compute);
  if (h->hs_copy) cudaStreamDestroy(h->hs_copy);
  delete h;
}

int a3d_enc2d_num_weights(const a3d_enc2d* h) { return h ? (int)h->w_numel.size() : 0; }
int64_t a3d_enc2d_weight_numel(const a3d_enc2d* h, int index) {
  return (h && index >= 0 && index < (int)h->w_numel.size()) ? h->w_numel[index] : -1;
}

int a3d_enc2d_set_weight(a3d_enc2d* h, int index, const float* host, size_t nbytes) {
  int rc = check(h);
  if (rc) return rc;
  if (index < 0 || index >= (int)h->w_numel.size() || !host) { set_error("bad encoder weight index %d", index); return A3D_ERR_INVALID; }
  if (nbytes != (size_t)h->w_numel[index] * 4) {
    set_error("encoder weight %d: expected %lld fp32 values, got %zu bytes", index, (long long)h->w_numel[index], nbytes);
    return A3D_ERR_WEIGHTS;
  }
  h->w[index].assign(host, host + h->w_numel[index]);
  h->w_set[index] = true;
  h->dirty = true;
  return A3D_OK;
}

int a3d_enc2d_get_weight(const a3d_enc2d* h, int index, float* host, size_t nbytes) {
  if (!h || index < 0 || index >= (int)h->w_numel.size() || !host) { set_error("bad encoder weight index %d", index); return A3D_ERR_INVALID; }
  if (!h->w_set[index]) { set_error("encoder weight %d was never set", index); return A3D_ERR_WEIGHTS; }
  if (nbytes != (size_t)h->w_numel[index] * 4) { set_error("encoder weight %d: size mismatch", index); return A3D_ERR_WEIGHTS; }
  memcpy(host, h->w[index].data(), nbytes);
  return A3D_OK;
}

int a3d_enc2d_output_shape(const a3d_enc2d* h, int32_t* out_dims) {
  if (!h || !out_dims) { set_error("null argument"); return A3D_ERR_INVALID; }
  out_dims[0] = h->out_h; out_dims[1] = h->out_w; out_dims[2] = h->out_c;
  return A3D_OK;
}

int a3d_enc2d_forward(a3d_enc2d* h, const void* in_dev, int in_dtype, int64_t n, void* out_dev, int out_dtype,
                      void* stream) {
  int rc = check(h);
  if (rc) return rc;
  if (n < 0 || (n > 0 && (!in_dev || !out_dev))) { set_error("a3d_enc2d_forward: bad arguments"); return A3D_ERR_INVALID; }
  const bool image_in = h->desc.in_ch == 3;
  if ((image_in && in_dtype != A3D_IO_F32) || (!image_in && in_dtype != A3D_IO_F32 && in_dtype != h->desc.operand_dtype)) {
    set_error("a3d_enc2d_forward: input must be fp32%s", image_in ? " (images)" : " or the handle's operand dtype");
    return A3D_ERR_INVALID;
  }
  if ((h->out_is_pooled_f32 && out_dtype != A3D_IO_F32) || (out_dtype != A3D_IO_F32 && out_dtype != h->desc.operand_dtype)) {
    set_error("a3d_enc2d_forward: output must be fp32%s", h->out_is_pooled_f32 ? " (global pool)" : " or the handle's operand dtype");
    return A3D_ERR_INVALID;
  }
  if ((rc = sticky(h, finalize(h)))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t in_es = in_dtype == A3D_IO_F32 ? 4 : 2, out_es = out_dtype == A3D_IO_F32 ? 4 : 2;
  const size_t in_per = (size_t)h->desc.in_h * h->desc.in_w * h->desc.in_ch * in_es;
  const size_t out_per = (size_t)h->out_h * h->out_w * h->out_c * out_es;
  for (int64_t off = 0; off < n; off += h->desc.max_batch) {
    const int64_t nc = n - off < h->desc.max_batch ? n - off : h->desc.max_batch;
    rc = run_chunk(h, reinterpret_cast<const uint8_t*>(in_dev) + off * in_per, in_dtype, nc,
                   reinterpret_cast<uint8_t*>(out_dev) + off * out_per, out_dtype, st);
    if ((rc = sticky(h, rc))) return rc;
  }
  return A3D_OK;
}

namespace {
// images_host: fp32 (u8_scale == 0) or uint8 bytes scaled by u8_scale on the device
int forward_host_impl(a3d_enc2d* h, const void* images_host, float u8_scale, int64_t n, float* out_dev_or_null,
                      float* out_host_or_null) {
  const bool u8 = u8_scale != 0.f;
  int rc = check(h);
  if (rc) return rc;
  if (n < 0 || (n > 0 && (!images_host || (!out_dev_or_null && !out_host_or_null)))) {
    set_error("a3d_enc2d_forward_host: bad arguments");
    return A3D_ERR_INVALID;
  }
  if (n == 0) return A3D_OK;
  if ((rc = sticky(h, finalize(h)))) return rc;
  const size_t in_per = (size_t)h->desc.in_h * h->desc.in_w * h->desc.in_ch;       // fp32 elements per image
  const size_t out_per = (size_t)h->out_h * h->out_w * h->out_c;
  const int64_t mb = h->desc.max_batch;
  if (!h->hs_compute) {
    A3D_CUDA_OK(cudaStreamCreateWithFlags(&h->hs_compute, cudaStreamNonBlocking));
    A3D_CUDA_OK(cudaStreamCreateWithFlags(&h->hs_copy, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      A3D_CUDA_OK(cudaEventCreateWithFlags(&h->hs_copied[i], cudaEventDisableTiming));
      A3D_CUDA_OK(cudaEventCreateWithFlags(&h->hs_consumed[i], cudaEventDisableTiming));
      A3D_CUDA_OK(cudaMalloc(&h->hs_in[i], (size_t)mb * in_per * 4));
    }
  }
  if (u8 && !h->hs_in_u8[0])
    for (int i = 0; i < 2; ++i) A3D_CUDA_OK(cudaMalloc(&h->hs_in_u8[i], (size_t)mb * in_per));
  float* out_dev = out_dev_or_null;
  if (!out_dev) {
    if (n > h->hs_out_n) {
      cudaFree(h->hs_out);
      h->hs_out = nullptr;
      A3D_CUDA_OK(cudaMalloc(&h->hs_out, (size_t)n * out_per * 4));
      h->hs_out_n = n;
    }
    out_dev = h->hs_out;
  }
  const int64_t chunks = (n + mb - 1) / mb;
  auto copy_chunk = [&](int64_t c) -> int {
    const int b = (int)(c & 1);
    const int64_t off = c * mb, nc = n - off < mb ? n - off : mb;
    if (c >= 2) A3D_CUDA_OK(cudaStreamWaitEvent(h->hs_copy, h->hs_consumed[b], 0));   // stage b is free again
    if (u8)
      A3D_CUDA_OK(cudaMemcpyAsync(h->hs_in_u8[b], (const uint8_t*)images_host + off * in_per, (size_t)nc * in_per, cudaMemcpyHostToDevice, h->hs_copy));
    else
      A3D_CUDA_OK(cudaMemcpyAsync(h->hs_in[b], (const float*)images_host + off * in_per, (size_t)nc * in_per * 4, cudaMemcpyHostToDevice, h->hs_copy));
    A3D_CUDA_OK(cudaEventRecord(h->hs_copied[b], h->hs_copy));
    return A3D_OK;
  };
  if ((rc = copy_chunk(0))) return sticky(h, rc);
  for (int64_t c = 0; c < chunks; ++c) {
    const int b = (int)(c & 1);
    const int64_t off = c * mb, nc = n - off < mb ? n - off : mb;
    if (c + 1 < chunks && (rc = copy_chunk(c + 1))) return sticky(h, rc);               // overlaps the forward of chunk c
    A3D_CUDA_OK(cudaStreamWaitEvent(h->hs_compute, h->hs_copied[b], 0));
    if (u8 && (rc = sticky(h, launch_u8_to_f32(h->hs_in_u8[b], h->hs_in[b], (int64_t)nc * in_per, u8_scale, h->hs_compute, &h->launches)))) return rc;
    if ((rc = sticky(h, run_chunk(h, h->hs_in[b], A3D_IO_F32, nc, out_dev + off * out_per, A3D_IO_F32, h->hs_compute)))) return rc;
    A3D_CUDA_OK(cudaEventRecord(h->hs_consumed[b], h->hs_compute));
  }
  if (out_host_or_null)
    A3D_CUDA_OK(cudaMemcpyAsync(out_host_or_null, out_dev, (size_t)n * out_per * 4, cudaMemcpyDeviceToHost, h->hs_compute));
  cudaError_t e = cudaStreamSynchronize(h->hs_compute);
  if (e != cudaSuccess) {
    set_error("a3d_enc2d_forward_host: %s", cudaGetErrorString(e));
    h->sticky = A3D_ERR_CUDA;
    return A3D_ERR_CUDA;
  }
  return A3D_OK;
}
}  // namespace

int a3d_enc2d_forward_host(a3d_enc2d* h, const float* images_host, int64_t n, float* out_dev_or_null, float* out_host_or_null) {
  return forward_host_impl(h, images_host, 0.f, n, out_dev_or_null, out_host_or_null);
}

int a3d_enc2d_forward_host_u8(a3d_enc2d* h, const uint8_t* images_host, float scale, int64_t n, float* out_dev_or_null,
                              float* out_host_or_null) {
  if (!(scale > 0.f)) { set_error("a3d_enc2d_forward_host_u8: scale must be positive"); return A3D_ERR_INVALID; }
  return forward_host_impl(h, images_host, scale, n, out_dev_or_null, out_host_or_null);
}

int a3d_enc2d_split_sample(a3d_enc2d* h, const float* enc_out_dev, int64_t n, int D, int out_stride, float clip,
                           int seed_enable, uint64_t seed, uint64_t obj_offset, float* mean_dev, float* logvar_dev,
                           float* z_dev, void* stream) {
  int rc = check(h);
  if (rc) return rc;
  if (n < 0 || D < 1 || out_stride < 2 * D || (n > 0 && !enc_out_dev)) { set_error("a3d_enc2d_split_sample: bad arguments"); return A3D_ERR_INVALID; }
  return sticky(h, launch_split_sample(enc_out_dev, n, D, out_stride, clip, seed_enable, seed, obj_offset, mean_dev,
                                       logvar_dev, z_dev, (cudaStream_t)stream, &h->launches));
}

int a3d_enc2d_layer_shape(const a3d_enc2d* h, int layer, int32_t* dims) {
  if (!h || !dims || layer < 0 || layer >= h->desc.num_layers) { set_error("bad layer index"); return A3D_ERR_INVALID; }
  const int bi = h->layer_buf[layer];
  if (bi < 0) { dims[0] = 1; dims[1] = 1; dims[2] = h->out_c; dims[3] = h->out_c; return A3D_OK; }
  const Buf2d& b = h->bufs[bi];
  dims[0] = b.H; dims[1] = b.W; dims[2] = b.C; dims[3] = b.C_pad;
  return A3D_OK;
}

int a3d_enc2d_debug_read_layer(a3d_enc2d* h, int layer, int64_t n, float* host, size_t nbytes) {
  int rc = check(h);
  if (rc) return rc;
  if (layer < 0 || layer >= h->desc.num_layers || h->layer_buf[layer] < 0 || n <= 0 || n > h->desc.max_batch || !host) {
    set_error("a3d_enc2d_debug_read_layer: bad arguments");
    return A3D_ERR_INVALID;
  }
  const Buf2d& b = h->bufs[h->layer_buf[layer]];
  const size_t elems = (size_t)n * b.H * b.W * b.C_pad;
  if (nbytes != elems * 4) { set_error("a3d_enc2d_debug_read_layer: expected %zu bytes", elems * 4); return A3D_ERR_INVALID; }
  A3D_CUDA_OK(cudaDeviceSynchronize());
  if (b.f32) {
    A3D_CUDA_OK(cudaMemcpy(host, b.ptr, nbytes, cudaMemcpyDeviceToHost));
    return A3D_OK;
  }
  float* tmp = nullptr;
  A3D_CUDA_OK(cudaMalloc(&tmp, nbytes));
  rc = launch_to_f32(b.ptr, tmp, (int64_t)elems, h->desc.operand_dtype, 0);
  if (rc == A3D_OK) {
    cudaError_t e = cudaMemcpy(host, tmp, nbytes, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_error("debug copy failed: %s", cudaGetErrorString(e)); rc = A3D_ERR_CUDA; }
  }
  cudaFree(tmp);
  return rc;
}

int64_t a3d_enc2d_launch_count(const a3d_enc2d* h) { return h ? h->launches : 0; }
size_t a3d_enc2d_workspace_bytes(const a3d_enc2d* h) { return h ? h->arena_bytes : 0; }

}  // extern "C"
