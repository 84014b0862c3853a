// Host side of the voxel encoder (a3d_enc3d_* entry points of include/a3d.h): encoder3D(structure),
// src/net_core/autoencoder3D.py:72-102, called as self._encoder(voxels, training=False) at src/module/nolbo.py:1463.
// Owns the Keras-order weights, folds BatchNorm, repacks the Conv3D kernels [kd,kh,kw,Cin,Cout] -> [tap][co][ci] 16-bit,
// builds the element-strided 5-D TMA tensor maps once and drives the launch sequence on the caller's stream:
//   conv3d_first_tc (1 -> 64, CTA-built A rows)  ->  conv3d_tc x (L-2) (stride 2)  ->  conv3d_tc (stride 1, fp32 out)
//   ->  global mean / max over the grid  [-> sigmoid].
#include <cmath>
#include <cstring>

#include "internal.h"

using namespace a3d;

namespace {

struct Layer3d {
  int cin = 0, cout = 0, cout_pad = 0, stride = 0, g_in = 0, g_out = 0;
  bool bn = false, out_f32 = false;
  int w_index = -1, bn_tile = 0;
  Conv3dGeom g;
  CUtensorMap tmap_act, tmap_wgt;
  void* wgt = nullptr;
  float *scale = nullptr, *shift = nullptr;
  void* out = nullptr;          // arena buffer holding this layer's output
  size_t out_elems = 0;         // per object (padded channels)
};

inline int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }
inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

}  // namespace

struct a3d_enc3d {
  a3d_enc3d_desc desc{};
  int num_sms = 0;
  std::vector<Layer3d> layers;
  std::vector<std::vector<float>> w;
  std::vector<int64_t> w_numel;
  std::vector<bool> w_set;
  bool dirty = true;
  int64_t alloc_n = 0;
  size_t arena_bytes = 0;
  int64_t launches = 0;
  int sticky = 0;
};

namespace {

int check(const a3d_enc3d* h) {
  cudaGetLastError();   // drop a stale non-sticky error left by another runtime user (see check_handle in handle.cu)
  if (!h) { set_error("null voxel-encoder handle"); return A3D_ERR_INVALID; }
  if (h->sticky) { set_error("voxel-encoder handle is in a sticky CUDA error state (%d)", h->sticky); return h->sticky; }
  return A3D_OK;
}
int sticky(a3d_enc3d* h, int rc) {
  if (rc == A3D_ERR_CUDA) h->sticky = rc;
  return rc;
}

int build_plan(a3d_enc3d* h) {
  const a3d_enc3d_desc& d = h->desc;
  const int L = d.num_layers;
  bool ok = L >= 3 && L <= A3D_MAX_LAYERS && d.in_grid >= 32 && (d.in_grid & (d.in_grid - 1)) == 0;
  for (int i = 0; ok && i < L; ++i)
    ok = d.ksizes[i] == 4 && d.strides[i] == (i < L - 1 ? 2 : 1) && d.filters[i] >= 1;
  ok = ok && d.filters[0] == 64;
  for (int i = 1; ok && i < L - 1; ++i) ok = d.filters[i] % 128 == 0;
  if (ok) ok = (d.in_grid >> (L - 1)) >= 1;
  if (!ok) {
    set_error("unsupported encoder3D structure: this build implements filter_size_list [4]*L, strides_list [2]*(L-1)+[1], "
              "a cubic power-of-two one-channel input grid >= 32, filter_num_list[0] == 64 and hidden filter counts that "
              "are multiples of 128 (the reference: [64,128,256,512,2*latent], autoencoder3D.py:5-14)");
    return A3D_ERR_INVALID;
  }
  if (d.activation < 0 || d.activation > A3D_ACT_LRELU || d.final_activation < 0 || d.final_activation > 1 ||
      d.final_pool < 0 || d.final_pool > 2) {
    set_error("invalid activation / final_activation / final_pool field");
    return A3D_ERR_INVALID;
  }
  int g = d.in_grid, c = 1, wi = 0;
  for (int i = 0; i < L; ++i) {
    Layer3d ly;
    ly.cin = c; ly.cout = d.filters[i]; ly.stride = d.strides[i];
    ly.g_in = g; ly.g_out = g / ly.stride;
    ly.bn = i < L - 1; ly.out_f32 = i == L - 1;
    ly.cout_pad = i == 0 ? 64 : round_up(ly.cout, 128);
    ly.w_index = wi;
    h->w_numel.push_back((int64_t)64 * c * ly.cout);
    if (ly.bn) for (int k = 0; k < 4; ++k) h->w_numel.push_back(ly.cout);
    wi += ly.bn ? 5 : 1;
    if (i > 0) {
      ly.bn_tile = ly.cout_pad % 256 == 0 ? 256 : 128;
      Conv3dGeom& q = ly.g;
      const int G = ly.g_out;
      const int bw = G < 16 ? G : 16;
      int bh = 128 / bw; if (bh > G) bh = G;
      int bd = 128 / (bw * bh); if (bd > G) bd = G;
      q.G = G; q.stride = ly.stride;
      q.lw = ilog2(bw); q.lh = ilog2(bh); q.ld = ilog2(bd);
      q.tiles_w = G / bw; q.tiles_h = G / bh; q.tiles_d = G / bd;
      q.cin_chunks = ly.cin / 64;
      q.cout_pad = ly.cout_pad; q.cout_real = ly.cout;
      q.n_tiles = ly.cout_pad / ly.bn_tile;
    }
    ly.out_elems = (size_t)ly.g_out * ly.g_out * ly.g_out * (ly.out_f32 ? ly.cout : ly.cout_pad);
    h->layers.push_back(ly);
    g = ly.g_out; c = ly.cout;
  }
  h->w.assign(h->w_numel.size(), {});
  h->w_set.assign(h->w_numel.size(), false);
  return A3D_OK;
}

int make_maps(a3d_enc3d* h, int li) {
  Layer3d& ly = h->layers[li];
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return A3D_ERR_CUDA; }
  const CUtensorMapDataType dt =
      h->desc.operand_dtype == A3D_DTYPE_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const uint64_t C = ly.cin, Gi = ly.g_in;
  const cuuint32_t s = (cuuint32_t)ly.stride;
  cuuint64_t dims[5] = {C, Gi, Gi, Gi, (cuuint64_t)h->alloc_n};              // (c, w, h, d, n)
  cuuint64_t strides[4] = {C * 2, Gi * C * 2, Gi * Gi * C * 2, Gi * Gi * Gi * C * 2};
  // box = the input voxels s*o + k - 1 of one output brick: traversal extent s * brick with element stride s
  cuuint32_t box[5] = {64, (1u << ly.g.lw) * s, (1u << ly.g.lh) * s, (1u << ly.g.ld) * s,
                       128u >> (ly.g.lw + ly.g.lh + ly.g.ld)};
  cuuint32_t es[5] = {1, s, s, s, 1};
  CUresult r = enc(&ly.tmap_act, dt, 5, h->layers[li - 1].out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(voxel-encoder activations, layer %d) failed: %d", li, (int)r); return A3D_ERR_CUDA; }
  cuuint64_t wd[2] = {C, (cuuint64_t)64 * ly.cout_pad};
  cuuint64_t ws[1] = {C * 2};
  cuuint32_t wb[2] = {64, (cuuint32_t)ly.bn_tile};
  cuuint32_t e2[2] = {1, 1};
  r = enc(&ly.tmap_wgt, dt, 2, ly.wgt, wd, ws, wb, e2, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(voxel-encoder weights, layer %d) failed: %d", li, (int)r); return A3D_ERR_CUDA; }
  return A3D_OK;
}

int finalize(a3d_enc3d* h) {
  if (!h->dirty) return A3D_OK;
  for (size_t i = 0; i < h->w_set.size(); ++i)
    if (!h->w_set[i]) { set_error("voxel-encoder weight %zu of %zu was never set", i, h->w_set.size()); return A3D_ERR_WEIGHTS; }
  const int fmt = h->desc.operand_dtype;
  int rc;
  std::vector<float> sc, sf;
  std::vector<uint16_t> p16;
  for (size_t li = 0; li < h->layers.size(); ++li) {
    Layer3d& ly = h->layers[li];
    const std::vector<float>& k = h->w[ly.w_index];       // Keras Conv3D kernel [tap 64][cin][cout]
    if (ly.bn) fold_bn(h->w[ly.w_index + 1], h->w[ly.w_index + 2], h->w[ly.w_index + 3], h->w[ly.w_index + 4], sc, sf);
    else { sc.assign(ly.cout, 1.f); sf.assign(ly.cout, 0.f); }
    sc.resize(ly.cout_pad, 0.f);
    sf.resize(ly.cout_pad, 0.f);
    if ((rc = upload(sc.data(), sc.size() * 4, (void**)&ly.scale))) return rc;
    if ((rc = upload(sf.data(), sf.size() * 4, (void**)&ly.shift))) return rc;
    if (li == 0) {
      p16.assign(64 * 64, cvt16(0.f, fmt));                // [co][tap]
      for (int t = 0; t < 64; ++t)
        for (int co = 0; co < 64; ++co) p16[(size_t)co * 64 + t] = cvt16(k[(size_t)t * 64 + co], fmt);
      if ((rc = upload(p16.data(), p16.size() * 2, &ly.wgt))) return rc;
    } else {
      p16.assign((size_t)64 * ly.cout_pad * ly.cin, cvt16(0.f, fmt));
      for (int t = 0; t < 64; ++t)
        for (int ci = 0; ci < ly.cin; ++ci) {
          const float* src = &k[((size_t)t * ly.cin + ci) * ly.cout];
          for (int co = 0; co < ly.cout; ++co) p16[((size_t)t * ly.cout_pad + co) * ly.cin + ci] = cvt16(src[co], fmt);
        }
      if ((rc = upload(p16.data(), p16.size() * 2, &ly.wgt))) return rc;
      if ((rc = make_maps(h, (int)li))) return rc;
    }
  }
  h->dirty = false;
  return A3D_OK;
}

int run_chunk(a3d_enc3d* h, const float* vox, int64_t n, float* out_dev, cudaStream_t st) {
  const int fmt = h->desc.operand_dtype, act = h->desc.activation;
  int rc;
  Layer3d& l0 = h->layers[0];
  if ((rc = launch_conv3d_first_tc(vox, l0.wgt, l0.scale, l0.shift, l0.out, n, l0.g_in, fmt, act, h->num_sms, st, &h->launches)))
    return rc;
  for (size_t li = 1; li < h->layers.size(); ++li) {
    Layer3d& ly = h->layers[li];
    Conv3dGeom g = ly.g;
    const int nt = 128 >> (g.lw + g.lh + g.ld);
    g.n_objects = (int)n;
    g.m_tiles = (int)((n + nt - 1) / nt) * g.tiles_w * g.tiles_h * g.tiles_d;
    const bool last = li + 1 == h->layers.size();
    void* dst = (last && h->desc.final_pool == A3D_POOL_NONE) ? (void*)out_dev : ly.out;
    if ((rc = launch_conv3d_tc(ly.tmap_act, ly.tmap_wgt, dst, ly.scale, ly.shift, g, ly.bn_tile, fmt,
                               last ? A3D_ACT_NONE : act, ly.out_f32, h->num_sms, st, &h->launches)))
      return rc;
  }
  const Layer3d& lz = h->layers.back();
  const int64_t vox_out = (int64_t)lz.g_out * lz.g_out * lz.g_out;
  int64_t out_elems = n * lz.cout * (h->desc.final_pool == A3D_POOL_NONE ? vox_out : 1);
  if (h->desc.final_pool != A3D_POOL_NONE)
    if ((rc = launch_global_pool(reinterpret_cast<const float*>(lz.out), out_dev, n, (int)vox_out, lz.cout,
                                 h->desc.final_pool == A3D_POOL_MAX, st, &h->launches)))
      return rc;
  if (h->desc.final_activation == A3D_FINAL_SIGMOID)
    if ((rc = launch_sigmoid_inplace(out_dev, out_elems, st, &h->launches))) return rc;
  return A3D_OK;
}

}  // namespace

extern "C" {

int a3d_enc3d_create(const a3d_enc3d_desc* d, a3d_enc3d** out) {
  if (!d || !out) { set_error("null argument"); return A3D_ERR_INVALID; }
  *out = nullptr;
  if (d->abi_version != A3D_ABI_VERSION) { set_error("ABI version mismatch: %d vs %d", d->abi_version, A3D_ABI_VERSION); return A3D_ERR_INVALID; }
  if (d->operand_dtype != A3D_DTYPE_F16 && d->operand_dtype != A3D_DTYPE_BF16) { set_error("invalid operand dtype"); return A3D_ERR_INVALID; }
  if (d->max_batch < 1) { set_error("max_batch must be >= 1"); return A3D_ERR_INVALID; }
  a3d_enc3d* h = new a3d_enc3d();
  h->desc = *d;
  int rc = build_plan(h);
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
  int ntmax = 1;   // TMA boxes hold 128 / brick objects: keep them inside the allocation
  for (size_t li = 1; li < h->layers.size(); ++li) {
    const Conv3dGeom& g = h->layers[li].g;
    const int nt = 128 >> (g.lw + g.lh + g.ld);
    if (nt > ntmax) ntmax = nt;
  }
  h->alloc_n = round_up(d->max_batch, ntmax);
  for (auto& ly : h->layers) {
    const size_t bytes = ly.out_elems * (ly.out_f32 ? 4 : 2) * (size_t)h->alloc_n;
    cudaError_t e = cudaMalloc(&ly.out, bytes);
    if (e != cudaSuccess) {
      set_error("voxel-encoder arena allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
      a3d_enc3d_destroy(h);
      return A3D_ERR_CUDA;
    }
    cudaMemset(ly.out, 0, bytes);
    h->arena_bytes += bytes;
  }
  *out = h;
  return A3D_OK;
}

void a3d_enc3d_destroy(a3d_enc3d* h) {
  if (!h) return;
  cudaDeviceSynchronize();
  for (auto& ly : h->layers) { cudaFree(ly.out); cudaFree(ly.wgt); cudaFree(ly.scale); cudaFree(ly.shift); }
  delete h;
}

int a3d_enc3d_num_weights(const a3d_enc3d* h) { return h ? (int)h->w_numel.size() : 0; }
int64_t a3d_enc3d_weight_numel(const a3d_enc3d* h, int index) {
  return (h && index >= 0 && index < (int)h->w_numel.size()) ? h->w_numel[index] : -1;
}

int a3d_enc3d_set_weight(a3d_enc3d* h, int index, const float* host, size_t nbytes) {
  int rc = check(h);
  if (rc) return rc;
  if (index < 0 || index >= (int)h->w_numel.size() || !host) { set_error("bad voxel-encoder weight index %d", index); return A3D_ERR_INVALID; }
  if (nbytes != (size_t)h->w_numel[index] * 4) {
    set_error("voxel-encoder weight %d: expected %lld fp32 values, got %zu bytes", index, (long long)h->w_numel[index], nbytes);
    return A3D_ERR_WEIGHTS;
  }
  h->w[index].assign(host, host + h->w_numel[index]);
  h->w_set[index] = true;
  h->dirty = true;
  return A3D_OK;
}

int a3d_enc3d_get_weight(const a3d_enc3d* h, int index, float* host, size_t nbytes) {
  if (!h || index < 0 || index >= (int)h->w_numel.size() || !host) { set_error("bad voxel-encoder weight index %d", index); return A3D_ERR_INVALID; }
  if (!h->w_set[index]) { set_error("voxel-encoder weight %d was never set", index); return A3D_ERR_WEIGHTS; }
  if (nbytes != (size_t)h->w_numel[index] * 4) { set_error("voxel-encoder weight %d: size mismatch", index); return A3D_ERR_WEIGHTS; }
  memcpy(host, h->w[index].data(), nbytes);
  return A3D_OK;
}

int a3d_enc3d_forward(a3d_enc3d* h, const float* voxels_dev, int64_t n, float* out_dev, void* stream) {
  int rc = check(h);
  if (rc) return rc;
  if (n < 0 || (n > 0 && (!voxels_dev || !out_dev))) { set_error("a3d_enc3d_forward: bad arguments"); return A3D_ERR_INVALID; }
  if ((rc = sticky(h, finalize(h)))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t G = h->desc.in_grid, in_per = G * G * G;
  const Layer3d& lz = h->layers.back();
  const int64_t out_per = (int64_t)lz.cout * (h->desc.final_pool == A3D_POOL_NONE ? (int64_t)lz.g_out * lz.g_out * lz.g_out : 1);
  for (int64_t off = 0; off < n; off += h->desc.max_batch) {
    const int64_t nc = n - off < h->desc.max_batch ? n - off : h->desc.max_batch;
    if ((rc = sticky(h, run_chunk(h, voxels_dev + off * in_per, nc, out_dev + off * out_per, st)))) return rc;
  }
  return A3D_OK;
}

int a3d_enc3d_split_sample(a3d_enc3d* h, const float* enc_out_dev, int64_t n, int D, int out_stride, float clip,
                           int seed_enable, uint64_t seed, uint64_t obj_offset, float* mean_dev, float* logvar_dev,
                           float* z_dev, void* stream) {
  int rc = check(h);
  if (rc) return rc;
  if (n < 0 || D < 1 || out_stride < 2 * D || (n > 0 && !enc_out_dev)) { set_error("a3d_enc3d_split_sample: bad arguments"); return A3D_ERR_INVALID; }
  return sticky(h, launch_split_sample(enc_out_dev, n, D, out_stride, clip, seed_enable, seed, obj_offset, mean_dev,
                                       logvar_dev, z_dev, (cudaStream_t)stream, &h->launches));
}

int a3d_enc3d_debug_read_layer(a3d_enc3d* h, int layer, int64_t n, float* host, size_t nbytes) {
  int rc = check(h);
  if (rc) return rc;
  if (layer < 0 || layer >= (int)h->layers.size() - 1 || n <= 0 || n > h->desc.max_batch || !host) {
    set_error("a3d_enc3d_debug_read_layer: bad arguments");
    return A3D_ERR_INVALID;
  }
  const Layer3d& ly = h->layers[layer];
  const size_t elems = ly.out_elems * (size_t)n;
  if (nbytes != elems * 4) { set_error("a3d_enc3d_debug_read_layer: expected %zu bytes", elems * 4); return A3D_ERR_INVALID; }
  float* tmp = nullptr;
  A3D_CUDA_OK(cudaMalloc(&tmp, nbytes));
  A3D_CUDA_OK(cudaDeviceSynchronize());
  rc = launch_to_f32(ly.out, tmp, (int64_t)elems, h->desc.operand_dtype, 0);
  if (rc == A3D_OK) {
    cudaError_t e = cudaMemcpy(host, tmp, nbytes, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_error("debug copy failed: %s", cudaGetErrorString(e)); rc = A3D_ERR_CUDA; }
  }
  cudaFree(tmp);
  return rc;
}

int64_t a3d_enc3d_launch_count(const a3d_enc3d* h) { return h ? h->launches : 0; }
size_t a3d_enc3d_workspace_bytes(const a3d_enc3d* h) { return h ? h->arena_bytes : 0; }

}  // extern "C"
