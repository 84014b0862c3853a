// Memory-bound helper kernels of the anytime path:
//   impute   -- sampling() + mask / fill logic, function.py:35-38, nolbo.py:1472-1486,1505-1510,431-439
//   counts   -- voxelPrecisionRecall(), function.py:100-115 (stand-alone form on fp32 grids)
//   pack     -- fp32 {0,1} targets (loader layout, pascal3D.py:149-152) -> 1 bit / voxel
#include "internal.h"
#include "philox.cuh"
#include "ptx.cuh"

namespace a3d {
namespace {

// One block per object.  z, mask: [B, D]; mu: [C, D]; z_out: [B, K, D].
__global__ void __launch_bounds__(128)
impute_kernel(const float* __restrict__ z, const float* __restrict__ mask, const float* __restrict__ mu, int C, int K,
              int D, uint64_t seed, uint64_t obj_offset, int fill, float* __restrict__ z_out,
              int32_t* __restrict__ cstar_out) {
  extern __shared__ float sm[];
  float* zf = sm;            // [D] filled latent
  float* mk = zf + D;        // [D] mask
  float* dist = mk + D;      // [C]
  __shared__ int cstar_s;
  const int64_t b = blockIdx.x;
  const uint64_t obj = obj_offset + (uint64_t)b;
  ptx::pdl_sync();     // first kernel of a call's chain: lets the Dense kernel's launch overlap this one
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float m = mask[b * D + d];
    float v = z[b * D + d] * m;                       // nolbo.py:1477
    if (fill != A3D_FILL_NORMAL && fill != A3D_FILL_NONE && v == 0.f) {   // nolbo.py:1481-1482: where(z == 0) <- mean_c(mu)
      float s = 0.f;
      for (int c = 0; c < C; ++c) s += mu[(size_t)c * D + d];
      v = s / (float)C;
    }
    zf[d] = v;
    mk[d] = m;
  }
  __syncthreads();
  if (fill != A3D_FILL_NORMAL && fill != A3D_FILL_NONE) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {   // nolbo.py:1505: sum_d mask * (z - mu_c)^2
      float s = 0.f;
      for (int d = 0; d < D; ++d) {
        const float t = zf[d] - mu[(size_t)c * D + d];
        s = fmaf(mk[d] * t, t, s);
      }
      dist[c] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int best = 0;
      for (int c = 1; c < C; ++c)
        if (dist[c] < dist[best]) best = c;             // tf.argmin: first minimum
      cstar_s = best;
      if (cstar_out) cstar_out[b] = best;
    }
    __syncthreads();
  } else if (threadIdx.x == 0 && cstar_out) {
    cstar_out[b] = -1;
  }
  const int nq = (D + 3) >> 2;
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  for (int i = threadIdx.x; i < K * nq; i += blockDim.x) {
    const int k = i / nq, q = i % nq;
    float nrm[4] = {0.f, 0.f, 0.f, 0.f};
    if (fill != A3D_FILL_MEAN && fill != A3D_FILL_NONE) {
      const uint4 w = philox4x32_10(make_uint4((uint32_t)q, (uint32_t)k, (uint32_t)obj, (uint32_t)(obj >> 32)), key);
      box_muller(w.x, w.y, nrm[0], nrm[1]);
      box_muller(w.z, w.w, nrm[2], nrm[3]);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int d = q * 4 + e;
      if (d >= D) break;
      float v = zf[d];
      if (fill == A3D_FILL_PRIOR_SAMPLE) {
        if (mk[d] == 0.f) v = mu[(size_t)cstar_s * D + d] + nrm[e];   // nolbo.py:1508-1510, sigma = 1
      } else if (fill == A3D_FILL_NORMAL) {
        if (v == 0.f) v = nrm[e];                                      // nolbo.py:437-439
      }
      A3D_DEV_CHECK(k >= 0 && k < K && d >= 0 && d < D && (fill != A3D_FILL_PRIOR_SAMPLE || (unsigned)cstar_s < (unsigned)C));
      z_out[((size_t)b * K + k) * D + d] = v;
    }
  }
}

// sampling(mu, logVar), function.py:35-38: z = mu + sqrt(exp(logVar)) * eps, eps = Philox4x32-10 + Box-Muller with the
// counter (dim / 4, 0, obj_offset + row) of the imputation sampler's k = 0 draw.  One thread per (row, 4 dims).
__global__ void __launch_bounds__(256)
sampling_kernel(const float* __restrict__ mu, const float* __restrict__ logvar, int64_t n, int D, uint64_t seed,
                uint64_t obj_offset, float* __restrict__ z) {
  const int nq = (D + 3) >> 2;
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n * nq; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / nq;
    const int q = (int)(i % nq);
    const uint64_t obj = obj_offset + (uint64_t)b;
    const uint4 w = philox4x32_10(make_uint4((uint32_t)q, 0u, (uint32_t)obj, (uint32_t)(obj >> 32)), key);
    float nrm[4];
    box_muller(w.x, w.y, nrm[0], nrm[1]);
    box_muller(w.z, w.w, nrm[2], nrm[3]);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int d = q * 4 + e;
      if (d < D) z[b * D + d] = mu[b * D + d] + sqrtf(expf(logvar[b * D + d])) * nrm[e];
    }
  }
}

// Nearest-prior classification of getEval, nolbo.py:1488-1494 / :1511-1518: idx[b] = argmin_c sum_d (z[b,d] - mu[c,d])^2
// (first minimum, like tf.argmin), hit = (idx[b] == argmax_c labels[b,c]) (first maximum, like tf.argmax); *hits counts
// the matches of the batch (acc_cat = hits / B).  One warp per object; z rows are z_stride floats apart.
__global__ void __launch_bounds__(128)
nearest_prior_kernel(const float* __restrict__ z, int64_t z_stride, const float* __restrict__ mu, int C, int D,
                     const float* __restrict__ labels, int64_t B, int32_t* __restrict__ idx_out,
                     int32_t* __restrict__ hits) {
  const int lane = threadIdx.x & 31;
  const int64_t b = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const float* zb = z + b * z_stride;
  float best = INFINITY;
  int best_c = 0x7fffffff;
  for (int c = lane; c < C; c += 32) {
    float s = 0.f;
    for (int d = 0; d < D; ++d) {
      const float t = zb[d] - mu[(size_t)c * D + d];
      s = fmaf(t, t, s);
    }
    if (s < best) { best = s; best_c = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oc = __shfl_xor_sync(0xffffffffu, best_c, o);
    if (ob < best || (ob == best && oc < best_c)) { best = ob; best_c = oc; }
  }
  if (labels) {
    float lb = -INFINITY;
    int lc = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {
      const float v = labels[b * C + c];
      if (v > lb) { lb = v; lc = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, lb, o);
      const int oc = __shfl_xor_sync(0xffffffffu, lc, o);
      if (ob > lb || (ob == lb && oc < lc)) { lb = ob; lc = oc; }
    }
    if (lane == 0 && hits && lc == best_c) atomicAdd(hits, 1);
  }
  if (lane == 0 && idx_out) idx_out[b] = best_c;
}

// fp32 probability grid -> fp16 grid / 1 bit per voxel (p >= thr), the compact return formats of a3d_decode_host
__global__ void __launch_bounds__(256)
grid_to_f16_kernel(const float* __restrict__ in, int64_t n4, uint2* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(in) + i);
    const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    out[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
  }
}
__global__ void __launch_bounds__(256)
grid_to_bits_kernel(const float* __restrict__ in, int64_t total_bytes, float thr, uint8_t* __restrict__ bits) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total_bytes;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = __ldcs(reinterpret_cast<const float4*>(in) + 2 * i);
    const float4 b = __ldcs(reinterpret_cast<const float4*>(in) + 2 * i + 1);
    const uint32_t v = (a.x >= thr) | ((a.y >= thr) << 1) | ((a.z >= thr) << 2) | ((a.w >= thr) << 3) |
                       ((b.x >= thr) << 4) | ((b.y >= thr) << 5) | ((b.z >= thr) << 6) | ((b.w >= thr) << 7);
    bits[i] = (uint8_t)v;
  }
}

// counts[b] += {TP, FP, FN};  grid = (chunks, B).  V % 4 == 0.
__global__ void __launch_bounds__(256)
counts_kernel(const float* __restrict__ target, const float* __restrict__ pred, int64_t V, float thr,
              unsigned long long* __restrict__ counts) {
  const int64_t b = blockIdx.y;
  const float4* t4 = reinterpret_cast<const float4*>(target + b * V);
  const float4* p4 = reinterpret_cast<const float4*>(pred + b * V);
  int tp = 0, fp = 0, fn = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < V / 4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 t = __ldg(t4 + i), p = __ldg(p4 + i);
    const float tv[4] = {t.x, t.y, t.z, t.w}, pv[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int y = pv[e] >= thr, g = tv[e] > 0.5f;
      tp += g & y;
      fp += (1 - g) & y;
      fn += g & (1 - y);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    tp += __shfl_xor_sync(0xffffffffu, tp, o);
    fp += __shfl_xor_sync(0xffffffffu, fp, o);
    fn += __shfl_xor_sync(0xffffffffu, fn, o);
  }
  __shared__ int red[3][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = tp; red[1][warp] = fp; red[2][warp] = fn; }
  __syncthreads();
  if (threadIdx.x < 3) {
    int s = 0;
    for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
    if (s) atomicAdd(counts + b * 3 + threadIdx.x, (unsigned long long)s);
  }
}

// binary_loss (function.py:73-82, b_range = False): loss[b] = -sum_v gamma*y*log(p) + (1-gamma)*(1-y)*log(1-p),
// p = clip(pred, 1e-7, 1-1e-7) in fp32.  grid = (chunks, B).
__global__ void __launch_bounds__(256)
binary_loss_kernel(const float* __restrict__ pred, const float* __restrict__ target, int64_t V, float gamma,
                   double* __restrict__ loss) {
  const int64_t b = blockIdx.y;
  const float4* t4 = reinterpret_cast<const float4*>(target + b * V);
  const float4* p4 = reinterpret_cast<const float4*>(pred + b * V);
  float acc = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < V / 4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 t = __ldg(t4 + i), p = __ldg(p4 + i);
    const float tv[4] = {t.x, t.y, t.z, t.w}, pv[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float pc = fminf(fmaxf(pv[e], 1e-7f), 1.f - 1e-7f);
      acc -= gamma * tv[e] * logf(pc) + (1.f - gamma) * (1.f - tv[e]) * logf(1.f - pc);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(loss + b, (double)acc);
}

// Threshold sweep (modelnetAE3.ipynb cell 2): counts[b][t] += {TP, FP, FN} for T <= 32 thresholds.
struct SweepThr { float v[32]; };
__global__ void __launch_bounds__(256)
counts_sweep_kernel(const float* __restrict__ target, const float* __restrict__ pred, int64_t V, SweepThr thr, int T,
                    int strict, unsigned long long* __restrict__ counts) {
  const int64_t b = blockIdx.y;
  const float4* t4 = reinterpret_cast<const float4*>(target + b * V);
  const float4* p4 = reinterpret_cast<const float4*>(pred + b * V);
  __shared__ int red[32][3];
  for (int i = threadIdx.x; i < 96; i += blockDim.x) red[i / 3][i % 3] = 0;
  __syncthreads();
  for (int t = 0; t < T; ++t) {
    const float th = thr.v[t];
    int tp = 0, fp = 0, fn = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < V / 4; i += (int64_t)gridDim.x * blockDim.x) {
      const float4 tt = __ldg(t4 + i), pp = __ldg(p4 + i);   // re-read per threshold: the chunk stays in L1/L2
      const float tv[4] = {tt.x, tt.y, tt.z, tt.w}, pv[4] = {pp.x, pp.y, pp.z, pp.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int y = strict ? (pv[e] > th) : (pv[e] >= th), g = tv[e] > 0.5f;
        tp += g & y;
        fp += (1 - g) & y;
        fn += g & (1 - y);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      tp += __shfl_xor_sync(0xffffffffu, tp, o);
      fp += __shfl_xor_sync(0xffffffffu, fp, o);
      fn += __shfl_xor_sync(0xffffffffu, fn, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&red[t][0], tp); atomicAdd(&red[t][1], fp); atomicAdd(&red[t][2], fn); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * 3; i += blockDim.x)
    if (red[i / 3][i % 3]) atomicAdd(counts + (b * T + i / 3) * 3 + i % 3, (unsigned long long)red[i / 3][i % 3]);
}

// bits[i] packs voxels 8i..8i+7 (bit e = voxel 8i + e).  total = B*V/8 bytes.
__global__ void pack_kernel(const float* __restrict__ target, int64_t total_bytes, uint8_t* __restrict__ bits) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total_bytes;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(target) + 2 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(target) + 2 * i + 1);
    uint32_t v = (a.x > 0.5f) | ((a.y > 0.5f) << 1) | ((a.z > 0.5f) << 2) | ((a.w > 0.5f) << 3) | ((b.x > 0.5f) << 4) |
                 ((b.y > 0.5f) << 5) | ((b.z > 0.5f) << 6) | ((b.w > 0.5f) << 7);
    bits[i] = (uint8_t)v;
  }
}

}  // namespace

int launch_sampling(const float* mu, const float* logvar, int64_t n, int D, uint64_t seed, uint64_t obj_offset, float* z,
                    cudaStream_t st, int64_t* launches) {
  if (n <= 0) return A3D_OK;
  const int64_t work = n * ((D + 3) / 4);
  const int grid = (int)((work + 255) / 256 < 148 * 8 ? (work + 255) / 256 : 148 * 8);
  sampling_kernel<<<grid, 256, 0, st>>>(mu, logvar, n, D, seed, obj_offset, z);
  A3D_CUDA_OK(cudaGetLastError());
  if (launches) ++*launches;
  return A3D_OK;
}

int launch_nearest_prior(const float* z, int64_t z_stride, const float* mu, int C, int D, const float* labels, int64_t B,
                         int32_t* idx, int32_t* hits, cudaStream_t st, int64_t* launches) {
  if (B <= 0) return A3D_OK;
  nearest_prior_kernel<<<(unsigned)((B + 3) / 4), 128, 0, st>>>(z, z_stride, mu, C, D, labels, B, idx, hits);
  A3D_CUDA_OK(cudaGetLastError());
  if (launches) ++*launches;
  return A3D_OK;
}

int launch_grid_convert(const float* in, int64_t voxels, int out_dtype, float thr, void* out, cudaStream_t st,
                        int64_t* launches) {
  if (voxels <= 0) return A3D_OK;
  const int64_t work = out_dtype == A3D_OUT_F16 ? voxels / 4 : voxels / 8;
  const int grid = (int)((work + 255) / 256 < 148 * 16 ? (work + 255) / 256 : 148 * 16);
  if (out_dtype == A3D_OUT_F16)
    grid_to_f16_kernel<<<grid, 256, 0, st>>>(in, work, reinterpret_cast<uint2*>(out));
  else
    grid_to_bits_kernel<<<grid, 256, 0, st>>>(in, work, thr, reinterpret_cast<uint8_t*>(out));
  A3D_CUDA_OK(cudaGetLastError());
  if (launches) ++*launches;
  return A3D_OK;
}

int launch_impute(const float* z, const float* mask, const float* mu, int C, int64_t B, int K, int D, uint64_t seed,
                  uint64_t obj_offset, int fill, float* z_out, int32_t* cstar, cudaStream_t st, int64_t* launches) {
  if (B <= 0) return A3D_OK;
  const size_t smem = (size_t)(2 * D + (C > 0 ? C : 1)) * sizeof(float);
  if (smem > 48 * 1024) {
    set_error("a3d_impute: latent_dim %d / %d categories need %zu bytes of shared memory (limit 48 KB)", D, C, smem);
    return A3D_ERR_INVALID;
  }
  A3D_CUDA_OK(launch_chain(impute_kernel, dim3((unsigned)B), dim3(128), smem, st, 1, z, mask, mu, C, K, D, seed, obj_offset,
                           fill, z_out, cstar));
  if (launches) ++*launches;
  return A3D_OK;
}

int launch_counts(const float* target, const float* pred, int64_t B, int64_t V, float thr, unsigned long long* counts,
                  cudaStream_t st, int64_t* launches) {
  if (B <= 0) return A3D_OK;
  int chunks = (int)((V / 4 + 256 * 8 - 1) / (256 * 8));
  if (chunks < 1) chunks = 1;
  dim3 grid(chunks, (unsigned)B);
  counts_kernel<<<grid, 256, 0, st>>>(target, pred, V, thr, counts);
  A3D_CUDA_OK(cudaGetLastError());
  if (launches) ++*launches;
  return A3D_OK;
}

int launch_binary_loss(const float* pred, const float* target, int64_t B, int64_t V, float gamma, double* loss,
                       cudaStream_t st, int64_t* launches) {
  if (B <= 0) return A3D_OK;
  int chunks = (int)((V / 4 + 256 * 8 - 1) / (256 * 8));
  if (chunks < 1) chunks = 1;
  binary_loss_kernel<<<dim3(chunks, (unsigned)B), 256, 0, st>>>(pred, target, V, gamma, loss);
  A3D_CUDA_OK(cudaGetLastError());
  if (launches) ++*launches;
  return A3D_OK;
}

int launch_counts_sweep(const float* target, const float* pred, int64_t B, int64_t V, const float* thr, int T, int strict,
                        unsigned long long* counts, cudaStream_t st, int64_t* launches) {
  if (B <= 0 || T <= 0) return A3D_OK;
  SweepThr tv;
  for (int i = 0; i < 32; ++i) tv.v[i] = i < T ? thr[i] : 0.f;
  int chunks = (int)((V / 4 + 256 * 8 - 1) / (256 * 8));
  if (chunks < 1) chunks = 1;
  counts_sweep_kernel<<<dim3(chunks, (unsigned)B), 256, 0, st>>>(target, pred, V, tv, T, strict, counts);
  A3D_CUDA_OK(cudaGetLastError());
  if (launches) ++*launches;
  return A3D_OK;
}

int launch_pack(const float* target, int64_t B, int64_t V, uint8_t* bits, cudaStream_t st, int64_t* launches) {
  const int64_t total = B * V / 8;
  if (total <= 0) return A3D_OK;
  const int grid = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
  pack_kernel<<<grid, 256, 0, st>>>(target, total, bits);
  A3D_CUDA_OK(cudaGetLastError());
  if (launches) ++*launches;
  return A3D_OK;
}

}  // namespace a3d
