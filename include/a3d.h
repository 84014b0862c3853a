/*
 * a3d.h -- C ABI of the B200-native anytime voxel-decoder hot path.
 *
 * The reference (bogus2000/anytime-3D-reconstruction) is 100 % Python/TensorFlow: the path sits behind Python
 * callables, not an FFI.  Each entry point below names the reference interface it replaces (paths relative to
 * /root/reference).  All functions return 0 on success or a negative a3d_status; the message of the last error on
 * the calling thread is available from a3d_last_error().  CUDA errors are sticky on the handle.
 *
 * Ownership: the caller owns every input/output buffer; the library owns weights, tensor maps and one workspace
 * arena sized at create time.  No allocation happens inside the hot calls.  A handle is NOT thread-safe: use one
 * handle per (device, stream).  Every *_dev call is asynchronous on the stream passed in (cudaStream_t as void*).
 *
 * There is no CPU fallback: every compute entry point fails with A3D_ERR_NO_DEVICE when no sm_100 GPU is present.
 */
#ifndef A3D_H_
#define A3D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define A3D_ABI_VERSION 1
#define A3D_MAX_LAYERS 8
#define A3D_VOXELS 262144 /* 64^3, output grid of the reference decoder (autoencoder3D.py:17) */

typedef enum {
  A3D_OK = 0,
  A3D_ERR_INVALID = -1,     /* bad argument / unsupported structure */
  A3D_ERR_CUDA = -2,        /* CUDA runtime/driver error (sticky)    */
  A3D_ERR_NO_DEVICE = -3,   /* no sm_100 device: there is no CPU path */
  A3D_ERR_WEIGHTS = -4,     /* weights missing or of the wrong size  */
  A3D_ERR_WORKSPACE = -5    /* request exceeds the arena             */
} a3d_status;

enum { A3D_ACT_NONE = 0, A3D_ACT_ELU = 1, A3D_ACT_RELU = 2, A3D_ACT_LRELU = 3 /* slope 0.3 (Keras default) */,
       A3D_ACT_LRELU01 = 4 /* LeakyReLU(alpha=0.1), src/net_core/darknet.py:88,141 */ };
enum { A3D_FINAL_NONE = 0, A3D_FINAL_SIGMOID = 1 };
enum { A3D_DTYPE_F16 = 0, A3D_DTYPE_BF16 = 1 };          /* operand type of the tensor-core layers */
enum { A3D_IMPL_TCGEN05 = 0, A3D_IMPL_SIMT = 1 };        /* SIMT = CUDA-core kernels, bring-up/diagnostic only */
enum { A3D_FILL_PRIOR_SAMPLE = 0, A3D_FILL_MEAN = 1, A3D_FILL_NORMAL = 2,
       A3D_FILL_NONE = 3 /* z * mask only: getEval's missing_prob == 0 branch, nolbo.py:1485-1486 */ };
enum { A3D_OUT_F32 = 0, A3D_OUT_F16 = 1, A3D_OUT_BITS = 2 };   /* grid formats of a3d_decode_host */

/* Mirrors the `structure` dict consumed by decoder3D(structure), src/net_core/autoencoder3D.py:104-112. */
typedef struct {
  int32_t abi_version;               /* A3D_ABI_VERSION */
  int32_t latent_dim;                /* structure['input_dim'] */
  int32_t num_layers;                /* len(filter_num_list) */
  int32_t filters[A3D_MAX_LAYERS];   /* structure['filter_num_list'] */
  int32_t ksizes[A3D_MAX_LAYERS];    /* structure['filter_size_list'] */
  int32_t strides[A3D_MAX_LAYERS];   /* structure['strides_list'] */
  int32_t out_grid;                  /* structure['output_shape'][0] (cubic) */
  int32_t activation;                /* A3D_ACT_*   (structure['activation']) */
  int32_t final_activation;          /* A3D_FINAL_* (structure['final_activation']) */
  int32_t device;                    /* CUDA ordinal */
  int32_t max_chunk;                 /* decodes resident in the arena at once (multiple of 32) */
  int32_t operand_dtype;             /* A3D_DTYPE_* */
  int32_t impl;                      /* A3D_IMPL_* */
} a3d_desc;

typedef struct a3d_handle a3d_handle;

/* decoder3D(structure) -> model            src/net_core/autoencoder3D.py:104-139 (graph construction). */
int a3d_create(const a3d_desc* desc, a3d_handle** out);
void a3d_destroy(a3d_handle* h);

/* Number / element count of the Keras variables in model.get_weights() order (27 for the stock decoder). */
int a3d_num_weights(const a3d_handle* h);
int64_t a3d_weight_numel(const a3d_handle* h, int index);

/* model.set_weights()/get_weights()/load_weights()   src/module/nolbo.py:1572-1574,1585-1592.
 * `host` is fp32 in the Keras layout of that variable (Dense [D,512]; ConvT [k,k,k,Cout,Cin]; BN vectors).
 * Synchronous.  BN folding, operand conversion and per-tap repacking happen lazily before the next decode. */
int a3d_set_weight(a3d_handle* h, int index, const float* host, size_t nbytes);
int a3d_get_weight(const a3d_handle* h, int index, float* host, size_t nbytes);

/* decoder(z, training=False)               src/module/nolbo.py:1496,1520; src/module/nolbo_test.py:176,180.
 * z_dev: [n, latent_dim] fp32 on the device; prob_dev: [n, 64,64,64,1] fp32 (NDHWC, like the Keras output). */
int a3d_decode(a3d_handle* h, const float* z_dev, int64_t n, float* prob_dev, void* stream);

/* decoder(z, training=False) as the reference's scripts use it: numpy latents in, the full occupancy grid back on the
 * host (src/module/nolbo.py:1496 -> test_modelnet_VAE_dr.py:128-130 pulls the [72,64,64,64,1] prediction every iteration).
 * z_host: [n, latent_dim] fp32; out_host: [n, 262144] of out_dtype -- A3D_OUT_F32 (the Keras output), A3D_OUT_F16
 * (same values rounded to fp16, half the PCIe bytes) or A3D_OUT_BITS ([n, 32768] bytes, bit = p >= thr, packed like
 * the targets).  The call is processed in sub-chunks on internal streams: the device->host copy of sub-chunk i runs
 * while sub-chunk i+1 decodes (double-buffered staging), so with a PINNED out_host the call runs at the PCIe rate;
 * pageable buffers work too (the copies then serialise in the driver).  Synchronous. */
int a3d_decode_host(a3d_handle* h, const float* z_host, int64_t n, void* out_host, int out_dtype, float thr);

/* sampling(mu, logVar)                      src/module/function.py:35-38: z = mu + sqrt(exp(logVar)) * eps with
 * eps ~ N(0,1) from Philox4x32-10 + Box-Muller, counter (dim/4, 0, obj_offset + row), key = seed (the k = 0 stream of
 * a3d_impute).  mu_dev, logvar_dev, z_dev: [n, D] fp32 on the current device.  No handle needed. */
int a3d_sampling(const float* mu_dev, const float* logvar_dev, int64_t n, int D, uint64_t seed, uint64_t obj_offset,
                 float* z_dev, void* stream);

/* Nearest-prior classification of getEval   src/module/nolbo.py:1488-1494,1511-1518: idx[b] = argmin_c ||z_b - mu_c||^2
 * over ALL dims (tf.argmin: first minimum); with labels_dev ([B, C] one-hot category_list) *hits_dev receives the number
 * of b with idx[b] == argmax_c labels[b, c], i.e. acc_cat * B (overwritten).  z rows are z_stride floats apart (pass
 * K * D to classify the first of K completed latents per object).  idx_out_dev / labels_dev / hits_dev may be NULL. */
int a3d_nearest_prior(const float* z_dev, int64_t z_stride, const float* mu_table_dev, int C, int D,
                      const float* labels_dev, int64_t B, int32_t* idx_out_dev, int32_t* hits_dev, void* stream);

/* sampling() + mask / fill logic           src/module/function.py:35-38; src/module/nolbo.py:1472-1486 (mean fill),
 * :1505-1510 (nearest-prior sample fill), :431-439 (N(0,1) fill).  K Philox4x32-10 normal draws per object are
 * scattered into the missing dims.  z, mask: [B, D]; mu_table: [C, D] (category_vectors); z_out: [B, K, D];
 * cstar_out: [B] int32 or NULL.  Counter = (dim/4, k, obj_offset + b); key = seed. */
int a3d_impute(a3d_handle* h, const float* z_dev, const float* mask_dev, const float* mu_table_dev, int C,
               int64_t B, int K, uint64_t seed, uint64_t obj_offset, int fill_mode, float* z_out_dev,
               int32_t* cstar_out_dev, void* stream);

/* K-sample anytime reconstruction + scoring: decoder over all B*K completed latents, mean of the K post-sigmoid
 * grids (src/module/nolbo_test.py:167-177), threshold `>= thr` and TP/FP/FN (src/module/function.py:100-115).
 * z_bkd_dev: [B, K, D] fp32; target_bits_dev: [B, 32768] bytes, voxel v -> bit (v & 7) of byte v >> 3, v in NDHW
 * order; counts_dev: [B, 3] int64 (TP, FP, FN), overwritten; mean_prob_dev: [B, 262144] fp32 or NULL. */
int a3d_anytime_eval(a3d_handle* h, const float* z_bkd_dev, int64_t B, int K, const uint8_t* target_bits_dev,
                     float thr, int64_t* counts_dev, float* mean_prob_dev, void* stream);

/* a3d_anytime_eval plus the weighted-BCE shape loss the reference's getEval reports next to the counts:
 * binary_loss(xPred, xTarget, gamma)   src/module/function.py:73-82, called with gamma = 0.60 at src/module/nolbo.py:1497,1521.
 * loss_dev: [B] double = -sum_v (gamma*y*log(p) + (1-gamma)*(1-y)*log(1-p)), p = clip(mean grid, 1e-7, 1-1e-7); overwritten. */
int a3d_anytime_eval_loss(a3d_handle* h, const float* z_bkd_dev, int64_t B, int K, const uint8_t* target_bits_dev,
                          float thr, float gamma, int64_t* counts_dev, double* loss_dev, float* mean_prob_dev,
                          void* stream);

/* The four stand-alone scoring helpers below need no decoder: `h` may be NULL (the current CUDA device is used and
 * there is no launch counting / sticky state).
 *
 * binary_loss(xPred, xTarget, epsilon = 1e-7, gamma)   src/module/function.py:73-82 (b_range = False), stand-alone form on
 * fp32 grids [B, V]; loss_dev [B] double. */
int a3d_binary_loss(a3d_handle* h, const float* pred_dev, const float* target_dev, int64_t B, int64_t V, float gamma,
                    double* loss_dev, void* stream);

/* Precision/recall threshold sweep of the evaluation notebooks (modelnetAE3.ipynb cell 2): for each of T thresholds
 * yPred = strict ? (xPred > thr) : (xPred >= thr); counts_dev [B, T, 3] int64 (TP, FP, FN).  thresholds: host, T <= 32. */
int a3d_counts_sweep(a3d_handle* h, const float* target_dev, const float* pred_dev, int64_t B, int64_t V,
                     const float* thresholds, int T, int strict, int64_t* counts_dev, void* stream);

/* voxelPrecisionRecall(xTarget, xPred, prob)   src/module/function.py:100-115.  Both [B, V] fp32 on the device;
 * counts_dev [B, 3] int64 = TP, FP, FN with yPred = (xPred >= thr), yTarget = (xTarget > 0.5). */
int a3d_counts(a3d_handle* h, const float* target_dev, const float* pred_dev, int64_t B, int64_t V, float thr,
               int64_t* counts_dev, void* stream);

/* Bit-pack fp32 {0,1} targets [B, V] (loader layout, src/dataset_loader/pascal3D.py:149-152) to [B, V/8] bytes. */
int a3d_pack_targets(a3d_handle* h, const float* target_dev, int64_t B, int64_t V, uint8_t* bits_dev, void* stream);

/* End-to-end call with HOST buffers (what a getEval-style caller holds, src/module/nolbo.py:1449-1528): copies
 * z / mask / mu_table / packed targets host->device, imputes, decodes, scores, copies counts (and the mean grid if
 * requested) back and synchronises.  Host buffers may be pageable or pinned. */
int a3d_anytime_eval_host(a3d_handle* h, const float* z, const float* mask, const float* mu_table, int C, int64_t B,
                          int K, uint64_t seed, uint64_t obj_offset, int fill_mode, const uint8_t* target_bits,
                          float thr, int64_t* counts, float* mean_prob_or_null);

/* Bytes of device memory the arena holds / would need for `n` resident decodes. */
size_t a3d_workspace_bytes(const a3d_handle* h, int64_t n);

/* Diagnostics for per-layer parity tests: copy hidden layer `layer` (0 = dense output [n,4,4,4,8], 1..4 = ConvT
 * outputs, NDHWC) of the most recent chunk to `host` as fp32.  Synchronous. */
int a3d_debug_read_layer(a3d_handle* h, int layer, int64_t n, float* host, size_t nbytes);

/* Diagnostics for roofline studies: run ONLY the fused tail `reps` times on the 32^3 x 64 activations the most recent
 * chunk left in the arena (B objects x K samples, B * K <= that chunk's decodes), timed with CUDA events on `stream`;
 * *ms_per_launch receives the average.  counts_dev [B,3] is zeroed first and accumulates over the repetitions.
 * Synchronous.  (Inside a step the tail shares the GPU's power budget with the tensor-bound layers; alone it shows what
 * the kernel and the memory system do at the clock a memory-bound kernel is given.) */
int a3d_debug_time_tail(a3d_handle* h, int64_t B, int K, const uint8_t* target_bits_dev, int64_t* counts_dev, int reps,
                        float* ms_per_launch, void* stream);

/* Kernel launches issued by this handle since creation (what bench.py reports as gpu_launches). */
int64_t a3d_launch_count(const a3d_handle* h);

/* Device-side time of the last a3d_decode / a3d_anytime_eval per stage, in ms, when profiling was enabled with
 * a3d_set_profiling(h, 1): stages = dense+L1, L2, L3, L4, tail.  Returns number of stages written. */
int a3d_set_profiling(a3d_handle* h, int enable);
int a3d_stage_times_ms(a3d_handle* h, float* out, int max_stages);

/* ------------------------------------------------------------------------------------------------------------------
 * Image encoder of the Pascal3D path: Darknet19 backbone + head2D,
 * src/net_core/darknet.py:83-133 (Darknet19Conv / Darknet19), :135-168 (convHead / head2D).  A handle holds one
 * Keras-style functional model given as a flat layer list; the reference builds two models (backbone, head) and
 * calls head(backbone(images)) (src/module/nolbo.py:869), so two handles chain through a 16-bit NHWC feature buffer,
 * or one handle holds both lists concatenated.
 * ------------------------------------------------------------------------------------------------------------------ */
#define A3D_ENC_MAX_LAYERS 40
enum { A3D_L2D_CONV = 0,         /* Conv2D(filters, ksize 1|3, strides 1, 'same', use_bias=False) [+ BN] [+ act] */
       A3D_L2D_MAXPOOL = 1,      /* MaxPool2D(2, 2, 'same') on even sizes                                         */
       A3D_L2D_GLOBAL_MAX = 2,   /* tf.reduce_max(axis=[1,2])   darknet.py:159-160                                */
       A3D_L2D_GLOBAL_AVG = 3 }; /* tf.reduce_mean(axis=[1,2])  darknet.py:162-163                                */
enum { A3D_IO_F16 = 0, A3D_IO_BF16 = 1, A3D_IO_F32 = 2 };   /* element type of a forward() input / output buffer */

typedef struct {
  int32_t kind;         /* A3D_L2D_*                                  */
  int32_t filters;      /* conv only                                  */
  int32_t ksize;        /* conv only: 1 or 3                          */
  int32_t batch_norm;   /* conv only: BatchNormalization() follows    */
  int32_t activation;   /* conv only: A3D_ACT_*                       */
} a3d_layer2d;

typedef struct {
  int32_t abi_version;  /* A3D_ABI_VERSION */
  int32_t in_h, in_w;   /* input height / width, fixed at create; every MaxPool2D must see even sizes (multiples of 32
                           for Darknet19; the reference runs 256 x 256, test_pascal_VAE_dr.py:52) */
  int32_t in_ch;        /* 3 for images; a multiple of 64 for feature maps (head-only model) */
  int32_t num_layers;
  a3d_layer2d layers[A3D_ENC_MAX_LAYERS];
  int32_t device;
  int32_t max_batch;    /* images resident in the arena at once; larger calls are processed in chunks */
  int32_t operand_dtype;/* A3D_DTYPE_* of the tensor-core layers */
} a3d_enc2d_desc;

typedef struct a3d_enc2d a3d_enc2d;

/* Darknet19(name, activation) / head2D(name, input_shape, output_dim, ...) -> model     darknet.py:96-133,149-168 */
int a3d_enc2d_create(const a3d_enc2d_desc* desc, a3d_enc2d** out);
void a3d_enc2d_destroy(a3d_enc2d* h);

/* Keras variables in model.get_weights() order: per conv layer kernel [kh,kw,Cin,Cout], then (if BN) gamma, beta,
 * moving_mean, moving_variance.  fp32, host, synchronous; folding / repacking happens lazily before the next forward. */
int a3d_enc2d_num_weights(const a3d_enc2d* h);
int64_t a3d_enc2d_weight_numel(const a3d_enc2d* h, int index);
int a3d_enc2d_set_weight(a3d_enc2d* h, int index, const float* host, size_t nbytes);
int a3d_enc2d_get_weight(const a3d_enc2d* h, int index, float* host, size_t nbytes);

/* Output shape of the model for one input: out_dims[0..2] = (h, w, c); h = w = 1 after a global pool. */
int a3d_enc2d_output_shape(const a3d_enc2d* h, int32_t* out_dims);

/* model(images, training=False)            src/module/nolbo.py:869 (backbone, head).
 * in_dev:  [n, in_h, in_w, in_ch] NHWC on the device, in_dtype = A3D_IO_F32 (images) or the handle's operand dtype.
 * out_dev: [n, h, w, c] NHWC ([n, c] after a global pool), out_dtype = A3D_IO_F32 or the operand dtype
 * (a global pool always writes fp32).  Asynchronous on `stream`. */
int a3d_enc2d_forward(a3d_enc2d* h, const void* in_dev, int in_dtype, int64_t n, void* out_dev, int out_dtype,
                      void* stream);

/* Same model call with HOST images (what the reference's getEval holds: numpy batches from the loader,
 * src/module/nolbo.py:855-869): fp32 NHWC images in host memory (pinned memory gives the full PCIe rate), processed in
 * chunks of max_batch with the H2D copy of chunk i+1 overlapping the forward of chunk i on internal streams.  The
 * result ([n, h*w*c] fp32) goes to out_dev (device pointer) and / or out_host; synchronous. */
int a3d_enc2d_forward_host(a3d_enc2d* h, const float* images_host, int64_t n, float* out_dev_or_null, float* out_host_or_null);

/* Same call with the loader's raw bytes: uint8 NHWC images, multiplied by `scale` on the device (the reference's loader
 * does `image = image / 255.` on the host before the model sees it, src/dataset_loader/pascal3D.py:242; pass 1/255).
 * The H2D copy carries one byte per sample instead of four: the host path of a 256 x 256 batch is PCIe bound. */
int a3d_enc2d_forward_host_u8(a3d_enc2d* h, const uint8_t* images_host, float scale, int64_t n, float* out_dev_or_null,
                              float* out_host_or_null);

/* Latent split of the callers + sampling()  src/module/nolbo.py:869-875; src/module/function.py:35-38:
 * mean = enc_out[:, :D]; logvar = clip(enc_out[:, D:2D], -clip, clip); z = mean + sqrt(exp(logvar)) * eps,
 * eps ~ N(0,1) from Philox4x32-10 with counter (dim/4, 0x5A4D504C, obj_offset + b), key = seed (eps = 0 if
 * seed_enable == 0).  enc_out_dev: [n, out_stride] fp32; mean/logvar/z: [n, D] fp32 (any may be NULL). */
int a3d_enc2d_split_sample(a3d_enc2d* h, const float* enc_out_dev, int64_t n, int D, int out_stride, float clip,
                           int seed_enable, uint64_t seed, uint64_t obj_offset, float* mean_dev, float* logvar_dev,
                           float* z_dev, void* stream);

/* Diagnostics: copy the output of layer `layer` (index into the layer list; pooled if a fused pool follows is NOT
 * applied -- see a3d_enc2d_layer_shape) of the most recent chunk to `host` as fp32 NHWC with the padded channel
 * count.  Synchronous. */
int a3d_enc2d_layer_shape(const a3d_enc2d* h, int layer, int32_t* dims /* h, w, c_real, c_padded */);
int a3d_enc2d_debug_read_layer(a3d_enc2d* h, int layer, int64_t n, float* host, size_t nbytes);

int64_t a3d_enc2d_launch_count(const a3d_enc2d* h);
size_t a3d_enc2d_workspace_bytes(const a3d_enc2d* h);

/* ------------------------------------------------------------------------------------------------------------------
 * Voxel encoder: encoder3D(structure), src/net_core/autoencoder3D.py:72-102 -- Conv3D(k4,
 * strides 2, 'same', no bias) + BN + activation x (L-1), a bare Conv3D(k4, strides 1, 'same'), reduce_mean / reduce_max
 * over the grid, optional sigmoid.  Called as self._encoder(voxels, training=False) at src/module/nolbo.py:1463.
 * ------------------------------------------------------------------------------------------------------------------ */
enum { A3D_POOL_NONE = 0, A3D_POOL_AVERAGE = 1, A3D_POOL_MAX = 2 };

/* Mirrors the `structure` dict of encoder3D(structure), autoencoder3D.py:73-80. */
typedef struct {
  int32_t abi_version;               /* A3D_ABI_VERSION */
  int32_t in_grid;                   /* structure['input_shape'][0] (cubic, one channel): 64 */
  int32_t num_layers;                /* len(filter_num_list) */
  int32_t filters[A3D_MAX_LAYERS];   /* structure['filter_num_list'] */
  int32_t ksizes[A3D_MAX_LAYERS];    /* structure['filter_size_list'] */
  int32_t strides[A3D_MAX_LAYERS];   /* structure['strides_list'] */
  int32_t final_pool;                /* A3D_POOL_*   (structure['final_pool']) */
  int32_t activation;                /* A3D_ACT_*    (structure['activation']; 'lrelu' = LeakyReLU() = slope 0.3) */
  int32_t final_activation;          /* A3D_FINAL_*  (structure['final_activation']) */
  int32_t device;
  int32_t max_batch;                 /* objects resident in the arena at once */
  int32_t operand_dtype;             /* A3D_DTYPE_* */
} a3d_enc3d_desc;

typedef struct a3d_enc3d a3d_enc3d;

int a3d_enc3d_create(const a3d_enc3d_desc* desc, a3d_enc3d** out);
void a3d_enc3d_destroy(a3d_enc3d* h);
/* Keras get_weights() order: per conv3DEnc kernel [kd,kh,kw,Cin,Cout], gamma, beta, moving_mean, moving_variance; last: kernel. */
int a3d_enc3d_num_weights(const a3d_enc3d* h);
int64_t a3d_enc3d_weight_numel(const a3d_enc3d* h, int index);
int a3d_enc3d_set_weight(a3d_enc3d* h, int index, const float* host, size_t nbytes);
int a3d_enc3d_get_weight(const a3d_enc3d* h, int index, float* host, size_t nbytes);
/* encoder(voxels, training=False): voxels_dev [n, G, G, G, 1] fp32 on the device -> out_dev fp32 [n, filters[-1]]
 * ([n, g, g, g, filters[-1]] when final_pool is A3D_POOL_NONE).  Asynchronous on `stream`. */
int a3d_enc3d_forward(a3d_enc3d* h, const float* voxels_dev, int64_t n, float* out_dev, void* stream);
/* same contract as a3d_enc2d_split_sample (src/module/nolbo.py:1464-1470) */
int a3d_enc3d_split_sample(a3d_enc3d* h, const float* enc_out_dev, int64_t n, int D, int out_stride, float clip,
                           int seed_enable, uint64_t seed, uint64_t obj_offset, float* mean_dev, float* logvar_dev,
                           float* z_dev, void* stream);
/* Diagnostics: hidden layer `layer` (0 .. num_layers-2, NDHWC, after BN + activation) of the most recent chunk as fp32. */
int a3d_enc3d_debug_read_layer(a3d_enc3d* h, int layer, int64_t n, float* host, size_t nbytes);
int64_t a3d_enc3d_launch_count(const a3d_enc3d* h);
size_t a3d_enc3d_workspace_bytes(const a3d_enc3d* h);

/* Host utility for the TensorFlow-checkpoint reader (anytime-3d-reconstruction_b200/tf_checkpoint.py; reference weights
 * are Keras save_weights bundles, src/module/nolbo.py:1568-1592): CRC-32C (Castagnoli) of `n` bytes, continuing from
 * `crc` (0 to start).  Pure host code, no device needed. */
uint32_t a3d_crc32c(const void* data, size_t n, uint32_t crc);

const char* a3d_last_error(void);
int a3d_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* A3D_H_ */
