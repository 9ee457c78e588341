/*
 * crimac_b200.h — C ABI of libcrimac_b200.so: the B200 (sm_100a) implementation of the CRIMAC echogram U-Net hot path.
 *
 * The reference (CRIMAC-classifiers-unet) is pure Python/PyTorch and has NO native boundary of its own; the seam this
 * library sits behind is the nn.Module stored in SegPipe.model.  Each entry point below names the reference call it
 * replaces (paths relative to the reference's crimac_unet/ directory):
 *
 *   crimac_forward_infer   models/unet.py:327-343 (UNet_Baseline.forward, eval) + pipeline_train_predict/pipeline.py:218 (F.softmax)
 *   crimac_forward_train   models/unet.py:327-343 under model.train()  (pipeline.py:167-171)
 *   crimac_loss            pipeline.py:135-138,176 (nn.CrossEntropyLoss(weight=[10,300,250]))
 *   crimac_backward        pipeline.py:177 (loss.backward(): autograd of every layer of models/unet.py)
 *   crimac_train_step      pipeline.py:171-177 in one call
 *   crimac_sgd_step        pipeline.py:156,178 (optim.SGD(momentum) step)
 *   crimac_preprocess      batch/dataset.py:192-205 + utils/np.py:362-375 + batch/data_transforms/{remove_nan_inf,db_with_limits}.py
 *   crimac_train_patches   batch/dataset.py:75-108,358-407 (Dataset.__getitem__ / get_crop_zarr) + batch/data_augmentation/*.py +
 *                          batch/label_transforms/{refine_label_boundary,convert_label_indexing}.py + data transforms
 *   crimac_meta_channels   batch/dataset.py:296-349 (metadata input channels of get_crop_memmap; pipeline.py:413-425)
 *   crimac_stitch          pipeline_train_predict/save_predict.py:41-65 (fill_out_array) + label masks of
 *                          batch/label_transforms/mask_label_{overlap,seabed}.py
 *   crimac_eval_loss       pipeline.py:222-239 (set_label_ignore_val) + :264 (validation loss) + :269-270 (softmax,
 *                          SANDEEL channel) of get_predictions_dataloader, on the eval-mode logits
 *   crimac_forward_infer_fp32   the same forward in plain fp32 (validation mode, 1e-4 parity)
 *   crimac_op_* / crimac_dbg_*   single-kernel entry points used by the parity tests only.
 *
 * Conventions: every function returns 0 on success, 1 for an invalid argument, 2 for a CUDA failure;
 * crimac_last_error() returns a thread-local description.  All pointers named *_dev are device pointers owned by the
 * caller (PyTorch); `stream` is a cudaStream_t.  The library never allocates device memory, never synchronises and
 * never falls back to the CPU.  One context per (process, device); calls on one context are not re-entrant.
 *
 * Parameter tables.  `state` is an array of 136 device pointers in the order of UNet_Baseline.state_dict()
 * (SURVEY.md App. B): for each encoder block i: main.0.{weight,bias}, main.1.{weight,bias,running_mean,running_var,
 * num_batches_tracked}, main.3.{weight,bias}, main.4.{...5}; for each decoder block j: upconv.{weight,bias},
 * conv1.{weight,bias}, conv2.{weight,bias}, bn1.{weight,bias,running_mean,running_var,num_batches_tracked}, bn2.{...5};
 * then conv_final.{weight,bias}.  fp32 except num_batches_tracked (int64).  `grads` is an array of 82 fp32 device
 * pointers in the order of UNet_Baseline.parameters() (the same list without the BN buffers).
 */
#ifndef CRIMAC_B200_H
#define CRIMAC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct crimac_ctx crimac_ctx;

typedef struct crimac_config {
  int in_channels;   /* frequencies + metadata channels, 1..12             (unet.py:200 in_channels) */
  int n_classes;     /* 1..8                                                (unet.py:200 n_classes)   */
  int depth;         /* encoder blocks, 2..5 (reference default 5)          (unet.py:206)             */
  int start_filts;   /* must be 64                                          (unet.py:207)             */
  int max_batch;     /* largest batch a call may pass                                                 */
  int height, width; /* patch size, multiples of 2^(depth-1) (the reference's own constraint)          */
  int train;         /* 1: allocate saved activations + gradient scratch for forward_train/backward   */
  int deterministic; /* 1 (train): split-K weight gradients are summed in a FIXED order (per-split slabs,     */
                     /* ~0.9 GB more workspace) instead of red.global.add: bit-reproducible steps, the        */
                     /* equivalent of the reference's torch.backends.cudnn.deterministic (utils/general.py)   */
  int up_mode;       /* 0: "transpose" (ConvTranspose2d k2 s2), 1: "upsample" (bilinear 2x + conv1x1)         */
                     /*    (unet.py:47-56; the state table then holds upconv.1.{weight (Cout,Cin,1,1),bias})   */
  int merge_mode;    /* 0: "concat" (torch.cat((up, skip), 1)), 1: "add" (up + skip; conv1 takes C channels)   */
                     /*    (unet.py:113-118,131-134); up_mode 1 with merge_mode 1 is rejected as in the ctor   */
} crimac_config;

const char* crimac_last_error(void);
int crimac_state_count(const crimac_config* cfg); /* 136 for depth 5 */
int crimac_grad_count(const crimac_config* cfg);  /* 82 for depth 5  */

int crimac_workspace_bytes(const crimac_config* cfg, size_t* bytes);
int crimac_create(crimac_ctx** ctx, const crimac_config* cfg, void* workspace_dev, size_t workspace_bytes, int device);
int crimac_destroy(crimac_ctx* ctx);

/* Re-pack fp32 parameters into the bf16 GEMM operands (and, for train==0, fold BatchNorm running statistics into the
 * conv epilogue).  Must be called after every change of the parameters and before the forward call that uses them. */
int crimac_prepare(crimac_ctx* ctx, const void* const* state, int train, void* stream);

/* x_dev: fp32 NCHW (nb, in_channels, H, W), or NULL when crimac_preprocess_staged has staged nb patches in ctx.
 * out_dev: fp32 NCHW (nb, n_classes, H, W): class probabilities (softmax != 0) or raw logits. */
int crimac_forward_infer(crimac_ctx* ctx, const void* const* state, const float* x_dev, int nb, float* out_dev,
                         int softmax, void* stream);
/* The same eval forward with softmax AND the overlap stitching of crimac_stitch fused into the last conv's epilogue:
 * classes cls[0..K) of every kept pixel go straight into out_dev (K, R, Pc) fp16; no probability tensor is written.
 * Arguments as crimac_stitch; x_dev may be NULL after crimac_preprocess_staged. */
int crimac_forward_infer_stitch(crimac_ctx* ctx, const void* const* state, const float* x_dev, int nb,
                                const int32_t* centres_dev, const uint8_t* nan_dev, const int16_t* labels_dev,
                                const int32_t* seabed_dev, int seabed_pad, int overlap, int ping_start, int Pc, int R,
                                const int32_t* cls, int K, void* out_dev, void* stream);
/* Train-mode forward (batch statistics, running-stat update, activations kept for backward). logits_dev as above. */
int crimac_forward_train(crimac_ctx* ctx, const void* const* state, const float* x_dev, int nb, float* logits_dev,
                         void* stream);
/* Class-weighted cross-entropy with ignore_index.  labels_dev: int64 (nb,H,W).  out3_dev: {loss, 1/sum_w, sum_w}.
 * dlogits_dev (optional): UNNORMALISED gradient w[y]*(softmax-onehot); multiply by out3[1] (crimac_backward does). */
int crimac_loss(crimac_ctx* ctx, const float* logits_dev, const int64_t* labels_dev, const float* class_w_dev,
                int64_t ignore_index, int nb, float* out3_dev, float* dlogits_dev, void* stream);
/* Backward of the last crimac_forward_train.  dlogits_dev: fp32 NCHW gradient of the loss w.r.t. the logits;
 * gscale_dev (optional) points at one device float multiplied into it.  Writes (not accumulates) all 82 gradients. */
int crimac_backward(crimac_ctx* ctx, const void* const* state, const float* x_dev, const float* dlogits_dev,
                    const float* gscale_dev, int nb, float* const* grads, void* stream);
/* prepare(train) + forward_train + loss + backward.  loss3_dev as crimac_loss. */
int crimac_train_step(crimac_ctx* ctx, const void* const* state, const float* x_dev, const int64_t* labels_dev,
                      const float* class_w_dev, int64_t ignore_index, int nb, float* const* grads, float* loss3_dev,
                      void* stream);

/* SGD with momentum on flat fp32 arrays: v = momentum*v + g*gscale ; p -= lr*v   (torch.optim.SGD, dampening 0). */
int crimac_sgd_step(float* params_dev, float* momentum_dev, const float* grads_dev, size_t n, float lr, float momentum,
                    float gscale, void* stream);

/* Validation step (pipeline.py:249-270): remaps the raw label codes as set_label_ignore_val does (-70, -30, -100, -10 ->
 * ignore; -50 -> background), evaluates nn.CrossEntropyLoss(weight) on the remapped labels and writes the softmax
 * probability of class `prob_class` (SANDEEL = 1 in the reference).  logits_dev: fp32 NCHW (nb, n_classes, H, W);
 * labels_dev: int16 (label_bits = 16, what the dataset emits) or int64 (label_bits = 64), (nb, H, W); prob_out_dev
 * (optional) fp32 (nb, H, W); labels_out_dev (optional) int64 (nb, H, W) remapped labels; out3_dev {loss, 1/sum_w,
 * sum_w}; scratch_dev: >= 16 KB.  A label outside [0, n_classes) after the remap makes the loss NaN. */
int crimac_eval_loss(const float* logits_dev, int nb, int n_classes, int H, int W, const void* labels_dev,
                     int label_bits, const float* class_w_dev, int prob_class, float* prob_out_dev,
                     int64_t* labels_out_dev, float* out3_dev, void* scratch_dev, void* stream);

/* ---- data-parallel gradient exchange over NVLink peer memory (the reference has no multi-GPU path)
 * Sum-all-reduce of floats [offset, offset+count) of a flat fp32 gradient arena that every replica allocated
 * SYMMETRICALLY (same size, mapped on every GPU of the node, e.g. torch.distributed._symmetric_memory), in ONE kernel on
 * `stream`: flag exchange, each rank reduces its 1/world slice in fixed rank order and stores the sums into all arenas
 * (multimem.ld_reduce / multimem.st through the NVSwitch when multicast_arena_dev != NULL).  Results are bit-identical on
 * all replicas.  peer_arenas / peer_pads: HOST arrays of `world` device pointers (entry `rank` = the local one); pads:
 * symmetric, zero-initialised, >= CRIMAC_AR_PAD_BYTES each; local_state_dev: zero-initialised, >= 8 *
 * CRIMAC_AR_MAX_BUCKETS bytes, private to this rank.  `bucket` (0..CRIMAC_AR_MAX_BUCKETS-1) selects the flag rows, so
 * several buckets may be in flight.  offset, count: multiples of 4.  ctas <= 0: default.  The kernel keeps its own epoch
 * counter in local_state_dev: every rank must launch the same sequence of calls (CUDA-graph replay safe). */
#define CRIMAC_AR_MAX_BUCKETS 8
#define CRIMAC_AR_PAD_BYTES (CRIMAC_AR_MAX_BUCKETS * 2 * 16 * 4)
int crimac_peer_allreduce(float* const* peer_arenas, void* const* peer_pads, float* multicast_arena_dev,
                          void* local_state_dev, int rank, int world, int bucket, size_t offset, size_t count, int ctas,
                          void* stream);

typedef struct crimac_comm_config {
  int world, rank;                /* replicas on this node (<= 8), this replica's index                          */
  float* peer_arenas[8];          /* every replica's symmetric gradient arena as mapped on THIS device            */
  void* peer_pads[8];             /* every replica's symmetric signal pad (CRIMAC_AR_PAD_BYTES, zeroed)           */
  float* multicast_arena;         /* NVSwitch multicast address of the arena, or NULL (peer loads / stores)       */
  void* local_state;              /* private device scratch, 8 * CRIMAC_AR_MAX_BUCKETS bytes, zeroed              */
  size_t arena_floats;            /* arena size in floats, multiple of 4 (pad the 31 044 227 parameters up)       */
  int ctas;                       /* CTAs of the exchange kernel (<= 0: default 128)                              */
} crimac_comm_config;
/* Switch the bucketed, backward-overlapped gradient all-reduce of crimac_backward / crimac_train_step on (cfg != NULL)
 * or off (NULL); see csrc/net_api.cu.  The gradient pointers of those calls must then be views of peer_arenas[rank]. */
int crimac_set_comm(crimac_ctx* ctx, const crimac_comm_config* cfg);

typedef struct crimac_optimizer_config {
  float* params;        /* flat fp32 parameters, parameters() order (the module's tensors are views of it)          */
  float* momentum;      /* flat fp32 momentum buffer                                                               */
  const float* grads;   /* flat fp32 gradient arena the `grads` table points into                                  */
  size_t n;             /* number of parameters                                                                    */
  float lr, momentum_coef, gscale;   /* v = momentum_coef*v + g*gscale ; p -= lr*v  (gscale = 1/world)            */
} crimac_optimizer_config;
/* Fuse optim.SGD(momentum).step() (pipeline.py:156,178) into crimac_backward / crimac_train_step, bucket by bucket,
 * overlapped with the rest of backward; NULL switches it off.  See csrc/net_api.cu. */
int crimac_set_optimizer(crimac_ctx* ctx, const crimac_optimizer_config* cfg);

/* Patch gather + sv->dB transform.  sv_dev: fp32 (F, R, P) preloaded pings [frequency][range][ping] whose column 0 is
 * survey ping data_ping0; centres_dev: int32 (n,2) patch centres (y, x) in survey coordinates (batch/samplers/
 * gridded.py:22-54); out_dev: fp32 NCHW (n, F, ph, pw) = clip(10*log10(sv+1e-10), -75, 0) with out-of-data and
 * non-finite samples -> 0 before the transform; nan_dev (optional) uint8 (n, ph, pw) = 1 where frequency 0 was
 * non-finite (the pixels remove_nan_inf marks LABEL_IGNORE_VAL). */
int crimac_preprocess(const float* sv_dev, int F, int R, int P, int data_ping0, const int32_t* centres_dev, int n,
                      int ph, int pw, float* out_dev, uint8_t* nan_dev, void* stream);
/* The same gather + transform written STRAIGHT into the first conv's operand (bf16 hi/lo pairs, NHWC) inside ctx: the
 * next crimac_forward_infer(ctx, state, x_dev = NULL, nb = n, ...) reads it by TMA and the fp32 NCHW patch tensor is
 * never materialised.  Patch size = the context's (height, width); F = its in_channels; n <= max_batch. */
int crimac_preprocess_staged(crimac_ctx* ctx, const float* sv_dev, int F, int R, int P, int data_ping0,
                             const int32_t* centres_dev, int n, uint8_t* nan_dev, void* stream);
/* Metadata input channels (batch/dataset.py:296-349), written as fp32 planes [c_off, c_off + M) of the network input
 * x_out (n, c_total, ph, pw).  mask bits: 1 portion_year (1 channel), 2 portion_day (2: sin, cos), 4 time_diff, 8
 * depth_rel, 16 depth_abs_surface, 32 depth_abs_seabed - appended in that order, as the reference does.  centres_dev:
 * int32 (n,2) (y, x); the three vectors are per-ping doubles of the echogram (portion_of_day_vector, time_vector_diff,
 * _seabed); out-of-range pings read element 0 resp. the last one, rows are not clamped (the reference's conventions). */
int crimac_meta_channels(const int32_t* centres_dev, int n, int ph, int pw, unsigned mask, double portion_year,
                         const double* portion_of_day_dev, int n_pod, const double* time_diff_dev, int n_tvd,
                         const double* seabed_dev, int n_sb, float* x_out_dev, int c_total, int c_off, void* stream);
/* Overlap-stitch (fill_out_array): for every patch pixel whose label would not be one of {-70 overlap frame,
 * -50 below seabed+pad on background, -100 outside [ping_start, ping_start+Pc) x [0,R) or non-finite}, write
 * probs[:, cls[k]] as fp16 into out_dev (K, R, Pc).  labels_dev: optional int16 (R, Pc) chunk labels after the
 * label-only transforms (NULL = all background); seabed_dev: optional int32 (Pc) first below-seabed range index per
 * ping (NULL = no seabed mask); cls: HOST array of K (<=4) class indices. Pixels never written keep their value. */
int crimac_stitch(const float* probs_dev, int n, int n_classes, int ph, int pw, const int32_t* centres_dev,
                  const uint8_t* nan_dev, const int16_t* labels_dev, const int32_t* seabed_dev, int seabed_pad,
                  int overlap, int ping_start, int Pc, int R, const int32_t* cls, int K, void* out_dev, void* stream);

/* Training-sample path: one batch of random crops with augmentation, label refinement
 * and the dB transform, from a survey resident in HBM in the zarr store's own order.
 *   sv_dev      fp32 (F, P, R) [frequency][ping][range];  labels_dev fp32 (P, R) raw annotation category per sample
 *               (0 background, 27 sandeel, 1 other, other species > 0, NaN / -100 = no data)
 *   centres_dev int32 (n,2) crop centres (range, ping): sample (py,px) of crop b is survey sample
 *               (cy - ph/2 + 1 + py, cx - pw/2 + 1 + px) (utils/np.py:378-380); outside the survey -> sv 0, label -100
 *   flags_dev   uint8 (n): bit0 = add_noise fires for this crop, bit1 = flip_x_axis fires (the two coin flips of
 *               add_noise.py:25 / flip_x_axis.py:22 are the caller's)
 *   noise_mult_dev  optional fp32 (n,F,ph,pw): explicit multipliers in un-flipped crop coordinates (replay / tests);
 *               NULL = draw add_noise.py:28-38's distribution from Philox4x32-10 keyed by (noise_seed, sample index)
 *   thr_freq, thr_lo, thr_hi  refine_label_boundary's threshold channel index and (1e-7, 1e-4) window on linear sv
 *   scaled      0: db_with_limits, 1: db_with_limits_scaled;  border_zero: set_data_border_value on the final labels
 *   x_out       fp32 (n,F,ph,pw) network input;  labels_out int64 (n,ph,pw) in {0,1,2,-100}
 * ph == pw, multiples of 32.  Two launches on `stream`, no allocation, no synchronisation. */
int crimac_train_patches(const float* sv_dev, const float* labels_dev, int F, int P, int R, const int32_t* centres_dev,
                         const uint8_t* flags_dev, const float* noise_mult_dev, uint64_t noise_seed, int n, int ph,
                         int pw, int thr_freq, double thr_lo, double thr_hi, int scaled, int border_zero,
                         float* x_out, int64_t* labels_out, void* stream);

/* ---- fp32 VALIDATION mode of the eval forward (models/unet.py:327-343 + pipeline.py:218): an independent plain-fp32
 * CUDA-core implementation reading the fp32 parameters directly (no packing, no tensor cores, no bf16).  ~50x slower
 * than crimac_forward_infer; for parity checks at 1e-4, not for production.  Stateless: the caller passes a scratch
 * buffer of crimac_fp32_workspace_bytes(cfg, nb).  cfg->max_batch and cfg->train are ignored. */
int crimac_fp32_workspace_bytes(const crimac_config* cfg, int nb, size_t* bytes);
int crimac_forward_infer_fp32(const crimac_config* cfg, const void* const* state, const float* x_dev, int nb,
                              float* out_dev, int softmax, void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---- measurement aids (bench.py): launch counter and per-launch CUDA-event timing of the network-level calls */
unsigned long long crimac_launch_count(void);      /* kernels launched by this process through the library so far */
int crimac_profile_enable(int on);                 /* clears earlier records; on!=0 brackets every launch with events */
int crimac_profile_read(const char** names, float* ms, double* flops, double* bytes, int* launches, int cap);

/* ---- single-kernel entry points (parity tests) */
int crimac_op_igemm(int mode, const void* x, int NB, int H, int W, int cin, int x_pitch, const void* w, int n_total,
                    const float* scale, const float* shift, int relu, void* out, int out_pitch, int convt_cout,
                    void* pool_out, int pool_pitch, float* stats, const float* head_w, const float* head_b,
                    float* head_out, int n_classes, int head_softmax, int block_n, void* stream);
int crimac_op_wgrad(int mode, const void* f, int f_pitch, int m_total, const void* t, int t_pitch, int n_total, int NB,
                    int H, int W, float* scratch, float* dw, int splits, int block_n, void* stream);
int crimac_op_wgrad_halo(const void* dy, int dy_pitch, int cout, const void* x, int x_pitch, int cin, int NB, int H, int W,
                         float* scratch, float* dw, int splits, void* stream);
/* HBM-bound kernels one at a time (scratch: crimac_op_scratch_bytes() bytes of device memory); see csrc/ops_api.cu */
size_t crimac_op_scratch_bytes(void);
int crimac_op_bn_finalize(const float* partials, int rows, int C, double count, const float* gamma, const float* beta,
                          float* running_mean, float* running_var, int64_t* num_batches_tracked, float momentum,
                          float eps, float* scale, float* shift, float* save_mean, float* save_invstd, void* stream);
int crimac_op_bn_apply(const void* raw, int raw_pitch, int N, int H, int W, int C, const float* scale,
                       const float* shift, void* act, int act_pitch, void* pool, int pool_pitch, uint16_t* pool_arg,
                       void* stream);
int crimac_op_bn_bwd(const void* dact, int dact_pitch, const void* raw, int raw_pitch, int N, int H, int W, int C,
                     const float* scale, const float* shift, const float* mean, const float* invstd, void* draw,
                     int draw_pitch, float* dgamma, float* dbeta, float* dbias, const float* gscale, void* scratch,
                     void* stream);
int crimac_op_bn_bwd_pool(const uint16_t* pool_arg, const void* dpool, int dpool_pitch, const void* dskip,
                          int dskip_pitch, const void* raw, int raw_pitch, int N, int H, int W, int C, const float* scale,
                          const float* shift, const float* mean, const float* invstd, void* draw, int draw_pitch,
                          float* dgamma, float* dbeta, float* dbias, void* scratch, void* stream);
int crimac_op_pool_bwd_add(const uint16_t* pool_arg, const void* dpool, int dpool_pitch, const void* dskip,
                           int dskip_pitch, void* dact, int dact_pitch, int N, int H, int W, int C, void* stream);
int crimac_op_head_ce(const void* act, int act_pitch, int N, int H, int W, const float* hw, const float* hb, int ncls,
                      const int64_t* labels, const float* cw, int64_t ignore_index, void* dact, int dact_pitch,
                      float* dw, float* db, float* out3, void* scratch, void* stream);
int crimac_op_head_fwd(const void* act, int act_pitch, int N, int H, int W, const float* hw, const float* hb, int ncls,
                       float* logits, void* stream);
int crimac_op_head_bwd(const float* dlogits, const float* gscale, const void* act, int act_pitch, int N, int H, int W,
                       const float* hw, int ncls, void* dact, int dact_pitch, float* dw, float* db, void* scratch,
                       void* stream);
int crimac_op_colsum(const void* v, int pitch, int N, int H, int W, int C, float* out, void* scratch, void* stream);
int crimac_op_upsample2x(void* lo, int lo_pitch, void* hi, int hi_pitch, int N, int H, int W, int C, int backward,
                         void* stream);
int crimac_op_pack(int kind, const float* w, int cout, int cin, void* out, void* stream);
int crimac_op_wgrad_unpack_all(int n, float* const* scratch, float* const* dw, const int64_t* mn, const int* taps,
                               void* stream);
/* Test hook: copy a tensor saved by the last train-mode forward (raw conv output / activation / pooled / ConvTranspose
 * output) into a dense NHWC bf16 buffer; see csrc/net_api.cu. */
int crimac_dbg_saved(crimac_ctx* ctx, int index, int which, int nb, void* dst_dev, int* dims_out, void* stream);
int crimac_dbg_umma(const void* image_dev, int image_bytes, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                    int n_mma, int a_step_bytes, int b_step_bytes, float* out_dev, int N, void* stream);
int crimac_dbg_tma_box(const void* x_dev, int NB, int H, int W, int C, int pitch, int box_h, int sub, int ky, int kx,
                       int c0, int x0, int y0, int n0, void* out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CRIMAC_B200_H */
