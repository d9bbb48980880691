/* xptwarp.h -- C-ABI of libxptwarp.so: B200-native (sm_100a) view synthesis +
 * photometric SSIM/L1 + edge-aware smoothness loss, forward and backward.
 *
 * The reference (goodgodgd/xpt-mde-2021) has NO FFI/plugin interface: the
 * boundary of this path is a Python call surface.  Each
 * entry point below therefore names the reference Python interface it stands
 * behind (file:line relative to the reference checkout); the ctypes binding that
 * a maintainer would add is in INTEGRATION.md and shipped in
 * xpt-mde-2021_b200/xptwarp/_cabi.py.
 *
 * Conventions
 *  - plain C: pointers, sizes, POD structs; no torch / TF / DLPack types.
 *    (The Python host unwraps DLPack capsules to these pointers, zero-copy.)
 *  - all tensor pointers are DEVICE pointers to fp32, channel-last, dense in the
 *    inner dims; only the snippet tensor may carry batch / frame strides so that
 *    source = image5d[:, :-1] and target = image5d[:, -1] are zero-copy views
 *    (reference model/loss_and_metric/losses.py:77-83).
 *  - the caller owns every input and output buffer; the library borrows them for
 *    the duration of the call and owns only the scratch inside xpt_ctx.
 *  - every call is asynchronous on the cudaStream_t passed as `stream`
 *    (a void* so that this header needs no CUDA include); 0 = default stream.
 *  - return value: XPT_OK (0) or a negative xpt_status; xpt_last_error() gives
 *    the thread-local message.  There is no CPU fallback.
 *  - one xpt_ctx per (device, host thread); calls on one ctx must be serialised
 *    by the caller (stream order is enough when one stream is used).
 *  - the fused tile kernel reads the camera geometry through the constant bank; every ctx owns a private slot
 *    of it, so calls of different contexts on different streams of one device do not interfere.  (Only when
 *    the 60 KB bank is exhausted -- more than ~15 live contexts of config-2 size -- does a new ctx share a slot;
 *    such contexts must be ordered by their caller.  xpt_scratch_bytes' sibling xpt_geometry_slot_shared tells.)
 */
#ifndef XPTWARP_H_
#define XPTWARP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define XPT_API __attribute__((visibility("default")))
#else
#define XPT_API
#endif

#define XPT_VERSION 200        /* 0.2.0 */
#define XPT_MAX_SCALES 8

typedef enum {
  XPT_OK = 0,
  XPT_BAD_ARGUMENT = -1,   /* null pointer, bad enum, bad flag combination        */
  XPT_BAD_SHAPE = -2,      /* sizes not divisible by the scales, too small, ...   */
  XPT_CUDA_ERROR = -3,     /* a CUDA runtime call or kernel launch failed         */
  XPT_NO_DEVICE = -4,      /* no usable sm_100 device                            */
  XPT_OUT_OF_MEMORY = -5,  /* scratch allocation failed                          */
  XPT_BAD_DTYPE = -6,      /* a tensor that is not float32 (xpt_check_dlpack)     */
  XPT_BAD_DEVICE = -7,     /* a tensor that does not live on the ctx's CUDA device */
  XPT_NOT_CONTIGUOUS = -8, /* inner dimensions not dense (xpt_check_dlpack)       */
  XPT_NCCL_ERROR = -9      /* libnccl missing, or an NCCL call failed             */
} xpt_status;

/* photometric term selector: reference loss_util.py:6-25 / :29-48 / :52-96 */
typedef enum { XPT_PHOTO_L1 = 0, XPT_PHOTO_L2 = 1, XPT_PHOTO_SSIM = 2 } xpt_photo_method;

/* Static description of the problem: the reference reads all of these from
 * static tensor shapes (synthesize_base.py:61-64) and from config/opts
 * (config-example.py:22,67-71,76-89).  Bound once at xpt_create.            */
typedef struct {
  int32_t batch;            /* B: snippets on THIS rank                               */
  int32_t num_src;          /* N: source frames per snippet (SNIPPET_LEN-1)           */
  int32_t height, width;    /* full-resolution H, W                                   */
  int32_t num_scales;       /* S <= XPT_MAX_SCALES                                    */
  int32_t scales[XPT_MAX_SCALES];      /* integer down-scale of each level (1,2,4,8) */
  float scale_weights[XPT_MAX_SCALES]; /* losses.py:147-154                           */
  float w_l1, w_ssim, w_smooth;        /* loss weights; 0 drops the term (loss_factory.py:41-43) */
  float img_grad_factor;    /* opts.IMAGE_GRADIENT_FACTOR = 4 (losses.py:429)         */
  int32_t global_batch;     /* tf.nn.compute_average_loss divisor (losses.py:49); 0 -> batch */
  int32_t device;           /* CUDA device ordinal                                    */
  uint32_t flags;           /* XPT_FLAG_*                                             */
} xpt_config;

/* xpt_total_loss runs one kernel per reference stage with tensors through HBM
 * (warp -> synth -> loss -> dL/dsynth -> warp adjoint) instead of the fused
 * tile kernel.  Same results; kept for A/B parity tests and profiling.        */
#define XPT_FLAG_UNFUSED 1u
/* xpt_total_loss captures its launches into a CUDA graph the first time it sees a
 * given argument set (all pointers, strides, grad_scale, stream) and replays the
 * instantiated graph afterwards (up to 64 cached argument sets per ctx).      */
#define XPT_FLAG_GRAPH 2u
/* xpt_total_loss_host runs the batch as ONE chunk (no copy/compute pipelining).                */
#define XPT_FLAG_NO_PIPELINE 4u
/* xpt_total_loss runs its training steps (gradients wanted, no synthesis tensors, no dL/dsource, <= 4 sources) on
 * the streaming strip kernel (k_strip: warp-specialised roles marching down 64-column strips, no halo re-warp,
 * per-ctx geometry from global memory) instead of the tile kernel (k_fused).  Same results to summation order.
 * Opt-in: on B200 it is 13-24 % slower than the tile kernel; kept for A/B.                   */
#define XPT_FLAG_STRIP 8u
/* multi-rank data parallelism (reference distributer.py:93-110 ReplicaOutputIntegrator): xpt_total_loss ends with an
 * in-place ncclAllReduce(sum) of out->losses[4] over the communicator bound by xpt_comm_init / xpt_comm_attach,
 * enqueued on the call's own stream right behind the epilogue kernel -- and therefore captured in the same CUDA
 * graph as the step under XPT_FLAG_GRAPH.  After xpt_comm_init across processes of one node the fused path does not
 * even launch that collective: the epilogue kernel itself pushes the 4 scalars into every rank's inbox with peer
 * stores over NVLink and sums what it received (xpt_comm_status tells which).  xpt_total_loss_host does NOT reduce: its losses land in host memory
 * chunk by chunk and stay rank-local (a host caller sums the 4 floats with whatever host-side collective it has). */
#define XPT_FLAG_ALLREDUCE 16u
/* depth_ms[] holds the depth net's LOGITS.  xpt_total_loss applies the net's last op itself --
 * InverseSigmoidActivation, depth = safe_reciprocal_number(sigmoid(x) + 0.01) (model/build_model/model_factory.py:
 * 133-137) -- when it loads a depth tile, and d_depth_ms[] receives dL/dlogit.  With disp_ms == NULL the disparity
 * of the smoothness term is formed in the kernel as well (model_wrappers.py:47-48), so the boundary takes one
 * tensor per level from the net and returns one gradient per level.  Fused path only.                           */
#define XPT_FLAG_DEPTH_LOGIT 32u
/* xpt_photometric_min_loss runs on the round-1 tile kernel (k_photo_min: 32x16 tiles, one thread per statistics
 * position) instead of the strip kernel (k_min_strip: 64x13 tiles, 2-pixel strips, running minimum in registers,
 * in-tile reduction of the up-sampling adjoint).  Same results to summation order; kept for A/B parity tests.    */
#define XPT_FLAG_MIN_TILES 64u

/* The snippet frames + intrinsics (features of losses.py:26-37).              */
typedef struct {
  const float* source;           /* [B,N,H,W,3]                                       */
  int64_t source_batch_stride;   /* elements between snippets (5*H*W*3 for an image5d view) */
  int64_t source_frame_stride;   /* elements between frames   (H*W*3)                 */
  const float* target;           /* [B,H,W,3]; may be NULL for xpt_synthesize*        */
  int64_t target_batch_stride;   /* elements between snippets                         */
  const float* intrinsic;        /* [B,3,3] dense                                     */
} xpt_frames;

/* Outputs of xpt_total_loss.  Every pointer is optional (NULL = not wanted)
 * except `losses`.  Arrays are indexed by level 0..S-1.                       */
typedef struct {
  float* losses;                       /* [4]: total, L1, SSIM, smoothe -- by-type entries are the
                                          UNWEIGHTED per-type means (losses.py:49-52), already divided
                                          by global_batch; multi-rank callers all-reduce(sum) this.   */
  float* loss_batch;                   /* [3][B] per-snippet L1, SSIM, smoothe (scale-merged)  */
  float* synth_ms[XPT_MAX_SCALES];     /* [B,N,H_s,W_s,3]  augm_data["synth_target_ms"]         */
  float* mask_ms[XPT_MAX_SCALES];      /* [B,N,H_s,W_s,1]  validity mask (bilinear_interp.py:53-76) */
  float* target_ms[XPT_MAX_SCALES];    /* [B,H_s,W_s,3]    augm_data["target_ms"] (level 0 = copy) */
  float* d_depth_ms[XPT_MAX_SCALES];   /* [B,H_s,W_s,1]    dL/d depth_ms                        */
  float* d_disp_ms[XPT_MAX_SCALES];    /* [B,H_s,W_s,1]    dL/d disp_ms                         */
  float* d_pose;                       /* [B,N,6]          dL/d pose                            */
  float* d_source;                     /* [B,N,H,W,3] dense, dL/d source (through the pyramid)  */
  float grad_scale;                    /* upstream dL/d total_loss (1 for a plain backward)     */
} xpt_loss_outputs;

typedef struct xpt_ctx xpt_ctx;

XPT_API int xpt_version(void);
XPT_API const char* xpt_last_error(void);
XPT_API const char* xpt_status_string(int status);

/* Binds shapes/weights, allocates scratch on cfg->device.                     */
XPT_API int xpt_create(xpt_ctx** out, const xpt_config* cfg);
XPT_API void xpt_destroy(xpt_ctx* ctx);
XPT_API int xpt_get_config(const xpt_ctx* ctx, xpt_config* out);
/* bytes of device scratch held by the ctx */
XPT_API size_t xpt_scratch_bytes(const xpt_ctx* ctx);
/* 1 if this ctx found no private constant-bank slot for its camera geometry (see Conventions), else 0 */
XPT_API int xpt_geometry_slot_shared(const xpt_ctx* ctx);

/* utils/convert_pose.py:32-71  pose_rvec2matr_batch_tf: [B,N,6] -> [B,N,4,4] */
XPT_API int xpt_pose_rvec2matr(xpt_ctx* ctx, const float* pose, float* matr, void* stream);

/* utils/convert_pose.py:151-168 pose_matr2rvec_batch: `count` 4x4 matrices -> [count,6] = (t, rvec).
 * invert != 0 first applies tf.linalg.inv (losses.py:120 turns stereo_T_LR into T_RL this way).
 * Needs no ctx: `device` is the CUDA ordinal the pointers live on.                                    */
XPT_API int xpt_pose_matr2rvec(int device, const float* matr, int count, int invert, float* rvec, void* stream);

/* losses.py:481-495 StereoPoseLoss.__call__: stereo_T_LR [B,4,4], pose_lr / pose_rl [B,num,6] ->
 * loss_batch [B] (may be NULL).  d_pose_lr / d_pose_rl (may be NULL) receive
 * d(sum_b grad_loss_batch[b] * loss_batch[b]) / d pose_* (grad_loss_batch NULL = all ones).            */
XPT_API int xpt_stereo_pose_loss(int device, const float* stereo_T_LR, const float* pose_lr, const float* pose_rl,
                                 int batch, int num, float* loss_batch, const float* grad_loss_batch,
                                 float* d_pose_lr, float* d_pose_rl, void* stream);

/* utils/util_funcs.py:163-175 multi_scale_like_depth (target pyramid) and
 * synthesize_base.py:74-85 resize_source_images (source pyramid, kept inside
 * the ctx).  target_ms[s] may be NULL (level not wanted); level 0 is a copy.  */
XPT_API int xpt_build_pyramids(xpt_ctx* ctx, const xpt_frames* frames,
                       float* const target_ms[], void* stream);

/* synthesize_base.py:10-29 SynthesizeMultiScale.__call__ (+ the validity mask
 * of bilinear_interp.py:53-76 as an extra output; mask_ms may be NULL).       */
XPT_API int xpt_synthesize(xpt_ctx* ctx, const xpt_frames* frames,
                   const float* const depth_ms[], const float* pose,
                   float* const synth_ms[], float* const mask_ms[], void* stream);

/* Backward of xpt_synthesize for an arbitrary upstream gradient
 * grad_synth_ms[s] = dL/d synth_ms[s] (what tape.gradient does through
 * synthesize_base.py, train_val.py:85).  d_source may be NULL.                */
XPT_API int xpt_synthesize_backward(xpt_ctx* ctx, const xpt_frames* frames,
                            const float* const depth_ms[], const float* pose,
                            const float* const grad_synth_ms[],
                            float* const d_depth_ms[], float* d_pose, float* d_source,
                            void* stream);

/* losses.py:175-195 PhotometricLossMultiScale(method).__call__ on given
 * tensors -> loss_batch [B].  If d_synth_ms != NULL also writes
 * d_synth_ms[s] = d(sum_b grad_loss_batch[b]*loss_batch[b]) / d synth_ms[s]
 * (grad_loss_batch NULL = all ones).                                          */
XPT_API int xpt_photometric_loss(xpt_ctx* ctx, int method,
                         const float* const synth_ms[], const float* const target_ms[],
                         float* loss_batch, const float* grad_loss_batch,
                         float* const d_synth_ms[], void* stream);

/* losses.py:198-232 MonoDepth2LossMultiScale(method).__call__ and, with stereo_synth_ms != NULL,
 * losses.py:282-321 MoALossMultiScale(method).__call__: every scale's synthesis [B,N,h,w,3] (and the
 * stereo synthesis [B,1,h,w,3]) is bilinearly up-sampled to H x W (losses.py:377-383), compared with
 * the FULL-RESOLUTION target [B,H,W,3] per pixel and channel, the minimum over the N (+1) sources is
 * taken and averaged -> loss_batch [B] (scale-merged, losses.py:147-154).
 * If d_synth_ms != NULL also writes d(sum_b grad_loss_batch[b]*loss_batch[b]) / d synth_ms[s]
 * (and / d stereo_synth_ms[s] into d_stereo_synth_ms[s]); tf.reduce_min's rule: ties share equally.  */
XPT_API int xpt_photometric_min_loss(xpt_ctx* ctx, int method,
                             const float* const synth_ms[], const float* const stereo_synth_ms[],
                             const float* target, int64_t target_batch_stride, float* loss_batch,
                             const float* grad_loss_batch, float* const d_synth_ms[],
                             float* const d_stereo_synth_ms[], void* stream);

/* The "L1" and the "SSIM" loss object of one min-over-sources loss set -- moaL1 + moaSSIM or md2L1 + md2SSIM of
 * config-example.py:97-121, which the reference's loop (losses.py:43-47) evaluates one after the other over the same
 * augm_data -- in ONE launch: the up-sampled syntheses, the black-pixel masks and the up-sampling adjoint are shared,
 * the two minima are tracked independently.  loss_batch_l1 / loss_batch_ssim [B] as two calls of
 * xpt_photometric_min_loss would return them.  If d_synth_ms != NULL it receives
 * d(grad_l1 * sum_b loss_batch_l1[b] + grad_ssim * sum_b loss_batch_ssim[b]) / d synth_ms[s] (likewise
 * d_stereo_synth_ms): grad_l1 / grad_ssim are the two losses' upstream weights, loss_weight / global_batch in
 * TotalLoss (losses.py:48-49).  Needs every level at full or at most half resolution (XPT_BAD_SHAPE otherwise).   */
XPT_API int xpt_photometric_min_pair_loss(xpt_ctx* ctx, const float* const synth_ms[],
                                  const float* const stereo_synth_ms[], const float* target,
                                  int64_t target_batch_stride, float* loss_batch_l1, float* loss_batch_ssim,
                                  float grad_l1, float grad_ssim, float* const d_synth_ms[],
                                  float* const d_stereo_synth_ms[], void* stream);

/* losses.py:235-279 CombinedLossMultiScale(method).__call__: every scale's synthesis [B,N,h,w,3] and the
 * flow-warped view `warped` = warped_target_ms[0] [B,N,warped_height,warped_width,3] are bilinearly up-sampled to
 * H x W (losses.py:377-383); per pixel, channel and source the static term counts only where it is SMALLER than
 * the optical-flow term (the mask is a constant: no gradient reaches `warped`); mean over [N,H,W,3] ->
 * loss_batch [B] (scale-merged).  d_synth_ms as for xpt_photometric_min_loss.                                  */
XPT_API int xpt_photometric_cmb_loss(xpt_ctx* ctx, int method, const float* const synth_ms[], const float* warped,
                             int warped_height, int warped_width, const float* target,
                             int64_t target_batch_stride, float* loss_batch, const float* grad_loss_batch,
                             float* const d_synth_ms[], void* stream);

/* cmbL1 + cmbSSIM of one eye (config-example.py:90-96) in ONE launch, as xpt_photometric_min_pair_loss does for the
 * min-over-sources pair: loss_batch_l1 / loss_batch_ssim [B] as two calls of xpt_photometric_cmb_loss would return them,
 * d_synth_ms = d(grad_l1 * sum_b loss_batch_l1[b] + grad_ssim * sum_b loss_batch_ssim[b]) / d synth_ms[s].         */
XPT_API int xpt_photometric_cmb_pair_loss(xpt_ctx* ctx, const float* const synth_ms[], const float* warped,
                                  int warped_height, int warped_width, const float* target,
                                  int64_t target_batch_stride, float* loss_batch_l1, float* loss_batch_ssim,
                                  float grad_l1, float grad_ssim, float* const d_synth_ms[], void* stream);

/* model/synthesize/flow_warping.py:11-49 FlowWarpMultiScale.__call__: flow_ms[s] [B,N,H_s,W_s,2] (the ctx's
 * scales are the FLOW scales, e.g. 4,8,16,32 for PWC-Net, flow_net.py:44-48) -> warped_ms[s] [B,N,H_s,W_s,3]:
 * the source frames resized to the flow's size (flow_warping.py:36-49) and sampled at grid - flow (:51-71) by
 * BilinearInterpolation (bilinear_interp.py:7-147).  mask_ms (may be NULL) receives the validity mask.
 * frames->target and frames->intrinsic are not read.                                                           */
XPT_API int xpt_flow_warp(xpt_ctx* ctx, const xpt_frames* frames, const float* const flow_ms[],
                  float* const warped_ms[], float* const mask_ms[], void* stream);

/* Backward of xpt_flow_warp for an upstream gradient grad_warped_ms[s] = dL/d warped_ms[s]:
 * d_flow_ms[s] [B,N,H_s,W_s,2] (may be NULL) and d_source [B,N,H,W,3] dense (may be NULL; fp32 atomics).       */
XPT_API int xpt_flow_warp_backward(xpt_ctx* ctx, const xpt_frames* frames, const float* const flow_ms[],
                           const float* const grad_warped_ms[], float* const d_flow_ms[], float* d_source,
                           void* stream);

/* losses.py:522-534 L2Regularizer.__call__ ("flow_reg"): loss[0] = sum_i tf.nn.l2_loss(weights[i]) =
 * sum_i sum(weights[i]^2)/2 over `num` dense fp32 device tensors of counts[i] elements (the caller tiles the
 * scalar to [batch]).  With d_weights != NULL also writes d_weights[i] = grad_loss[0] * weights[i]
 * (grad_loss: device pointer to the upstream dL/d loss scalar).                                               */
XPT_API int xpt_l2_regularizer(xpt_ctx* ctx, const float* const weights[], const int64_t counts[], int num,
                       float* loss, const float* grad_loss, float* const d_weights[], void* stream);

/* losses.py:386-440 SmoothenessLossMultiScale.__call__ -> loss_batch [B];
 * optional backward to d_disp_ms as above.                                    */
XPT_API int xpt_smoothness_loss(xpt_ctx* ctx, const float* const disp_ms[],
                        const float* const target_ms[], float* loss_batch,
                        const float* grad_loss_batch, float* const d_disp_ms[], void* stream);

/* losses.py:26-55 TotalLoss.__call__ for the loss set {L1, SSIM, smoothe}
 * (config LOSS_RIGID_T1/T2), forward and -- when any gradient output is
 * non-NULL -- backward in the same call: pyramids, warp, losses, gradients.
 * disp_ms == NULL with a smoothness weight: the disparity is what model/model_wrappers.py:47-48
 * feeds in, safe_reciprocal_number(depth_ms) (utils/util_funcs.py:146-160), formed inside the fused
 * kernel; its gradient is folded into d_depth_ms and d_disp_ms (if given) is zero-filled.  Needs the
 * fused path (not XPT_FLAG_UNFUSED).                                           */
XPT_API int xpt_total_loss(xpt_ctx* ctx, const xpt_frames* frames,
                   const float* const depth_ms[], const float* const disp_ms[],
                   const float* pose, const xpt_loss_outputs* out, void* stream);

/* The same call with HOST buffers (pinned or pageable): copies inputs to
 * device staging owned by the ctx, runs xpt_total_loss, copies back the
 * outputs that are non-NULL, and synchronises before returning.  The batch is
 * cut into up to 8 chunks so that host->device copies, compute and
 * device->host copies of consecutive chunks overlap (copy streams inside the
 * ctx; compute on `stream`).  `frames` and all pointers in `out` are host
 * pointers here; losses of a pipelined call are the sum of the chunk losses
 * (same value up to fp32 rounding of the partial sums).                     */
XPT_API int xpt_total_loss_host(xpt_ctx* ctx, const xpt_frames* frames,
                        const float* const depth_ms[], const float* const disp_ms[],
                        const float* pose, const xpt_loss_outputs* out, void* stream);
/* The same call split in two, so that a training loop can keep TWO steps in flight (two contexts, two streams, two sets
 * of pinned buffers): while one ctx computes and drains step i, the other already copies in step i+1 -- what the
 * reference's tf.data prefetch (tfrecords/tfrecord_reader.py) does for its input pipeline.  _begin enqueues the whole
 * call (one graph launch once the ctx is warm and every buffer is pinned) and returns; _end waits for THAT call only (an
 * event, not the stream) and finishes out->losses.  The buffers passed to _begin must stay untouched until _end returns.
 * Calls that cannot be replayed as a graph (cold ctx, pageable buffers) complete inside _begin.  One call in flight per ctx. */
XPT_API int xpt_total_loss_host_begin(xpt_ctx* ctx, const xpt_frames* frames,
                              const float* const depth_ms[], const float* const disp_ms[],
                              const float* pose, const xpt_loss_outputs* out, void* stream);
XPT_API int xpt_total_loss_host_end(xpt_ctx* ctx);

/* Device timing of the dominant kernel (the fused photometric tile kernel): after
 * xpt_profile_begin, each of the next `max_records` eager launches of that kernel is bracketed
 * by CUDA events on its stream; xpt_profile_end waits for them and returns how many durations
 * (milliseconds) it wrote to ms_out.  Not recorded inside graph replays.          */
XPT_API int xpt_profile_begin(xpt_ctx* ctx, int max_records);
/* which kernel of xpt_total_loss the events bracket (default: the fused tile kernel) */
#define XPT_PROFILE_FUSED 0
#define XPT_PROFILE_PYRAMID 1      /* the single-pass tiled pyramid + geometry kernel (scales within 1,2,4,8) */
XPT_API int xpt_profile_select(xpt_ctx* ctx, int kernel);
XPT_API int xpt_profile_end(xpt_ctx* ctx, float* ms_out, int capacity);

/* Validation of a tensor exchanged through DLPack, for hosts that unwrap capsules themselves: `dl_managed_tensor`
 * points at a DLManagedTensor (dlpack.h v0.8 layout; a void* so that this header needs no DLPack include).
 * Returns XPT_BAD_DTYPE unless float32 x 1 lane, XPT_BAD_DEVICE unless kDLCUDA on `device`,
 * XPT_NOT_CONTIGUOUS unless the dimensions >= dense_from_dim are C-contiguous (leading dimensions may be strided:
 * source = image5d[:, :-1] is legal with dense_from_dim = 2).                                                  */
XPT_API int xpt_check_dlpack(const void* dl_managed_tensor, int device, int dense_from_dim);

/* dst[i][j] = src[i][j] * scale[0] for `num` dense fp32 device tensors of counts[i] elements in ONE launch
 * (`scale` is a DEVICE scalar: the upstream gradient of an autograd node, losses.py:49-54 under
 * train_val.py:85).  dst[i] may equal src[i].                                                                  */
XPT_API int xpt_scale_tensors(int device, const float* const src[], float* const dst[], const int64_t counts[], int num,
                              const float* scale, void* stream);

/* ---- multi-rank: one process per GPU, NCCL over NVLink (reference distributer.py:93-110) ----------------------
 * libnccl.so.2 is bound at run time (dlopen; the copy PyTorch ships is found when torch is loaded).
 * xpt_comm_unique_id: rank 0 creates the 128-byte rendezvous id and hands it to the other ranks (any channel).
 * xpt_comm_init: every rank joins; the ctx owns the communicator.  xpt_comm_attach: borrow an existing ncclComm_t.
 * xpt_allreduce: in-place sum of `num` dense fp32 device buffers as ONE NCCL group on `stream` (the loss vector and
 * the pose / depth net gradient buckets of a step).                                                            */
XPT_API int xpt_comm_unique_id(unsigned char id[128]);
XPT_API int xpt_comm_init(xpt_ctx* ctx, const unsigned char id[128], int nranks, int rank);
XPT_API int xpt_comm_attach(xpt_ctx* ctx, void* nccl_comm);
XPT_API int xpt_allreduce(xpt_ctx* ctx, float* const bufs[], const int64_t counts[], int num, void* stream);
/* Releases the ctx's communicator and peer mappings while the process group is still alive (collective teardown must
 * not be left to interpreter exit).  Synchronises the device and drops the ctx's captured graphs first: a captured
 * step references the communicator.                                                                               */
XPT_API int xpt_comm_destroy(xpt_ctx* ctx);
/* uses_peer_memory: 1 when XPT_FLAG_ALLREDUCE sums the loss scalars INSIDE the epilogue kernel through peer memory
 * (every rank's inbox mapped with CUDA IPC over NVLink) -- no collective launch behind the step; 0 = ncclAllReduce.
 * exchange_error: 1 after a step in which a peer's record did not arrive within ~2 s (ranks issuing different numbers
 * of steps); the losses of that step are invalid.  Synchronising (reads a device flag).                             */
XPT_API int xpt_comm_status(xpt_ctx* ctx, int* uses_peer_memory, int* exchange_error);

/* number of kernels the last call on this ctx launched (bench's gpu_launches) */
XPT_API int xpt_last_launch_count(const xpt_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* XPTWARP_H_ */
