/*
 * ddp_b200.h — C ABI of the B200-native DDP reverse-diffusion decode head.
 *
 * One shared library (libddp_b200.so, sm_100a) owns the whole hot path of
 * JiYuanFeng/DDP: the T-step "noise-to-map" sampling loop and the
 * time-conditioned deformable-attention denoiser it calls.  The Python
 * plug-in classes in ddp_b200/ (DDP, SelfAlignedDDP, DeformableHeadWithTime)
 * bind these entry points with ctypes; INTEGRATION.md shows the stub a
 * reference maintainer would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no C++/torch types.
 *   - Every function returns 0 on success or a negative ddp_status; nothing
 *     throws, nothing calls exit().  ddp_last_error() gives the message.
 *   - x / noise / out / workspace are CALLER-OWNED DEVICE memory (e.g. a
 *     torch tensor's data_ptr()); ddp_sample is asynchronous on `stream`
 *     (a cudaStream_t passed as void*), the caller synchronises.
 *   - After ddp_plan() the library neither allocates nor frees device memory.
 *   - A handle is bound to the CUDA device current at ddp_create(); it is not
 *     thread-safe, distinct handles are independent.
 *
 * Reference interfaces replaced (paths relative to the reference root):
 *   ddp_create / ddp_set_weight   DDP.__init__                       segmentation/mmseg/models/segmentors/ddp.py:57-112
 *                                 DeformableHeadWithTime.__init__    segmentation/mmseg/models/decode_heads/deformable_head_with_time.py:29-60
 *                                 depth DDP.__init__                 depth/depth/models/depther/ddp.py:42-95
 *   ddp_set_schedule              _get_sampling_timesteps, log_snr,  segmentation/mmseg/models/segmentors/ddp.py:14-28,204-213
 *                                 gamma                              depth/depth/models/depther/ddp.py:207-218
 *   ddp_sample                    DDP.ddim_sample                    segmentation/mmseg/models/segmentors/ddp.py:215-246
 *                                 (calls DeformableHeadWithTime.forward deformable_head_with_time.py:90-132,
 *                                  BaseTransformerLayer.forward      segmentation/mmseg/models/utils/transformer.py:374-419,
 *                                  MultiScaleDeformableAttention.forward + ms_deform_attn_forward
 *                                                                     mmcv/ops/multi_scale_deform_attn.py:252-358, 94-151)
 *                                 depth DDP.sample + encode_decode clamp  depth/depth/models/depther/ddp.py:229-247, 97-110
 */
#ifndef DDP_B200_H_
#define DDP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DDP_ABI_VERSION 1

typedef struct ddp_handle ddp_handle;

typedef enum ddp_status {
    DDP_OK = 0,
    DDP_ERR_INVALID = -1,      /* bad argument / bad config (reference: ValueError, asserts)      */
    DDP_ERR_UNSUPPORTED = -2,  /* valid in the reference, not built here (reference: NotImplementedError) */
    DDP_ERR_STATE = -3,        /* call order: weights not committed, plan missing, ...            */
    DDP_ERR_WORKSPACE = -4,    /* workspace too small or misaligned                               */
    DDP_ERR_CUDA = -5,         /* a CUDA call failed; message holds cudaGetErrorString            */
    DDP_ERR_WEIGHT = -6        /* unknown weight name or wrong element count                      */
} ddp_status;

enum { DDP_TASK_SEG = 0, DDP_TASK_DEPTH = 1 };
enum { DDP_SCHEDULE_COSINE = 0, DDP_SCHEDULE_LINEAR = 1 };
enum { DDP_DIFFUSION_DDIM = 0, DDP_DIFFUSION_DDPM = 1 };
/* GEMM arithmetic.  FP32 = CUDA-core fp32 FMA.  TC_3XF16 = tcgen05 fp16 tensor cores with each fp32
 * operand split into hi+lo halves and three MMAs per product (fp32-faithful, ~2^-22).  TC_F16 = one
 * fp16 MMA per product (fast, NOT parity-grade; reported separately).                              */
enum { DDP_GEMM_FP32 = 0, DDP_GEMM_TC_3XF16 = 1, DDP_GEMM_TC_F16 = 2 };

typedef struct ddp_config {
    int32_t abi_version;        /* DDP_ABI_VERSION */
    int32_t task;               /* DDP_TASK_* */
    int32_t num_classes;        /* seg: decode_head.num_classes (<= 256); depth: ignored */
    int32_t timesteps;          /* DDP(timesteps=) */
    int32_t time_difference;    /* DDP(time_difference=) */
    int32_t noise_schedule;     /* DDP_SCHEDULE_* (seg only) */
    int32_t diffusion;          /* DDP_DIFFUSION_* */
    int32_t accumulation;       /* seg: DDP(accumulation=) */
    int32_t learned_sinusoidal_dim; /* 16 */
    int32_t num_layers;         /* encoder.num_layers, 6 in every shipped config (<= 8) */
    int32_t gemm_mode;          /* DDP_GEMM_* */
    float   sample_range_lo;    /* DDP(sample_range=)[0] */
    float   bit_scale;          /* DDP(bit_scale=) */
    float   min_depth;          /* depth */
    float   max_depth;          /* depth */
} ddp_config;

/* Intermediate tensors that tests can copy out (ddp_add_tap) — token-major fp32,
 * [rows][N][width] with rows = B*R, row = b*R + r, token n = i*w + j. */
enum {
    DDP_TAP_HEAD_IN = 0,   /* width 256: transform(cat[x, m_t]) tokens, layer ignored            */
    DDP_TAP_VALUE = 1,     /* width 256: value_proj output of `layer`                            */
    DDP_TAP_SAMPLING = 2,  /* width  96: 64 sampling offsets (head,point,xy) + 32 softmaxed weights */
    DDP_TAP_GATHERED = 3,  /* width 256: ms_deform_attn output before output_proj                */
    DDP_TAP_LN1 = 4,       /* width 256: after first LayerNorm                                   */
    DDP_TAP_LAYER_OUT = 5, /* width 256: layer output (after FiLM)                               */
    DDP_TAP_LOGITS = 6,    /* width C (seg) or 1 (depth: relu(conv)+min_depth), layer ignored    */
    DDP_TAP_STATE = 7,     /* width Cin: m_t / depth_t AFTER the step's update, layer ignored    */
    DDP_TAP_TEMB = 8,      /* [1024] time embedding of `step`                                    */
    DDP_TAP_FILM = 9       /* [512] (scale | shift) of (`step`, `layer`)                         */
};

int ddp_abi_version(void);

/* Create a handle on the current CUDA device.  Fails with DDP_ERR_CUDA when no sm_100 device is
 * usable: there is no CPU fallback. */
int ddp_create(const ddp_config* cfg, ddp_handle** out);
void ddp_destroy(ddp_handle* h);
const char* ddp_last_error(const ddp_handle* h);   /* h may be NULL: last create error */

/* Weights, by the reference's state-dict key, as contiguous fp32 HOST arrays in the reference's
 * own shapes (e.g. "transform.conv.weight" (256,512,1,1), "decode_head.encoder.layers.3.ffns.0.
 * layers.0.0.weight" (1024,256); depth: "down.conv.weight" (256,257,1,1), "decode_head.conv_depth.
 * weight" (1,256,3,3)).  Unknown names (backbone.*, neck.*, auxiliary_head.*) are rejected with
 * DDP_ERR_WEIGHT so the caller can filter; ddp_weight_count/ddp_weight_name enumerate what is needed. */
int ddp_weight_count(const ddp_handle* h);
const char* ddp_weight_name(const ddp_handle* h, int index, int64_t* numel);
int ddp_set_weight(ddp_handle* h, const char* name, const float* host_data, int64_t numel);
/* Upload + repack into the kernels' layouts; fails if any weight is missing. */
int ddp_commit_weights(ddp_handle* h);

/* Optional: override the per-step schedule scalars computed by the library (plain C float math
 * following ddp.py:14-28) with values the host computed, e.g. with the very torch ops the reference
 * uses, so that they are bit-identical to a given reference run.  Arrays of length `timesteps`.
 * seg:   time_in = log_snr(t_now) (input of time_mlp), a_now/s_now = alpha,sigma(t_now),
 *        a_next/s_next = alpha,sigma(t_next).
 * depth: time_in = t_now, a_now = gamma(t_now), a_next = gamma(t_next); s_* ignored. */
int ddp_set_schedule(ddp_handle* h, int timesteps, const float* time_in, const float* a_now,
                     const float* s_now, const float* a_next, const float* s_next);
/* diffusion = DDP_DIFFUSION_DDPM (ddp.py:248-290): per-step scalars (1 - c), c, exp(0.5 log variance) with
 * c = -expm1(log_snr - log_snr_next), and whether t_next > 0 (noise is added).  Optional override like
 * ddp_set_schedule.  ddp_set_step_noise hands over the noise the reference draws with randn_like(mask_t) every step:
 * device fp32 (T, B, R, 256, h, w); it must be set before ddp_sample when diffusion is ddpm. */
int ddp_set_ddpm_schedule(ddp_handle* h, int timesteps, const float* one_minus_c, const float* c, const float* std_dev,
                          const int32_t* noise_on);
int ddp_set_step_noise(ddp_handle* h, const float* device_noise);
/* Per-pixel uncertainty maps (README "uncertainty awareness").  The reference exposes its stochastic
 * samples only through randsteps + the mean (ddp.py:219, 245); these two optional extra outputs of the following
 * ddp_sample / ddp_sample_host calls are DEFINED here and pinned by oracle/ddp_oracle.py::uncertainty:
 *   changes (B,h,w) int32, seg only: sum over the R samples of the number of steps k >= 1 whose argmax class differs
 *                                    from the same sample's class at step k-1;
 *   spread  (B,h,w) fp32: seg   1 - (number of samples whose LAST-step class equals the returned map's class) / R,
 *                         depth the population standard deviation over the R samples of the last-step prediction.
 * Device pointers, caller-owned, either may be NULL; they stay registered until changed.  Cost: one byte per (sample,
 * token) and step inside the existing step kernel — no extra launch. */
int ddp_set_uncertainty_outputs(ddp_handle* h, int32_t* changes, float* spread);
/* Read back the schedule in use (after ddp_plan). */
int ddp_get_schedule(const ddp_handle* h, float* time_in, float* a_now, float* s_now,
                     float* a_next, float* s_next);

/* Fix the geometry: B images, R stochastic samples per image (randsteps), h x w tokens.  Builds the
 * shape-only constants (sine positional encoding, its products with the offset/attention
 * projections, time embeddings and FiLM vectors of all steps) and reports the workspace size. */
int ddp_plan(ddp_handle* h, int B, int R, int height, int width, size_t* workspace_bytes);

/* The sampling loop.
 *   x      (B,256,h,w) fp32 NCHW   frozen conditioning feature (neck output)
 *   noise  (B,R,Cin,h,w) fp32      initial state (ddp.py:220 draws it; here the caller does), Cin = 256 seg / 1 depth
 *   out    (B,C,h,w) fp32          seg: mean logits, or mean softmax prob when accumulation; depth: (B,1,h,w) clamped
 *   cls    (B,h,w) int32 or NULL   seg: argmax_C of `out`
 * workspace: >= ddp_plan's size, 256-byte aligned. */
int ddp_sample(ddp_handle* h, const float* x, const float* noise, float* out, int32_t* cls,
               void* workspace, size_t workspace_bytes, void* stream);

/* One evaluation of the denoiser, for callers of decode_head.forward(inputs, times)
 * (segmentation/mmseg/models/decode_heads/deformable_head_with_time.py:90-132):
 *   feat            (rows,256,h,w) fp32 NCHW device, rows = B*R of ddp_plan (the transform()/down() output)
 *   time_embedding  [1024] fp32 device (one embedding for all rows, as in the sampling loop)
 *   out             seg: (rows,C,h,w) logits; depth: (rows,1,h,w) = relu(conv3x3)+min_depth
 * Uses (and clobbers) the same workspace as ddp_sample. */
int ddp_head_forward(ddp_handle* h, const float* feat, const float* time_embedding, float* out, void* workspace,
                     size_t workspace_bytes, void* stream);

/* Post-loop tail in one kernel: bilinear resize (align_corners=False) of (B,C,h,w) logits to (out_h,out_w),
 * softmax, argmax -> uint8 class map (B,out_h,out_w).  Replaces resize + F.softmax + argmax of
 * segmentation/mmseg/models/segmentors/ddp.py:124-128 and encoder_decoder.py:232-304 for the no-flip, whole-image case. */
int ddp_resize_argmax(ddp_handle* h, const float* logits, int B, int C, int in_h, int in_w, int out_h, int out_w,
                      uint8_t* cls, void* stream);

/* The general inference() tail of one view, for flipped / rescaled / multi-scale test-time augmentation
 * (encoder_decoder.py:229-283 whole_inference + inference, ddp.py:124-128; aug_test :295-304 sums the views):
 *   logits (B,C,in_h,in_w) --bilinear--> (img_h,img_w) --[rescale: crop to (crop_h,crop_w) = img_meta img_shape, bilinear to
 *   (out_h,out_w) = ori_shape]--> softmax over C --> flip back (1 horizontal, 2 vertical) --> probs (B,C,out_h,out_w) fp32,
 * overwritten (accumulate = 0) or added to (accumulate = 1).  Both resizes use align_corners=False and are evaluated
 * nested in one kernel with the reference's rounding of the intermediate image.  ddp_probs_argmax turns (accumulated)
 * probabilities into the uint8 class map. */
int ddp_tail_probs(ddp_handle* h, const float* logits, int B, int C, int in_h, int in_w, int img_h, int img_w, int crop_h,
                   int crop_w, int out_h, int out_w, int rescale, int flip, int accumulate, float* probs, void* stream);
int ddp_probs_argmax(ddp_handle* h, const float* probs, int B, int C, int H, int W, uint8_t* cls, void* stream);

/* Same call with HOST buffers (pinned or pageable): copies x and noise to the device, runs
 * ddp_sample, copies out (and cls) back and synchronises the stream.  Uses the tail of the workspace
 * for staging (ddp_plan's size already includes it). */
int ddp_sample_host(ddp_handle* h, const float* x_host, const float* noise_host, float* out_host,
                    int32_t* cls_host, void* workspace, size_t workspace_bytes, void* stream);
/* The same with two more arguments.  `chunks`: the batch is cut into that many contiguous groups of whole images which are
 * pipelined (H2D of group c+1 and D2H of group c-1 on two copy streams under the loop of group c on `stream`); 0 = automatic
 * (environment DDP_B200_HOST_CHUNKS, else 2 for inputs >= 64 MB, else 1).  Images are independent, so the result does not
 * depend on `chunks`.  `out_device`: optional (B,C,h,w) device buffer that also receives the result — e.g. the send
 * buffer of the one NCCL gather — instead of the workspace's staging area.  Pinned host memory is needed for
 * the copies to overlap; pageable memory works but serialises. */
int ddp_sample_host_ex(ddp_handle* h, const float* x_host, const float* noise_host, float* out_host, int32_t* cls_host,
                       float* out_device, int chunks, void* workspace, size_t workspace_bytes, void* stream);
/* Streaming form for a queue of batches (a serving process, an evaluation loop over a dataset — the reference's
 * single_gpu_test / multi_gpu_test loops, segmentation/mmseg/apis/test.py:91, 208): submit returns at once with a ticket,
 * at most TWO calls may be in flight (two staging sets inside the workspace); the upload of call i+1 and the download of
 * call i-1 run on copy streams under the loop of call i.  ddp_sample_host_wait blocks until that call's out_host (and
 * cls_host) are complete.  x_host / noise_host must stay valid and unchanged until the NEXT submit on the same handle has
 * been waited for or this call's wait returned, out_host until this call's wait returned.  Results are bit-identical to
 * ddp_sample_host.  Pinned host memory required for the overlap. */
int ddp_sample_host_submit(ddp_handle* h, const float* x_host, const float* noise_host, float* out_host, int32_t* cls_host,
                           float* out_device, void* workspace, size_t workspace_bytes, void* stream, int64_t* ticket);
int ddp_sample_host_wait(ddp_handle* h, int64_t ticket);

/* Test hooks.  A tap copies one intermediate of (step, layer) into caller-owned device memory during
 * the next ddp_sample calls; ddp_set_state_override makes step `step` start from the given state
 * ((B,R,Cin,h,w) device pointer, NCHW like `noise`) — teacher forcing.  ddp_clear_debug removes both. */
int ddp_add_tap(ddp_handle* h, int kind, int step, int layer, float* device_dst);
int ddp_set_state_override(ddp_handle* h, int step, const float* device_state);
int ddp_clear_debug(ddp_handle* h);

/* Number of kernel launches the last ddp_sample enqueued (for bench.py's gpu_launches). */
int64_t ddp_last_launch_count(const ddp_handle* h);

/* Latency mode (environment DDP_B200_GRAPH=1 at ddp_create; the reference's own deployment is one image per GPU,
 * segmentation/tools/test.py:216, tools/benchmark.py:62): ddp_sample's launch sequence is captured into a CUDA graph on the
 * second call with one buffer set + caller stream and replayed afterwards.  ddp_graph_replays = calls served by
 * cudaGraphLaunch so far, ddp_graph_captures = graphs instantiated so far, ddp_graph_last_fallback = why the LAST
 * ddp_sample used ordinary launches ("" when it replayed).  A capture / instantiate failure switches the mode off for the
 * handle and says so once on stderr. */
int64_t ddp_graph_replays(const ddp_handle* h);
int64_t ddp_graph_captures(const ddp_handle* h);
const char* ddp_graph_last_fallback(const ddp_handle* h);

/* Per-kernel-class device timing: while enabled, every launch of ddp_sample is bracketed by CUDA
 * events on the launching stream; ddp_profile_collect waits for them and returns, per class, the
 * summed milliseconds and the number of launches since the last collect. */
enum {
    DDP_K_COND = 0, DDP_K_HEAD_IN, DDP_K_VALUE, DDP_K_SAMPLING, DDP_K_GATHER, DDP_K_OUT_PROJ,
    DDP_K_FFN1, DDP_K_FFN2, DDP_K_HEAD_OUT, DDP_K_STEP, DDP_K_FINALIZE, DDP_K_LAYOUT, DDP_K_FFN_FUSED, DDP_K_QPROJ_FUSED, DDP_K_COUNT
};
int ddp_profile_enable(ddp_handle* h, int on);
int ddp_profile_collect(ddp_handle* h, float* ms_by_class, int64_t* launches_by_class, int n_classes);
const char* ddp_kernel_class_name(int cls);

/* ------------------------------------------------------------------------------------------------------------------
 * The neck in front of the loop: FPN + MultiStageMerging, the module pair every DDP config chains
 * between backbone and decode head.  It produces `x`, the frozen conditioning feature ddp_sample takes.
 *
 * Reference interfaces replaced:
 *   ddp_neck_create / set_weight   FPN.__init__                   segmentation/mmseg/models/necks/fpn.py:66-160
 *                                  MultiStageMerging.__init__     segmentation/mmseg/models/necks/multi_stage_merging.py:14-37
 *   ddp_neck_forward               FPN.forward                    segmentation/mmseg/models/necks/fpn.py:162-213
 *                                  MultiStageMerging.forward      segmentation/mmseg/models/necks/multi_stage_merging.py:40-52
 *                                  (depth/depth/models/necks/{fpn,multi_stage_merging}.py are copies)
 * Supported arguments = what the DDP configs pass: out_channels 256, GroupNorm after every conv (no conv bias),
 * act_cfg None, num_outs == number of inputs, start_level 0, nearest top-down upsampling, bilinear
 * align_corners=False merging; in_channels multiples of 16.  Same conventions as above (caller-owned device memory,
 * asynchronous on `stream`, no allocation after ddp_neck_commit_weights, status codes, no CPU path).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct ddp_neck ddp_neck;

enum { DDP_NECK_STAGE_FPN = 1, DDP_NECK_STAGE_MERGE = 2 };   /* stages bit mask: 1 = FPN only, 2 = MultiStageMerging only, 3 = fused */

typedef struct ddp_neck_config {
    int32_t abi_version;        /* DDP_ABI_VERSION */
    int32_t stages;             /* DDP_NECK_STAGE_* bit mask */
    int32_t num_levels;         /* len(in_channels) == num_outs, <= 4 */
    int32_t in_channels[4];     /* FPN(in_channels=); ignored (256 each) when stages == MERGE */
    int32_t out_channels;       /* 256 */
    int32_t num_groups;         /* norm_cfg=dict(type='GN', num_groups=32) */
    float   eps;                /* 1e-5 (nn.GroupNorm default) */
} ddp_neck_config;

int ddp_neck_create(const ddp_neck_config* cfg, ddp_neck** out);
void ddp_neck_destroy(ddp_neck* h);
const char* ddp_neck_last_error(const ddp_neck* h);   /* h may be NULL: last create error */

/* Weights by the reference's state-dict key, contiguous fp32 HOST arrays in the reference's shapes.  The key may carry
 * the module path in front ("neck.0.lateral_convs.2.conv.weight", "0.lateral_convs.2.conv.weight" and
 * "lateral_convs.2.conv.weight" are the same tensor): lateral_convs.{l}.conv.weight (256,C_l,1,1),
 * fpn_convs.{l}.conv.weight (256,256,3,3), {lateral_convs,fpn_convs}.{l}.gn.{weight,bias} (256), down.conv.weight
 * (256,256*L,1,1), down.gn.{weight,bias}. */
int ddp_neck_weight_count(const ddp_neck* h);
const char* ddp_neck_weight_name(const ddp_neck* h, int index, int64_t* numel);
int ddp_neck_set_weight(ddp_neck* h, const char* name, const float* host_data, int64_t numel);
int ddp_neck_commit_weights(ddp_neck* h);

/* Fix the geometry: B images, level l is heights[l] x widths[l] (level 0 = the 1/4-resolution map the decode loop runs on). */
int ddp_neck_plan(ddp_neck* h, int B, const int32_t* heights, const int32_t* widths, size_t* workspace_bytes);

/*   inputs[l]    (B,C_l,h_l,w_l) fp32 NCHW device      backbone pyramid (stages & FPN) or FPN outputs (stages == MERGE)
 *   x_out        (B,256,h_0,w_0) fp32 NCHW device      MultiStageMerging output; may be NULL when stages == FPN
 *   fpn_outs[l]  (B,256,h_l,w_l) fp32 NCHW device      FPN outputs; the array or single entries may be NULL unless stages == FPN
 * workspace: >= ddp_neck_plan's size, 256-byte aligned. */
int ddp_neck_forward(ddp_neck* h, const float* const* inputs, float* x_out, float* const* fpn_outs, void* workspace,
                     size_t workspace_bytes, void* stream);
int64_t ddp_neck_last_launch_count(const ddp_neck* h);

/* ------------------------------------------------------------------------------------------------------------------
 * BEV map segmentation: the same denoiser inside the BEV tree's sampling loop.
 *
 * Reference interfaces replaced (paths relative to bev/):
 *   ddp_bev_create / set_weight    DDP.__init__                    mmdet3d/models/fusion_models/ddp.py:65-116
 *                                  DeformableHeadWithTime.__init__ mmdet3d/models/heads/segm/deformable_head_with_time.py:109-142
 *   ddp_bev_plan                   BEVGridTransform                mmdet3d/models/heads/segm/deformable_head_with_time.py:58-98
 *   ddp_bev_sample                 DDP.ddim_sample                 mmdet3d/models/fusion_models/ddp.py:268-301
 *                                  (calls DeformableHeadWithTime.forward, heads/segm/deformable_head_with_time.py:178-241)
 * Two token grids: the state m_t lives on the fused BEV feature grid (in_h x in_w), the denoiser runs on the output map
 * grid (out_h x out_w) after a bilinear grid_sample; 6 sigmoid classes thresholded at `threshold`; the result is the mean
 * of the sigmoid maps over all timesteps * randsteps evaluations.  The reference's ddpm_sample for this model cannot
 * run (it indexes the repeated feature with [0] and feeds floats to nn.Embedding, ddp.py:303-342): only ddim is built.
 * Same conventions as ddp_sample (caller-owned device memory, asynchronous on `stream`, status codes, no CPU path).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct ddp_bev ddp_bev;

typedef struct ddp_bev_config {
    int32_t abi_version;            /* DDP_ABI_VERSION */
    int32_t timesteps;              /* DDP(timesteps=) */
    int32_t time_difference;        /* DDP(time_difference=) */
    int32_t noise_schedule;         /* DDP_SCHEDULE_* */
    int32_t diffusion;              /* DDP_DIFFUSION_DDIM */
    int32_t learned_sinusoidal_dim; /* 16 */
    int32_t num_layers;             /* heads.map.encoder.num_layers, 5 in the shipped configs */
    int32_t feat_channels;          /* DDP(feat_channels=): 256 camera-only, 512 (default) fusion */
    int32_t gemm_mode;              /* DDP_GEMM_* of the denoiser */
    float   bit_scale;              /* DDP(bit_scale=) */
    float   threshold;              /* DDP(threshold=), 0.5 */
} ddp_bev_config;

int ddp_bev_create(const ddp_bev_config* cfg, ddp_bev** out);
void ddp_bev_destroy(ddp_bev* h);
const char* ddp_bev_last_error(const ddp_bev* h);     /* h may be NULL: last create error */

/* Weights by the BEV model's state-dict keys: embedding_table.weight (7,256), transform.conv.weight
 * (256, feat_channels + 256, 1, 1), transform.conv.bias, time_mlp.*, heads.map.encoder.layers.N.*, heads.map.conv_seg.*. */
int ddp_bev_weight_count(const ddp_bev* h);
const char* ddp_bev_weight_name(const ddp_bev* h, int index, int64_t* numel);
int ddp_bev_set_weight(ddp_bev* h, const char* name, const float* host_data, int64_t numel);
int ddp_bev_commit_weights(ddp_bev* h);

/* Optional schedule override, same meaning as ddp_set_schedule (seg). */
int ddp_bev_set_schedule(ddp_bev* h, int timesteps, const float* time_in, const float* a_now, const float* s_now,
                         const float* a_next, const float* s_next);

/* Geometry.  grid_y [out_h] / grid_x [out_w]: HOST arrays of the normalised ([-1, 1]) sampling coordinates
 * BEVGridTransform computes from its input_scope / output_scope (deformable_head_with_time.py:82-88); the host computes
 * them with the reference's own ops so that they are bit-identical to a reference run. */
int ddp_bev_plan(ddp_bev* h, int B, int R, int in_h, int in_w, int out_h, int out_w, const float* grid_y, const float* grid_x,
                 size_t* workspace_bytes);

/*   x      (B,feat_channels,in_h,in_w) fp32 NCHW device    fused BEV feature (decoder neck output)
 *   noise  (B,R,256,in_h,in_w) fp32 device                  initial state (ddp.py:275 draws it; here the caller does)
 *   out    (B,6,out_h,out_w) fp32 device                    mean sigmoid map */
int ddp_bev_sample(ddp_bev* h, const float* x, const float* noise, float* out, void* workspace, size_t workspace_bytes,
                   void* stream);
int64_t ddp_bev_last_launch_count(const ddp_bev* h);

#ifdef __cplusplus
}
#endif
#endif /* DDP_B200_H_ */
