/*
 * clslam_b200 -- C ABI of the B200-native CL-SLAM online-adaptation hot path.
 *
 * The reference (robot-learning-freiburg/CL-SLAM) has no FFI/plugin layer: its boundary for
 * this path is the Python class depth_pose_prediction.DepthPosePrediction
 * (depth_pose_prediction/depth_pose_prediction.py:38).  This header is what the drop-in
 * Python class in clslam_b200/ binds with ctypes; every entry point names the reference
 * code it replaces (file:line relative to /root/reference).
 *
 * Conventions
 *   - plain C types only: raw DEVICE pointers, ints, floats, a cudaStream_t passed as void*.
 *   - no allocation, no ownership transfer: the caller (torch) owns every buffer; kernels that
 *     need scratch take (workspace, workspace_bytes) and a *_workspace_bytes() query exists.
 *   - stream ordered, no host synchronisation inside, CUDA-graph capturable.
 *   - return 0 on success, negative on error; clslam_last_error() returns the message of the
 *     last failing call on this host thread.
 *   - activations are fp32 NHWC ("pixel-major, channel-fastest"); images / disparities /
 *     depth maps at the API edge are fp32 NCHW exactly as in the reference's batch dict
 *     (SURVEY.md App. C).  Convolution weights are fp32 [Cout][kh][kw][Cin] (KRSC), i.e. the
 *     torch.channels_last storage of the reference's (Cout,Cin,kh,kw) tensors.
 */
#ifndef CLSLAM_B200_H
#define CLSLAM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define CLSLAM_ABI_VERSION 1

/* epilogue activations */
enum { CLSLAM_ACT_NONE = 0, CLSLAM_ACT_RELU = 1, CLSLAM_ACT_ELU = 2, CLSLAM_ACT_SIGMOID = 3 };
/* border handling of the convolution input */
enum { CLSLAM_PAD_ZERO = 0, CLSLAM_PAD_REFLECT = 1 };
/* how the convolution input is assembled */
enum { CLSLAM_IN_PLAIN = 0,   /* x: (n, hi, wi, ci) */
       CLSLAM_IN_UPCAT = 1 }; /* cat[nearest_up(x -> hi x wi) (c1 ch @ h1 x w1), x2 (ci-c1 ch @ hi x wi)] */
/* disp -> depth (depth_pose_prediction/utils.py:120-142) */
enum { CLSLAM_DEPTH_INV = 0,      /* 1/disp            (min_depth=None, max_depth=None) */
       CLSLAM_DEPTH_MIN_ONLY = 1, /* min_depth/disp    (max_depth=None; config_adapt.yaml:23-24) */
       CLSLAM_DEPTH_MIN_MAX = 2 };/* 1/(1/max + (1/min-1/max)*disp) */

/* One convolution layer.  Forward: y = act(scale * conv(x) + shift) (+ residual before act). */
typedef struct clslam_conv_desc {
  int n;            /* batch */
  int hi, wi, ci;   /* (virtual) input height, width, channels (ci = c1 + c2 for UPCAT) */
  int ho, wo, co;   /* output height, width, channels */
  int kh, kw, stride, pad;
  int pad_mode;     /* CLSLAM_PAD_* */
  int in_mode;      /* CLSLAM_IN_* */
  int c1, h1, w1;   /* UPCAT: channels / height / width of the low-resolution part */
  int act;          /* CLSLAM_ACT_* of THIS layer's output */
} clslam_conv_desc;

const char* clslam_last_error(void);
int clslam_abi_version(void);
/* number of kernel launches issued through this library since load (bench.py "gpu_launches") */
long long clslam_launch_count(void);
/* bind this library's CUDA runtime to `device` on the calling host thread (one process per GPU) */
int clslam_set_device(int device);
/* Timing events for the measurement harness (bench.py): cudaEvent_t handles.  clslam_event_record on a stream that is
 * being captured adds an external event-record node, so the per-kernel durations of the step graph can be read after a
 * replay (warm, in-graph: the time base of the roofline block).  Not used on the product path. */
void* clslam_event_create(void);
int clslam_event_destroy(void* ev);
int clslam_event_record(void* ev, void* stream);
int clslam_event_elapsed_ms(void* ev0, void* ev1, float* ms);

/* ---- input staging ------------------------------------------------------------------------
 * (x - 0.45) / 0.225 of networks/resnet_encoder.py:117 fused with NCHW -> NHWC and channel
 * padding; `img_b` may be NULL (depth net: 3 -> c_pad channels) or a second image that is
 * concatenated on the channel axis (pose net, depth_pose_prediction.py:951-955). */
int clslam_stage_images(const float* img_a, const float* img_b, float* out_nhwc, int n, int h, int w,
                        int c_pad, void* stream);

/* ---- convolution stack --------------------------------------------------------------------
 * Replaces nn.Conv2d + BatchNorm2d(eval) + ReLU / residual add of torchvision BasicBlock as
 * driven by networks/resnet_encoder.py:118-124, Conv3x3/ConvBlock (networks/layers.py:9-48),
 * the nearest-upsample + concat of networks/depth_decoder.py:56-66 (fused into the loader of
 * upconv_i_1) and PoseDecoder convs (networks/pose_decoder.py:40-47).
 *   scale may be NULL (=1); shift is the bias or the folded BatchNorm shift; residual may be NULL. */
/* arithmetic of the three GEMM-shaped entry points below:
 *   1 (default) = tcgen05 tensor cores, operands split on the fly into bf16 hi+lo, three products per k step
 *                 (A_lo.B_hi + A_hi.B_lo + A_hi.B_hi), fp32 accumulation in TMEM;
 *   0           = exact fp32 on the CUDA cores (parity anchor). */
int clslam_set_conv_mode(int mode);
int clslam_get_conv_mode(void);
/* debug aid: clock64 stamps of CTA 0 of the last tensor-core GEMM launch (synchronises the device) */
int clslam_debug_tc_stamps(unsigned long long* out16);
long long clslam_conv_fwd_workspace_bytes(const clslam_conv_desc* d);
int clslam_conv_fwd(const clslam_conv_desc* d, const float* x, const float* x2, const float* w,
                    const float* scale, const float* shift, const float* residual, float* y,
                    void* workspace, long long workspace_bytes, void* stream);

/* Same layer as clslam_conv_fwd with the weights additionally given as pre-split bf16 planes
 * (w = w_hi + w_lo up to ~2^-17, [Cout][kh][kw][Cin] each, produced by clslam_split_bf16): 3x3/stride-1
 * layers with Cin % 64 == 0 then run the patch-staged tcgen05 kernel, everything else falls back to
 * clslam_conv_fwd with the fp32 weights `w`. */
int clslam_split_bf16(const float* src, void* hi, void* lo, long long n, void* stream);
int clslam_conv_fwd_bf16w(const clslam_conv_desc* d, const float* x, const float* x2, const float* w, const void* w_hi,
                          const void* w_lo, const float* scale, const float* shift, const float* residual, float* y,
                          void* workspace, long long workspace_bytes, void* stream);

/* ---- bf16 hi/lo activation planes + TMA-fed tcgen05 3x3 convolution -------------------------
 * A "planes" tensor stores an activation x as two bf16 tensors hi = bf16(x), lo = bf16(x - hi)
 * (x = hi + lo up to ~2^-17): exactly the operands of the 3xBF16 tensor-core products, written
 * once by the producing kernel's epilogue, so that consumers feed them to tcgen05.mma with
 * plain TMA loads (no CUDA-core staging).  Layout: [2 planes][n][h + 2][w + 2][c_total] (channels
 * fastest) with a one-pixel halo ring: reflection-padded consumers (networks/layers.py:36,
 * nn.ReflectionPad2d(1)) read the stored halo, zero-padded consumers (torchvision BasicBlock,
 * pose decoder) address the interior and let TMA zero-fill outside.  A channel slice
 * [c_off, c_off + c) of a wider buffer realises torch.cat (networks/depth_decoder.py:66)
 * without a copy; writing with up = 2 realises F.interpolate(nearest) (:61-64). */
typedef struct clslam_planes {
  void* base;              /* bf16 element (hi plane, image 0, padded row 0, padded col 0, channel 0) */
  long long plane_stride;  /* elements between the hi and the lo plane */
  int n, h, w;             /* images, logical height and width (stored (h + 2) x (w + 2)) */
  int c_total;             /* channels per stored pixel */
  int c_off, c;            /* channel slice this operand refers to */
} clslam_planes;

enum { CLSLAM_PAD_FULL = 2 };  /* dgrad: "full" correlation, output grid (h + 2) x (w + 2), zero outside */

/* y = act(scale * conv3x3_s1(in) + shift (+ residual)), implicit GEMM on tcgen05 (M = 16 x 8 output
 * pixels, N = cout block, K = 9 * cin in (tap, channel chunk) steps), operands by TMA, fp32
 * accumulators double-buffered in TMEM, persistent CTAs.  Outputs (each optional): fp32 NHWC `y`
 * (n, ho, wo, cout) and/or planes `out` (logical dims (ho * out_up, wo * out_up); out_halo = 1 also
 * writes the reflection halo).  Needs in.c % 16 == 0 and cout % 16 == 0. */
typedef struct clslam_conv_tma {
  clslam_planes in;
  int pad_mode;            /* CLSLAM_PAD_ZERO | CLSLAM_PAD_REFLECT | CLSLAM_PAD_FULL */
  int cout;
  const void* w_hi;        /* bf16 [cout][3][3][in.c] */
  const void* w_lo;
  const float* scale;      /* [cout] or NULL */
  const float* shift;      /* [cout] or NULL */
  int act;                 /* CLSLAM_ACT_* */
  const clslam_planes* residual;  /* optional, logical dims of the output */
  const float* gate;       /* optional fp32 NHWC (n, ho, wo, cout): result *= act'(gate), dgrad through the activation */
  int gate_act;            /* CLSLAM_ACT_* whose derivative (expressed through its output) is applied */
  float* y;                /* optional */
  const clslam_planes* out;       /* optional */
  int out_up;              /* 1 or 2 (nearest up-sampling on write) */
  int out_halo;            /* write the reflection halo of `out` */
  int taps;                /* 0 / 3: 3x3.  1 (experimental): 1x1 kernel, weights [cout][in.c] (with `in` = the first c channels of
                              space-to-depth planes: a 1x1 / stride-2 convolution, torchvision BasicBlock.downsample).
                              2: 2x2 kernel over input rows / cols o-1 .. o (space-to-depth form of a 3x3 / stride-2 conv).
                              4: 4x4 kernel over input rows / cols o-2 .. o+1, zero padded, weights [cout][4][4][in.c]
                              (the space-to-depth form of the encoders' 7x7 / stride-2 conv1, clslam_stage_images_s2d) */
  int flags;               /* CLSLAM_TMA_* bits, 0 = none */
  int* splitk_counters;    /* optional, clslam_conv3x3_tma_ws only: CLSLAM_TMA_SPLITK_COUNTERS zero-initialised ints owned by THIS
                              launch site.  When the launch splits K, the CTA that delivers the last K slice of a tile adds the
                              slices (fixed order, deterministic) and runs the epilogue itself: no second kernel.  The kernel
                              leaves the counters at zero.  NULL: separate reduction / epilogue kernel (round-1 form). */
  int sm_budget;           /* 0 = every SM.  > 0: at most this many CTAs (a CTA of this kernel family holds a whole SM: ~200 KB of
                              shared memory), for launches that run NEXT TO another lane's launches -- two concurrent launches only
                              overlap when their CTA counts add up to <= the SM count.  Applies to the persistent grid and to the
                              split-K plan (same value in the workspace query and the launch). */
} clslam_conv_tma;
#define CLSLAM_TMA_SPLITK_COUNTERS 256
/* flags bit (experimental): the stored halo ring of `in` is valid padding for this layer -- the reflection halo for
 * CLSLAM_PAD_REFLECT (always required), all zeros for CLSLAM_PAD_ZERO (true when the producer wrote the planes with
 * out_halo = 0 into a zero-initialised buffer).  The kernel may then walk the n images as ONE stack of n * (h + 2) rows, so
 * that a 16-row tile covers several small images (6 x 20 maps: half the tiles) instead of padding each image to 16 rows. */
#define CLSLAM_TMA_IN_HALO_VALID 1
/* flags bit, clslam_conv3x3_tma_ws / clslam_dgrad3x3_tma_ws only: when the launch splits K, run the (at most 8) K slices of a
 * tile as the CTAs of one thread-block cluster and reduce them over distributed shared memory: no workspace round trip, no
 * second kernel, the epilogue spread over the cluster.  Deterministic (fixed rank order). */
#define CLSLAM_TMA_SPLITK_CLUSTER 2
int clslam_conv3x3_tma(const clslam_conv_tma* p, void* stream);
/* Split-K form for the layers whose output grid is too small to fill the persistent kernel (the <= 24 x 80 maps of
 * ResNet layer2-4, upconv_4_x .. upconv_2_0 and the pose decoder at batch 1): when the layer has fewer (16 x 8 tile,
 * N block) work items than half the SMs, the (64-channel chunk, column shift) steps of K = 9 * cin are spread over
 * ksplit work items per tile, each storing raw fp32 partial sums to `workspace` [ksplit][n][ho][wo][cout]; a second
 * kernel adds them in a fixed order (deterministic) and applies the same epilogue.  Same arguments, same results up to
 * fp32 summation order; layers that are large enough (or whose shape the split kernel is not built for: it needs
 * in.c % 64 == 0, cout % 64 == 0, taps 3 or 2) run exactly clslam_conv3x3_tma and need no workspace (query returns 0). */
long long clslam_conv3x3_tma_workspace_bytes(const clslam_conv_tma* p);
int clslam_conv3x3_tma_ws(const clslam_conv_tma* p, void* workspace, long long workspace_bytes, void* stream);
/* dX of a reflection-padded layer in one call: clslam_conv3x3_tma_ws(p) with p->pad_mode = CLSLAM_PAD_FULL and the fp32
 * scratch p->y (n, h + 2, w + 2, cout), followed by clslam_dgrad_fold(p->y, p->in.n, p->in.h, p->in.w, p->cout, up, ...).
 * When the launch is split over K the fold reads the partial sums of the workspace directly (it adds the slices and the
 * mirrored halo cells in one pass: no separate split-K epilogue, p->y untouched).  Workspace: clslam_conv3x3_tma_workspace_bytes(p). */
int clslam_dgrad3x3_tma_ws(const clslam_conv_tma* p, int up, const float* y_prev, int act_prev, const float* disp_dz,
                           const float* disp_w, float* dz_prev, const clslam_planes* dz_prev_planes, void* workspace,
                           long long workspace_bytes, void* stream);
/* dgrad of the 3x3 decoder convolutions on the same kernel (autograd of depth_pose_prediction.py:312):
 *   clslam_wt_flip: Wt[ci'][2-r][2-s][co] = W[co][r][s][ci'] (ci' < n_ci) for a batch of layers described by a DEVICE
 *     array of items {long long src, dst (element offsets into the planes); int co, ci, n_ci, pad}; planes are bf16 with
 *     the lo plane `*_plane_stride` elements after the hi plane.
 *   dX of a zero-padded layer = clslam_conv3x3_tma(dz planes, PAD_ZERO, Wt, gate = y_prev): one launch.
 *   dX of a reflection-padded layer = clslam_conv3x3_tma(dz planes, PAD_FULL, Wt) -> fp32 (n, h+2, w+2, c), then
 *   clslam_dgrad_fold: adds the halo cells onto the pixels they mirror, sums the 2x2 footprint when the layer read
 *     an up-sampled tensor (up = 2), adds the dispconv input gradient, multiplies by act'(y_prev). */
int clslam_wt_flip(const void* items_dev, int n_items, int max_elems, const void* w_hi, long long w_plane_stride, void* wt_hi,
                   long long wt_plane_stride, void* stream);
int clslam_dgrad_fold(const float* dxp, int n, int h, int w, int c, int up, const float* y_prev, int act_prev,
                      const float* disp_dz, const float* disp_w, float* dz_prev, const clslam_planes* dz_prev_planes,
                      void* stream);
/* wgrad of a 3x3 / stride-1 layer on tcgen05 with TMA-fed MN-major operands (no transposition): dW [co][3][3][ci] and
 * db [co] (may be NULL; reduced from the fp32 copy dz_f32 (n, h, w, co)) from the layer input planes `x` (pad_mode ZERO: interior view, REFLECT: stored halo) and the
 * output-gradient planes `dz` (zero halo), both (n, h, w).  Needs x->c % 32 == 0 and dz->c in {16, 32, 64, 128 * k}, or
 * x->c == dz->c == 16 (the last decoder layer).  Deterministic: per-CTA partial sums + fixed-order reduction. */
long long clslam_wgrad3x3_tma_workspace_bytes(int n, int h, int w, int ci, int co);
int clslam_wgrad3x3_tma(const clslam_planes* x, int pad_mode, const clslam_planes* dz, const float* dz_f32, float* dw, float* db,
                        void* workspace, long long workspace_bytes, void* stream);
/* timing experiments only (results are wrong when != 0; the tests are compiled out unless the library is built with
 * -DCLSLAM_TMA_DEBUG): 1 no MMA, 2 no epilogue body, 4 no weight ring, 8 no A loads */
int clslam_debug_tma_flags(int flags);
/* debug builds only (-DCLSLAM_TMA_DEBUG, tools/tma_stamps.py): clock64 stamps of CTA 0's first work item in the last
 * clslam_conv3x3_tma launch (synchronises the device); returns -1 in a normal build */
int clslam_debug_tma_stamps(unsigned long long* out32);

/* Input staging for conv1 on the TMA kernel: (x - 0.45) / 0.225 (networks/resnet_encoder.py:117), channel concat of
 * img_b (pose net) and 2x2 space-to-depth, written as planes (n, h/2, w/2, 16 | 32): channel (p*2 + q) * cp + c of
 * S2D pixel (y, x) is image pixel (2y + p, 2x + q), cp = 4 (one image) or 8 (two).  Images go to planes image
 * `image_offset` .. + n. */
int clslam_stage_images_s2d(const float* img_a, const float* img_b, const clslam_planes* out, int image_offset, int n, int h, int w,
                            void* stream);
/* fp32 NHWC (n, h, w, c) -> 2x2 space-to-depth planes (n, h/2, w/2, 4c), zero halo: the input of a 3x3 / stride-2 / pad-1
 * convolution (torchvision BasicBlock conv1 of layer2-4) run as clslam_conv3x3_tma(taps = 2) with
 * W'[co][r][s][(p, q, k)] = W[co][2r + p - 1][2s + q - 1][k]. */
int clslam_planes_s2d_from_f32(const float* x, const clslam_planes* out, int n, int h, int w, int c, void* stream);
/* clslam_conv_fwd_bf16w / clslam_maxpool3x3s2 that additionally emit their result as planes */
int clslam_conv_fwd_emit(const clslam_conv_desc* d, const float* x, const float* x2, const float* w, const void* w_hi,
                         const void* w_lo, const float* scale, const float* shift, const float* residual, float* y,
                         const clslam_planes* out, int out_up, int out_halo,
                         void* workspace, long long workspace_bytes, void* stream);
int clslam_maxpool3x3s2_emit(const float* x, float* y, int n, int hi, int wi, int c, const clslam_planes* out,
                             int out_halo, void* stream);
/* fp32 NHWC (n, h, w, c) -> planes (test / staging helper) */
int clslam_planes_from_f32(const float* x, const clslam_planes* out, int out_up, int out_halo, void* stream);

/* 3x3/stride-2/pad-1 max pooling, NHWC (networks/resnet_encoder.py:121). */
int clslam_maxpool3x3s2(const float* x, float* y, int n, int hi, int wi, int c, void* stream);

/* dispconv_s: reflection-padded 3x3 conv to ONE channel + sigmoid (networks/depth_decoder.py:69).
 * y: (n, 1, h, w). */
int clslam_dispconv_fwd(const float* x, const float* w, const float* bias, float* y, int n, int h, int w_,
                        int ci, void* stream);

/* ---- backward of the decoder convolutions (autograd of losses['loss'].backward(),
 * depth_pose_prediction.py:312; encoders are frozen, :807-827) --------------------------------
 * dgrad: gradient wrt the layer input, written as the pre-activation gradient of the layer
 *   that PRODUCED that input: dz_prev = act'(y_prev) * (dX [+ dispconv contribution]).
 *   For IN_UPCAT only the first c1 channels (the up-sampled decoder branch) are produced, already
 *   summed over the nearest-neighbour footprint, at (h1, w1); the skip branch feeds a frozen
 *   encoder and is not computed.
 *   disp_dz / disp_w (optional): pre-sigmoid gradient (n,1,hi,wi) and weights [1][3][3][ci] of
 *   the dispconv that also consumes y_prev; its input gradient is added before act'.
 * wgrad: dW [co][kh][kw][ci] and db [co] from the layer input (x, x2 as in forward) and dz. */
long long clslam_conv_dgrad_workspace_bytes(const clslam_conv_desc* d);
int clslam_conv_dgrad(const clslam_conv_desc* d, const float* dz, const float* w, const float* y_prev,
                      int act_prev, const float* disp_dz, const float* disp_w, float* dz_prev,
                      void* workspace, long long workspace_bytes, void* stream);
/* Input gradient of ANY convolution of the path for a slice [ci_off, ci_off + ci_n) of its input channels, at the layer's
 * input resolution: strided layers (the encoders' 3x3 / stride-2, 1x1 / stride-2 convolutions; zero padding), and the skip
 * part of an UPCAT layer (ci_off >= c1: the gradient that flows into the ENCODER feature, which the adaptation step -- frozen
 * encoders -- never needs).  dx (n, hi, wi, ci_n) = sum_taps W^T dz, times act'(y_prev) when y_prev is given.  Used by the
 * full training path (depth_pose_prediction.py:219-289, :797-800). */
long long clslam_conv_dgrad_ex_workspace_bytes(const clslam_conv_desc* d, int ci_n);
int clslam_conv_dgrad_ex(const clslam_conv_desc* d, const float* dz, const float* w, int ci_off, int ci_n, const float* y_prev,
                         int act_prev, float* dx, void* workspace, long long workspace_bytes, void* stream);
long long clslam_conv_wgrad_workspace_bytes(const clslam_conv_desc* d);
int clslam_conv_wgrad(const clslam_conv_desc* d, const float* x, const float* x2, const float* dz,
                      float* dw, float* db, void* workspace, long long workspace_bytes, void* stream);
/* dispconv backward: dW [1][3][3][ci], db [1] (wgrad) and, when the dispconv is the only consumer
 * of y_prev (scale 0), dz_prev = elu'(y_prev) * dX. */
long long clslam_dispconv_wgrad_workspace_bytes(int n, int h, int w_, int ci);
int clslam_dispconv_wgrad(const float* x, const float* dz, float* dw, float* db, int n, int h, int w_, int ci,
                          void* workspace, long long workspace_bytes, void* stream);
int clslam_dispconv_dgrad_act(const float* dz, const float* w, const float* y_prev, int act_prev,
                              float* dz_prev, int n, int h, int w_, int ci, void* stream);
/* the same, additionally writing dz_prev as bf16 hi/lo planes (the operand of the tcgen05 dgrad / wgrad of upconv_0_1):
 * saves the separate clslam_planes_from_f32 launch on the critical lane */
int clslam_dispconv_dgrad_act_emit(const float* dz, const float* w, const float* y_prev, int act_prev,
                                   float* dz_prev, const clslam_planes* dz_prev_planes, int n, int h, int w_, int ci,
                                   void* stream);

/* ---- pose head ----------------------------------------------------------------------------
 * mean over (W, H) * 0.01 -> (axis_angle, translation) of frame slot 0 (networks/pose_decoder.py:
 * 49-54; caller keeps [:,0], depth_pose_prediction.py:965), Rodrigues + translation
 * (utils.py:34-117; invert for the f=-1 pass), and P = (K @ cam_T_cam)[:3] of
 * networks/layers.py:94.  pose_out: (2*b, hw, 12) NHWC, pass 0 = frame -1, pass 1 = frame +1.
 * Outputs (pass-major): axis_angle (2,b,3), translation (2,b,3), cam_t_cam (2,b,16), proj (2,b,12). */
int clslam_pose_head_fwd(const float* pose_out, const float* k_mat, int b, int h, int w, int invert_mask,
                         float* axis_angle, float* translation, float* cam_t_cam, float* proj,
                         void* stream);
/* Backward of the above plus the velocity term (depth_pose_prediction.py:1125-1146):
 * dproj_partials (n_partials, 2, b, 12) are the per-block sums written by clslam_photo_bwd.
 * Writes dz of pose_2's output (2*b, hw, 12). */
int clslam_pose_head_bwd(const float* dproj_partials, int n_partials, const float* k_mat,
                         const float* axis_angle, const float* translation, const double* rel_dist,
                         const float* sample_w, float velocity_scale, int b, int h, int w, int invert_mask,
                         float* dz_pose_out, void* stream);

/* ---- photometric chain --------------------------------------------------------------------
 * identity: 0.85*mean_c SSIM + 0.15*mean_c L1 between the UNWARPED sources and the target
 *   (auto-masking candidates, depth_pose_prediction.py:1049-1053; SSIM networks/layers.py:123-137).
 *   src: 2 pointers' worth laid out (2, b, 3, h, w); out (b, 2, h, w).
 * fwd (per call: all `n_scales` scales): bilinear up (align_corners=False, :989-991) -> disp_to_depth
 *   (utils.py:120-142) -> BackprojectDepth/Project3D (networks/layers.py:74-104) -> grid_sample(bilinear,
 *   border, align_corners=True) (:1013-1017) -> SSIM+L1 (:1178-1192) -> + noise*1e-5 on the identity
 *   candidates, min over 4 (:1055-1061) -> per-sample sums.
 *   disp[s]: (b,1,h>>s,w>>s).  noise[s]: (b,2,h,w) standard normal or NULL.  Optional outputs
 *   depth[s] (b,1,h,w) and warped[s] (2,b,3,h,w) may be NULL.  argmin[s]: (b,h,w) uint8 stash
 *   for backward.  loss_partials: (n_scales, b, tiles) block sums (clslam_photo_tiles()).
 * bwd: gradient wrt disp at full resolution, ddisp_full[s] (b,h,w), and per-block sums of the
 *   gradient wrt proj, dproj_partials (n_scales*tiles, 2, b, 12). */
typedef struct clslam_photo_args {
  int b, h, w, n_scales;
  int depth_mode; float min_depth, max_depth;
  const float* target;        /* (b,3,h,w)  inputs['rgb',0,0] */
  const float* sources;       /* (2,b,3,h,w) inputs['rgb',-1,0], inputs['rgb',1,0] */
  const float* sources_rgba;  /* (2,b,h,w,4) the same images packed by clslam_pack_rgba (sampled by fwd / bwd) */
  const float* inv_k;         /* (b,4,4) inputs['inv_camera_matrix',0] */
  const float* proj;          /* (2,b,12) from clslam_pose_head_fwd */
  const float* identity;      /* (b,2,h,w) from clslam_photo_identity */
  const float* sample_w;      /* (b) loss weights w_b */
  const float* disp[4];
  const float* noise[4];
  float* depth[4];
  float* warped[4];
  unsigned char* argmin[4];
  float* loss_partials;
  float* ddisp_full[4];
  float* dproj_partials;
  /* mask_dynamic (depth_pose_prediction.py:1062-1075), both NULL otherwise.  dyn_mask: inputs['mask',0,0] as (b,h,w) floats,
   * >= 0.5 = potentially dynamic: such a pixel contributes nothing (forward: argmin code 4, which no gradient path matches).
   * reproj_w: (b) weights of the reprojection term in place of sample_w -- H*W / (static pixels of the batch) for every b, which
   * turns sum_b w_b * mean_hw(.) into the mean over the selected pixels (clslam_mask_prepare writes them). */
  const float* dyn_mask;
  const float* reproj_w;
} clslam_photo_args;
int clslam_photo_tiles(int h, int w);
/* (n,3,h,w) planar -> (n,h,w,4) packed (4th component 0): one 16-byte load per bilinear tap in the warp */
int clslam_pack_rgba(const float* planar, float* packed, int n_images, int h, int w, void* stream);
int clslam_photo_identity(const float* target, const float* sources, float* identity, int b, int h, int w,
                          void* stream);
int clslam_photo_fwd(const clslam_photo_args* a, void* stream);
int clslam_photo_bwd(const clslam_photo_args* a, void* stream);

/* ---- loss assembly ------------------------------------------------------------------------
 * depth_pose_prediction.py:1071-1113 incl. the smoothness term AS EXECUTED by the reference
 * (all b_global terms come from row 0 of batch element 0) and the velocity term.
 * losses out (24 floats): [0..3] reprojection_loss/scale_s, [4..7] smooth_loss, [8..11] reg_loss,
 * [12..15] depth_loss/scale_s, [16] depth_loss (== loss, aliased in the reference), [17] velocity_loss,
 * [18] loss, [19] nan flag.  smooth_grad out (n_scales, 2*(b_global+1)+1): direct d(loss)/d(disp)
 * at rows 0/1, cols 0..b_global of sample 0 followed by the uniform term from the mean
 * normalisation.  has_sample0 = 0 on ranks that do not hold global sample 0. */
typedef struct clslam_loss_args {
  int b, b_global, h, w, n_scales, n_tiles, has_sample0;
  float disparity_smoothness, velocity_scale;
  const float* loss_partials;   /* (n_scales, b, n_tiles) */
  const float* sample_w;        /* (b) local slice of the global weights */
  const float* sample_w_global; /* (b_global) */
  const float* disp[4];         /* (b,1,h>>s,w>>s) */
  const float* color[4];        /* (b,3,h>>s,w>>s) inputs['rgb',0,s] */
  const float* translation;     /* (2,b,3) */
  const double* rel_dist;       /* (2,b): relative_distance 0 and 1 */
  float* losses;                /* (24) */
  float* smooth_grad;           /* (n_scales, 2*(b_global+1)+1) */
  const float* reproj_w;        /* (b) or NULL: weights of the reprojection term in place of sample_w (mask_dynamic) */
} clslam_loss_args;
int clslam_loss_finalize(const clslam_loss_args* a, void* stream);

/* ---- mask_dynamic: pre-training with dynamic-object masks (depth_pose_prediction.py:1062-1068, :1080-1082, :1148-1176) -------
 * masks[s]: inputs['mask',0,s] as (b,1,h>>s,w>>s) floats, >= 0.5 = dynamic (the reference selects `1 - mask`).  Single GPU.
 * clslam_mask_prepare: reproj_w[bb] = H*W / (number of static pixels of the whole batch at scale 0), and for every scale the flat
 *   positions of the first b SELECTED x-gradients (grid (b,h_s,w_s-1), selected by static[..., :-1]) and y-gradients (grid
 *   (b,h_s-1,w_s), selected by static[..., :-1, :]) in raster order: pos (n_scales, 2, b) ints, -1 where fewer exist.  These are
 *   the elements the reference's per-sample loop ends up reading: "sample i" gets the i-th selected gradient of the batch.
 * clslam_smooth_masked: run AFTER clslam_loss_finalize (called with has_sample0 = 0, so that its own smoothness term is zero):
 *   evaluates the b terms per scale, patches losses (smooth / reg / depth_loss per scale, depth_loss, loss) and writes the
 *   gradient with respect to the (sigmoid) disparities: direct entries sm_idx / sm_val (n_scales, 4*b) -- pixel index into
 *   (b,h_s,w_s), -1 = unused -- and the per-sample uniform term of the mean normalisation sm_uniform (n_scales, b).
 * clslam_smooth_masked_apply: dz_disp[q] += (uniform[sample of q] + sum of the direct entries at q) * disp[q] * (1 - disp[q]),
 *   after clslam_disp_grad of that scale. */
int clslam_mask_prepare(const float* const* masks, int b, int h, int w, int n_scales, float* reproj_w, int* pos, void* stream);
int clslam_smooth_masked(const float* const* disp, const float* const* color, const int* pos, const float* sample_w, int b, int h,
                         int w, int n_scales, float disparity_smoothness, float* losses, int* sm_idx, float* sm_val,
                         float* sm_uniform, void* stream);
int clslam_smooth_masked_apply(const float* disp, const int* sm_idx, const float* sm_val, const float* sm_uniform, int b, int h,
                               int w, int scale, float* dz_disp, void* stream);
/* d(loss)/d(disp_s): transpose of the bilinear up-sampling applied to ddisp_full[s], plus the
 * smoothness terms, times sigmoid'; writes dz of dispconv_s (b,1,h>>s,w>>s). */
int clslam_disp_grad(const float* ddisp_full, const float* disp, const float* smooth_grad, int b_global,
                     int has_sample0, int b, int h, int w, int scale, float* dz_disp, void* stream);

/* depth = disp_to_depth(disp) (depth_pose_prediction/utils.py:120-142) elementwise over n values; depth_mode as in
 * clslam_photo_args (CLSLAM_DEPTH_*).  Used by the depth-only inference path (predict_from_image, :538-554). */
int clslam_disp_to_depth(const float* disp, float* depth, long long n, int depth_mode, float min_depth, float max_depth,
                         void* stream);

/* ---- replay-sample image path (slam/replay_buffer.py:263-291, datasets/utils.py:236-259) -------
 * One pass of Pillow's 8-bit ImagingResample over n interleaved uint8 images (n, h_in, w_in, c): horizontal
 * (vertical = 0: -> (n, h_in, out_size, c)) or vertical (-> (n, out_size, w_in, c)).  bounds (out_size x 2: first source
 * index, tap count) and kk (out_size x ksize fixed-point coefficients, 22 fractional bits) are device arrays built as
 * Pillow's precompute_coeffs + normalize_coeffs_8bpc build them (clslam_b200/replay_path.py::pil_coeffs).  Bit-exact. */
int clslam_resample_u8(const void* src, void* dst, int n, int h_in, int w_in, int c, int out_size, int vertical,
                       const int* bounds, const int* kk, int ksize, void* stream);
/* torchvision ToTensor: (n, h, w, 3) uint8 -> (n, 3, h, w) float32 / 255 */
int clslam_u8_to_f32_chw(const void* hwc, float* chw, int n, int h, int w, void* stream);
/* One colour-jitter step of torchvision.transforms.functional on n float images (n, 3, h, w) in [0, 1]:
 * op 0 adjust_brightness, 1 adjust_contrast (per-image grey mean), 2 adjust_saturation, 3 adjust_hue; `factor` is the
 * sampled factor (a Python float in the reference).  src == dst is allowed.  Workspace: clslam_color_jitter_workspace_bytes(n). */
long long clslam_color_jitter_workspace_bytes(int n);
int clslam_color_jitter_op(const float* src, float* dst, int n, int h, int w, int op, double factor, void* workspace,
                           long long workspace_bytes, void* stream);

/* ---- full training path: BatchNorm with batch statistics, max-pool backward, accumulation ---------
 * nn.BatchNorm2d (eps 1e-5, momentum 0.1) over fp32 NHWC (n, hw, c).  `groups` > 1: the n images are that many consecutive
 * module calls of n / groups images each (the pose encoder's two passes), each with its own batch statistics; the running
 * statistics are updated once per group, in order.  training = 0: running statistics (eval mode), nothing is updated.
 * y = act((x - mean) * invstd * gamma + beta (+ residual)); save_mean / save_invstd: (groups, c), kept for the backward. */
long long clslam_bn_workspace_bytes(int groups, int c);
int clslam_bn_fwd(const float* x, const float* gamma, const float* beta, float* running_mean, float* running_var, int training,
                  float momentum, float eps, const float* residual, int relu, int n, int hw, int c, int groups, float* y,
                  float* save_mean, float* save_invstd, void* workspace, long long workspace_bytes, void* stream);
/* backward: g = dy * relu'(y) (relu != 0), dgamma = sum g * xhat, dbeta = sum g, dx = gamma * invstd * (g - mean(g) -
 * xhat * mean(g * xhat)) with batch statistics, gamma * invstd * g in eval mode.  g_out (optional) receives g: the gradient of
 * the identity branch of a BasicBlock. */
int clslam_bn_bwd(const float* x, const float* dy, const float* y, int relu, const float* gamma, const float* save_mean,
                  const float* save_invstd, int training, int n, int hw, int c, int groups, float* dx, float* dgamma,
                  float* dbeta, float* g_out, void* workspace, long long workspace_bytes, void* stream);
/* backward of MaxPool2d(3, 2, 1) (resnet_encoder.py:121), NHWC; ties go to the first maximum like ATen */
int clslam_maxpool3x3s2_bwd(const float* x, const float* dy, float* dx, int n, int hi, int wi, int c, void* stream);
/* out = a + b + c (b, c optional), n floats: gradient accumulation where branches of the graph join */
int clslam_add3(const float* a, const float* b, const float* c, float* out, long long n, void* stream);

/* ---- optimiser ----------------------------------------------------------------------------
 * torch.optim.Adam (lr, betas, eps; no weight decay / amsgrad) as built at
 * depth_pose_prediction.py:203 and stepped at :313, over one flat fp32 buffer.  `step` is the
 * 1-based step number AFTER this update; if step_dev != NULL it is read from (and incremented on)
 * the device instead, which makes the call CUDA-graph replayable.
 * Bucketed use (data-parallel step: one call per gradient bucket as soon as its all-reduce has finished): pass
 * step = CLSLAM_ADAM_NO_TICK with step_dev, so that every bucket of an iteration reads the same counter, and call
 * clslam_adam_tick once after the last bucket. */
#define CLSLAM_ADAM_NO_TICK (-1)
int clslam_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                     double lr, double beta1, double beta2, double eps, int step, int* step_dev,
                     void* stream);
int clslam_adam_tick(int* step_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CLSLAM_B200_H */
