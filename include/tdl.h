/*
 * tdl.h -- C ABI of libtdl.so: the fused multi-scale view-synthesis loss of
 * TripleDNet / FeatDepth / monodepth2, hand-written CUDA for sm_100a (B200).
 *
 * The reference (pure Python, no FFI of its own -- SURVEY.md section 8b) exposes this
 * path only as methods of its net classes.  Each entry point below names the
 * reference call sites it replaces (paths relative to the reference repository):
 *
 *   tdl_photo_fwd / tdl_photo_bwd
 *       generate_images_pred            mono/model/mono_fm/net.py:157-170
 *                                       mono/model/mono_baseline/net.py:123-135
 *                                       mono/model/mono_fm_joint/net.py:181-194
 *         = F.interpolate + disp_to_depth (net.py:135-140) + Backproject.forward
 *           (mono/model/mono_fm/layers.py:57-61) + Project.forward (layers.py:73-82)
 *           + F.grid_sample(border)
 *       compute_reprojection_loss       mono/model/mono_fm/net.py:63-67  (SSIM layers.py:97-107,
 *                                       robust_l1 net.py:55-57)
 *       automask + min over frames      mono/model/mono_fm/net.py:90-106
 *       disp mean-normalisation         mono/model/mono_fm/net.py:123-125
 *       get_smooth_loss / gradient      mono/model/mono_fm/net.py:255-283
 *       ... and the autograd backward of all of the above.
 *   tdl_feat_fwd / tdl_feat_bwd
 *       generate_features_pred          mono/model/mono_fm/net.py:172-199
 *                                       mono/model/mono_fm_joint/net.py:196-223
 *       compute_perceptional_loss + min mono/model/mono_fm/net.py:59-61,111-118
 *                                       mono/model/mono_fm_joint_inpaint/net.py:58-70
 *   tdl_edge_smooth_fwd / tdl_edge_smooth_bwd, tdl_edge_smooth_multi_fwd / _bwd (all encoder levels in one call)
 *       get_feature_regularization_loss mono/model/mono_fm_joint/net.py:309-330 (called per level at :77-80)
 *   tdl_recon_fwd / tdl_recon_bwd
 *       masked img_reconstruct_loss     mono/model/mono_fm_joint_inpaint/net.py:80-91
 *   tdl_proj_fwd / tdl_proj_bwd
 *       K @ T, K/2 and its inverse      mono/model/mono_fm/layers.py:58,74; mono/model/mono_fm/net.py:185-191
 *   tdl_pose_fwd / tdl_pose_bwd
 *       transformation_from_parameters  mono/model/mono_fm/net.py:201-212 (get_translation_matrix :214-222,
 *                                       rot_from_axisangle :224-253), called by predict_poses (net.py:142-155)
 *   tdl_input_fwd  (the per-item host work of the data loader, on uint8 frames already resident on the device)
 *       to_tensor + color_aug           mono/datasets/mono_dataset.py:84-103 (ColorJitter ranges :62-73, sampled :182-187)
 *       erase masks                     mono/datasets/kitti_dataset.py:167-182 (cfg erase_count / erase_shape)
 *
 * Conventions
 *   - every tensor is fp32, NCHW, contiguous, resident in device memory; only
 *     min_index is int64 (the reference's torch.min index dtype).
 *   - calls are asynchronous on `stream`, never allocate, never synchronise the host and
 *     keep no state between calls: scratch lives in the caller-provided workspace, whose
 *     contents must be preserved from a *_fwd call to the matching *_bwd call.
 *   - return value: 0 = TDL_OK, < 0 = argument error (nothing was launched),
 *     > 0 = a cudaError_t raised by a launch.  tdl_strerror() explains either.
 *   - there is NO CPU implementation behind these symbols.
 */
#ifndef TDL_H_
#define TDL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TDL_ABI_VERSION 4
#define TDL_MAX_SRC 4      /* source frames per target (frame_ids[1:])       */
#define TDL_MAX_SCALES 4   /* disparity scales (opt.scales)                  */
#define TDL_MAX_LEVELS 5   /* encoder levels of get_feature_regularization_loss */

#define TDL_OK 0
#define TDL_ERR_NULL (-1)      /* a required pointer is NULL                  */
#define TDL_ERR_SHAPE (-2)     /* unsupported sizes (see each struct)         */
#define TDL_ERR_WORKSPACE (-3) /* workspace_bytes smaller than *_ws_bytes()   */
#define TDL_ERR_COUNT (-4)     /* S / nscales / C out of range                */
#define TDL_ERR_NODEVICE (-5)  /* no CUDA device / not sm_100                 */
#define TDL_ERR_OPTION (-6)    /* tdl_set_option / tdl_get_option: unknown name */

/* storage of the feature maps handed to tdl_feat_* (arithmetic is fp32 in registers either way)            */
#define TDL_LAYOUT_NCHW 0      /* (B,C,h,w) contiguous: what the reference's extractor returns               */
#define TDL_LAYOUT_NHWC 1      /* (B,h,w,C) contiguous: torch channels_last memory of a (B,C,h,w) tensor --
                                  one bilinear tap is one contiguous C-vector (256 B at C=64 fp32)             */
#define TDL_DTYPE_F32 0
#define TDL_DTYPE_BF16 1       /* opt-in storage mode (north_star "bf16/fp32 loads"); never the parity config */

typedef void* tdl_stream_t;    /* a cudaStream_t */

/* ------------------------------------------------------------------ photometric + smoothness */
typedef struct tdl_photo_args {
    int32_t B, H, W;            /* batch, full-resolution height / width       */
    int32_t S;                  /* number of source frames, 1..TDL_MAX_SRC      */
    int32_t nscales;            /* 1..TDL_MAX_SCALES                            */
    int32_t disp_h[TDL_MAX_SCALES], disp_w[TDL_MAX_SCALES];
                                /* disp[s] is (B,1,disp_h[s],disp_w[s]); H/disp_h == W/disp_w
                                   must be a power of two in {1,2,4,8,16,32}            */
    int32_t automask;           /* != 0: identity channels precede the warped ones (net.py:90-95) */
    int32_t disp_norm;          /* != 0: disp / (mean_hw(disp) + 1e-7) before smoothness           */
    int32_t align_corners;      /* F.grid_sample convention; 0 = torch >= 1.3 default              */
    int32_t reserved0;
    double min_depth, max_depth;
    float photo_coef[TDL_MAX_SCALES];   /* losses[s]          = photo_coef[s]  * mean_{b,y,x} min_c rho_c      */
    float smooth_coef[TDL_MAX_SCALES];  /* losses[nscales+s]  = smooth_coef[s] * (sum of the six means)       */
    float smooth_alpha;                 /* exp(-alpha * mean_c |d img|); the reference uses 0.5              */
    float reserved1;
    uint64_t noise_seed;        /* Philox seed used when automask and noise[s][f] == NULL               */

    const float* target;                     /* (B,3,H,W)  inputs[("color",0,0)]                       */
    const float* src[TDL_MAX_SRC];           /* (B,3,H,W)  inputs[("color",f,0)]                       */
    const float* disp[TDL_MAX_SCALES];       /* outputs[("disp",0,s)]                                  */
    const float* P;                          /* (B,S,3,4)  (K @ T_f)[:, :3, :]   (layers.py:74)        */
    const float* invK;                       /* (B,3,3)    inv_K[:, :3, :3]      (layers.py:58)        */
    const float* noise[TDL_MAX_SCALES][TDL_MAX_SRC];
                                             /* (B,1,H,W) N(0,1) draws, scaled by 1e-5 in-kernel; NULL => Philox */
    float* warped[TDL_MAX_SCALES][TDL_MAX_SRC];  /* out, optional: outputs[("color",f,s)] (B,3,H,W)      */
    int64_t* min_index[TDL_MAX_SCALES];          /* out, optional: outputs[("min_index",s)] (B,H,W)      */
    void* workspace;
    uint64_t workspace_bytes;
    float* losses;              /* out [2*nscales]: photometric per scale, then smoothness per scale    */

    /* backward only */
    const float* dlosses;       /* [2*nscales] upstream gradient of `losses` (device memory)           */
    float* d_disp[TDL_MAX_SCALES];  /* out (B,1,disp_h[s],disp_w[s]), overwritten                       */
    float* dP;                  /* out (B,S,3,4), overwritten                                          */
} tdl_photo_args;

/* ------------------------------------------------------------------ feature-metric */
typedef struct tdl_feat_args {
    int32_t B, C, h, w;         /* feature maps are (B,C,h,w); the reference uses h=H/2, w=W/2, C=64    */
    int32_t S;
    int32_t disp_h, disp_w;     /* disp is (B,1,disp_h,disp_w), resized bilinearly to (h,w)            */
    int32_t align_corners;
    int32_t layout;             /* TDL_LAYOUT_*: of tgt, src, warped, d_tgt, d_src (all the same)       */
    int32_t dtype;              /* TDL_DTYPE_*:  of tgt, src, warped, d_tgt, d_src (all the same);
                                   NHWC requires C % 8 == 0; BF16 requires NHWC                          */
    double min_depth, max_depth;
    float coef;                 /* loss = coef * mean_{b,y,x} min_f mean_c robust_l1                    */
    float reserved0;

    const void* tgt;                         /* (B,C,h,w)  extractor(color_0)[0]      (float or bf16)   */
    const void* src[TDL_MAX_SRC];            /* (B,C,h,w)  extractor(color_f)[0]                       */
    const float* disp;
    const float* P;                          /* (B,S,3,4) built from the half-resolution K            */
    const float* invK;                       /* (B,3,3)   pinv(K_half)[:, :3, :3]                      */
    void* warped[TDL_MAX_SRC];               /* out, optional: outputs[("feature",f,0)]               */
    int64_t* min_index;                      /* out, optional (B,h,w)                                  */
    void* workspace;
    uint64_t workspace_bytes;
    float* loss;                /* out [1]                                                              */

    /* backward only */
    const float* dloss;         /* [1] */
    void* d_tgt;                /* out, optional (B,C,h,w), overwritten                                 */
    void* d_src[TDL_MAX_SRC];   /* out, optional (all or none), overwritten (zero where not selected)   */
    float* d_disp;              /* out (B,1,disp_h,disp_w), overwritten                                 */
    float* dP;                  /* out (B,S,3,4), overwritten                                           */
    void* bwd_scratch;          /* optional, >= tdl_feat_bwd_scratch_bytes(): lets the d_src scatter of
                                   grid_sample's backward run as a bucketed GATHER (no global atomics);
                                   contents need not be preserved; NULL / too small / C % 4 != 0 =>
                                   the atomic scatter kernel is used instead (same results)             */
    uint64_t bwd_scratch_bytes;
} tdl_feat_args;

/* ------------------------------------------------------------------ edge-aware smoothness on C-channel maps */
typedef struct tdl_edge_args {
    int32_t B, C, h, w;         /* feature (B,C,h,w)                                                    */
    int32_t H, W;               /* image (B,3,H,W); H/h == W/w must be a power of two <= 32             */
    float alpha;                /* exp(-alpha * mean_c |d img|): 1.0 in get_feature_regularization_loss */
    float first_coef;           /* loss = first_coef * (dx+dy terms) + second_coef * (dxx+dxy+dyx+dyy)  */
    float second_coef;          /*        (the reference: -dis/2^i/5 and cvt/2^i/5)                     */
    float reserved0;
    const float* feature;
    const float* image;
    void* workspace;
    uint64_t workspace_bytes;
    float* loss;                /* out [1] */
    const float* dloss;         /* [1], backward only */
    float* d_feature;           /* out (B,C,h,w), overwritten, backward only */
} tdl_edge_args;

/* All levels of get_feature_regularization_loss in ONE call (mono/model/mono_fm_joint/net.py:77-80 loops over the five
 * encoder levels): one launch builds every level's area-downsampled image, one evaluates all levels, one forms the
 * scalars -- 3 launches instead of 5 x (memset + 3); the backward is one launch instead of five.  Every level keeps its
 * own tdl_edge_args (feature, workspace, loss, coefficients); B, image, H and W must be the same in all of them. */
typedef struct tdl_edge_multi_args {
    int32_t nlevels;            /* 1..TDL_MAX_LEVELS */
    int32_t reserved0;
    tdl_edge_args level[TDL_MAX_LEVELS];
} tdl_edge_multi_args;

/* one row of tdl_profile_end(): device time attributed to one kernel (CUDA events on the launch stream) */
typedef struct tdl_kernel_time {
    char name[32];
    int32_t launches;
    double total_ms;
} tdl_kernel_time;

/* ------------------------------------------------------------------ masked image reconstruction (TripleD) */
typedef struct tdl_recon_args {
    int32_t B, h, w;            /* pred / target / mask are (B,3,h,w)                                    */
    float coef;                 /* loss = coef * sum(rho * (1 - mask)) / sum(1 - mask),
                                   rho = 0.85 mean_c SSIM(pred,target) + 0.15 mean_c robust_l1 (B,1,h,w) */
    const float* pred;          /* outputs[("res_img",0,s)]                                              */
    const float* target;        /* target resized to (h,w) (F.interpolate bilinear, done by the caller)  */
    const float* mask;          /* mask resized to (h,w); NULL => plain mean                             */
    void* workspace;
    uint64_t workspace_bytes;
    float* loss;                /* out [1] */
    const float* dloss;         /* [1], backward only */
    float* d_pred;              /* out (B,3,h,w), overwritten, backward only */
} tdl_recon_args;

/* ------------------------------------------------------------------ pose prologue (axis-angle, translation -> cam_T_cam) */
typedef struct tdl_pose_args {
    int32_t B;
    int32_t invert;             /* != 0: the inverted transform the reference builds for frame ids < 0 (net.py:204-209) */
    const float* axisangle;     /* (B,3)   PoseDecoder output axisangle[:, 0]                              */
    const float* translation;   /* (B,3)   PoseDecoder output translation[:, 0]                            */
    float* T;                   /* out (B,4,4) outputs[("cam_T_cam", 0, f)]                                 */
    /* backward only */
    const float* dT;            /* (B,4,4) upstream gradient of T                                           */
    float* d_axisangle;         /* out (B,3), overwritten                                                   */
    float* d_translation;       /* out (B,3), overwritten                                                   */
} tdl_pose_args;

/* ------------------------------------------------------------------ projection prologue (K, inv_K, cam_T_cam -> P, inv_K 3x3) */
/* Everything the loss kernels need from the camera matrices, in one launch each way:
 *   P_full[b][f] = (K @ T_f)[:3, :]                      Project.forward                 mono/model/mono_fm/layers.py:74
 *   P_half[b][f] = (K' @ T_f)[:3, :], K' = K with rows 0, 1 halved   generate_features_pred   mono/model/mono_fm/net.py:185-187
 *   invK3        = inv_K[:, :3, :3]                      Backproject.forward             mono/model/mono_fm/layers.py:58
 *   invKh3       = pinv(K')[:, :3, :3] = inv_K[:, :3, :3] with columns 0, 1 doubled      mono/model/mono_fm/net.py:188-191
 * (the reference spends a matmul, a slice and -- per frame, per scale -- a clone, two divisions and a per-sample SVD on
 * these 12 + 9 numbers per image; eager PyTorch needs ~30 small kernels for them forward + backward).  Full fp32
 * whatever torch's TF32 matmul switch says.  Backward: dT_f = K[:3]^T (dP_full_f + diag(1/2, 1/2, 1) dP_half_f). */
typedef struct tdl_proj_args {
    int32_t B, S;
    const float* K;                      /* (B,4,4) inputs["K"]                                           */
    const float* inv_K;                  /* (B,4,4) inputs["inv_K"]                                       */
    const float* T[TDL_MAX_SRC];         /* (B,4,4) cam_T_cam (or stereo_T) of every source frame         */
    float* P_full;                       /* out (B,S,3,4)                                                 */
    float* P_half;                       /* out (B,S,3,4)                                                 */
    float* invK3;                        /* out (B,3,3)                                                   */
    float* invKh3;                       /* out (B,3,3)                                                   */
    /* backward only */
    const float* dP_full;                /* (B,S,3,4) or NULL (treated as zero)                           */
    const float* dP_half;                /* (B,S,3,4) or NULL                                             */
    float* dT[TDL_MAX_SRC];              /* out (B,4,4) per frame, overwritten; NULL entries are skipped  */
} tdl_proj_args;

/* ------------------------------------------------------------------ input pipeline (uint8 frames -> color, color_aug, mask) */
/* Byte-exact torchvision ColorJitter on PIL images (Pillow's Image.blend / convert("L") / convert("HSV") arithmetic),
 * transforms.ToTensor and the inpainting erase mask, for a whole batch in two launches.  The caller samples the
 * augmentation parameters the way torchvision's ColorJitter.get_params does (host side, a few numbers per image). */
typedef struct tdl_input_args {
    int32_t B, H, W;
    int32_t nframes;            /* frames per item (target + sources), 1..TDL_MAX_SRC+1                       */
    int32_t erase_count;        /* boxes per image (cfg.erase_count), 0..64; 0 = no mask                        */
    int32_t erase_h, erase_w;   /* cfg.erase_shape                                                             */
    int32_t reserved0;
    const uint8_t* frames[TDL_MAX_SRC + 1];  /* (B,H,W,3) uint8, the resized PIL frames' memory layout         */
    const float* jitter;        /* (B,nframes,4): brightness, contrast, saturation factors and the BYTE added to the hue
                                   channel, float(uint8(int32(hue_factor * 255))); NULL => color_aug = color    */
    const int32_t* order;       /* (B,nframes,4): ColorJitter's fn_idx permutation (0 brightness, 1 contrast,
                                   2 saturation, 3 hue); entries outside 0..3 skip the step                     */
    const uint8_t* do_aug;      /* (B): do_color_aug of the item; required when jitter != NULL                  */
    const uint8_t* do_flip;     /* (B) or NULL: horizontal flip of the item's frames                            */
    const int32_t* holes;       /* (B,erase_count,2): (row, col) of each box's top-left corner                  */
    float* color[TDL_MAX_SRC + 1];      /* out, optional (B,3,H,W): inputs[("color", f, 0)]                    */
    float* color_aug[TDL_MAX_SRC + 1];  /* out, optional (B,3,H,W): inputs[("color_aug", f, 0)]                */
    float* mask;                /* out, optional (B,3,H,W) of 0 / 1: inputs[("mask", 0, 0)]                     */
    void* workspace;            /* >= tdl_input_ws_bytes(B, nframes) when jitter != NULL                        */
    uint64_t workspace_bytes;
} tdl_input_args;

int tdl_abi_version(void);
const char* tdl_strerror(int code);
/* Process-wide switches for tests and kernel experiments (the defaults are the product path).  They are
 * initialised once from the environment (TDL_NO_TMA, TDL_FUSED_FWD, TDL_PHOTO_SPARSE_MAX, TDL_FEAT_ATOMIC,
 * TDL_FEAT_CHUNK, TDL_PHOTO_V1, TDL_PHOTO_LIST_MAX, TDL_FEAT_NO_BULK) and afterwards only change through this call -- the entry points
 * never call getenv().
 * Names: "no_tma", "fused_fwd", "photo_sparse_max", "feat_atomic", "feat_chunk", "photo_v1" (round-1 scoring kernel),
 * "feat_no_bulk" (channel-last feature forward with per-lane cp.async instead of TMA bulk copies),
 * "photo_list_max" (backward: (image, scale) pairs with at most this many selected windows run from the work list the
 * scoring kernel emits, default and maximum 4096; -1 = always the tile kernel). */
int tdl_set_option(const char* name, int value);
int tdl_get_option(const char* name, int* value);
/* number of CUDA kernels (not memsets) one call launches -- used by bench.py's gpu_launches */
int tdl_launch_count(const char* entry_point);

/* Per-kernel timing for bench.py (not thread-safe; do not use while a CUDA graph is being captured):
 * between begin and end every kernel / memset the library launches is bracketed by a cudaEvent pair. */
int tdl_profile_begin(void);
int tdl_profile_end(tdl_kernel_time* out, int max_entries);

uint64_t tdl_photo_ws_bytes(int32_t B, int32_t H, int32_t W, int32_t S, int32_t nscales,
                            const int32_t* disp_h, const int32_t* disp_w);
int tdl_photo_fwd(const tdl_photo_args* args, tdl_stream_t stream);
int tdl_photo_bwd(const tdl_photo_args* args, tdl_stream_t stream);

uint64_t tdl_feat_ws_bytes(int32_t B, int32_t C, int32_t h, int32_t w, int32_t S);
uint64_t tdl_feat_bwd_scratch_bytes(int32_t B, int32_t C, int32_t h, int32_t w, int32_t S);
int tdl_feat_fwd(const tdl_feat_args* args, tdl_stream_t stream);
int tdl_feat_bwd(const tdl_feat_args* args, tdl_stream_t stream);

uint64_t tdl_edge_ws_bytes(int32_t B, int32_t C, int32_t h, int32_t w);
int tdl_edge_smooth_fwd(const tdl_edge_args* args, tdl_stream_t stream);
int tdl_edge_smooth_bwd(const tdl_edge_args* args, tdl_stream_t stream);
int tdl_edge_smooth_multi_fwd(const tdl_edge_multi_args* args, tdl_stream_t stream);
int tdl_edge_smooth_multi_bwd(const tdl_edge_multi_args* args, tdl_stream_t stream);

uint64_t tdl_recon_ws_bytes(void);
int tdl_recon_fwd(const tdl_recon_args* args, tdl_stream_t stream);
int tdl_recon_bwd(const tdl_recon_args* args, tdl_stream_t stream);

int tdl_pose_fwd(const tdl_pose_args* args, tdl_stream_t stream);
int tdl_pose_bwd(const tdl_pose_args* args, tdl_stream_t stream);

int tdl_proj_fwd(const tdl_proj_args* args, tdl_stream_t stream);
int tdl_proj_bwd(const tdl_proj_args* args, tdl_stream_t stream);

uint64_t tdl_input_ws_bytes(int32_t B, int32_t nframes);
int tdl_input_fwd(const tdl_input_args* args, tdl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TDL_H_ */
