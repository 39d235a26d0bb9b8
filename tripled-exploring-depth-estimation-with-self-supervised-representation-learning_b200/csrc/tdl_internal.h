// Internal launch descriptors shared by tdl_api.cu and the kernel translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "tdl.h"

namespace tdl {

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) attribute: opt in once per device the
// process launches on, not once per process (a second GPU would otherwise fail with cudaErrorInvalidValue).
struct SmemOptIn {
    std::atomic<unsigned long long> done{0};
    template <typename Kernel>
    cudaError_t operator()(Kernel kernel, size_t bytes) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        const unsigned long long bit = 1ull << (dev & 63);
        if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
        return e;
    }
};

struct DepthParamsH {
    float min_disp, range;
};

// ---- photometric + smoothness, device-side view of tdl_photo_args -------------
constexpr int kListCap = 4096;           // work-list entries per (image, scale); 8192 measured worse: the list kernel's time is set
                                         // by its LONGEST list (74 CTAs per pair sweep it serially): 44 -> 85 us on the bench's smooth batch
constexpr int kListMagic = 0x5C0DE2;     // written by photo_score2_kernel after the lists of a forward are complete

struct PhotoDev {
    int B, H, W, S, nscales;
    int dh[TDL_MAX_SCALES], dw[TDL_MAX_SCALES], fac[TDL_MAX_SCALES];
    float sy[TDL_MAX_SCALES], sx[TDL_MAX_SCALES];     // dh/H, dw/W (F.interpolate scale)
    int automask, align_corners;
    int use_tma;                 // stage image tiles with TMA box copies when the tensors allow it
    int split_fwd;               // forward = warp kernel + TMA-staged scoring kernel (needs materialised warps)
    int v1;                      // use the round-1 scoring kernel (A/B comparisons and tests)
    int sparse_max;              // backward: tiles with at most this many selected windows (of 1156 incl. halo) scatter
                                 // their adjoint instead of running the dense box-sum gather (auto-masked regions)
    int list_max;                // backward: (image, scale) pairs with at most this many selected windows run from the work list
                                 // the scoring kernel emitted (static scenes: auto-masking leaves ~1 % of the windows); -1: off
    float min_disp, range;
    uint64_t seed;
    const float* target;
    const float* src[TDL_MAX_SRC];
    const float* disp[TDL_MAX_SCALES];
    const float* P;
    const float* invK;
    const float* noise[TDL_MAX_SCALES][TDL_MAX_SRC];
    float* warped[TDL_MAX_SCALES][TDL_MAX_SRC];
    long long* min_index[TDL_MAX_SCALES];
    // workspace
    double* acc;                 // [nscales][B][4]: photo sum, disp sum, smooth first, smooth second
    int acc_n;                   // doubles cleared at the start of every forward: the sums above + the work-list header
    int* lcnt;                   // [nscales*B] selected windows per (scale, image) | [nscales*B] = kListMagic once the lists are valid
    uint32_t* wlist;             // [nscales*B][kListCap] selected windows: pixel | frame << 28
    unsigned char* argmin;       // [nscales][B][H][W]
    float* J[TDL_MAX_SCALES];    // area-downsampled target (B,3,dh,dw)
    float* Wt[TDL_MAX_SCALES];   // smoothness edge weights (B,6,dh,dw)
    // backward
    const float* dlosses;
    float photo_coef[TDL_MAX_SCALES];
    float* d_disp[TDL_MAX_SCALES];
    float* dP;
};

// ---- edge-aware smoothness over a set of levels (one launch) ---------------------
struct SmoothLevel {
    int C, h, w;
    const float* x;        // (B,C,h,w)
    const float* J;        // (B,3,h,w) area-downsampled image
    float* Wt;             // (B,6,h,w) edge weights exp(-alpha*mean_c|stencil(J)|): written by the forward, re-read by the backward
    double* acc;           // per image b: acc[b*stride + 2] += first, acc[b*stride + 3] += second;
                           //              acc[b*stride + 1] holds sum(x) when norm
    int acc_stride;
    int norm;              // divide x by (mean_hw + 1e-7) first (C must be 1)
    float alpha;
    float first_coef, second_coef;
    // backward
    const float* dloss;    // scalar upstream
    float* dx;             // (B,C,h,w), overwritten
};

struct SmoothDev {
    int B, nlevels;
    float* zero_ptr;       // optional: zero_n floats cleared by the first CTA (folds a tiny memset of the NEXT kernel's
    int zero_n;            // accumulator into this launch; stream order makes it visible)
    SmoothLevel lv[TDL_MAX_LEVELS];      // disparity scales (<= TDL_MAX_SCALES) or encoder levels
};

// ---- feature-metric ----------------------------------------------------------------
struct FeatDev {
    int B, C, h, w, S;
    int bulk;                    // NHWC forward: rows travel by TMA bulk copies (option feat_no_bulk = 0)
    int layout, dtype;           // TDL_LAYOUT_* / TDL_DTYPE_* of tgt, src, warped, d_tgt, d_src (bf16 pointers are carried as float*)
    int Bnorm;                   // batch size of the mean (== B except for the batch chunks of the bucketed backward)
    int dh, dw;
    float sy, sx;
    int align_corners;
    float min_disp, range;
    float coef;
    const float* tgt;
    const float* src[TDL_MAX_SRC];
    const float* disp;
    const float* P;
    const float* invK;
    float* warped[TDL_MAX_SRC];
    long long* min_index;
    double* acc;                 // [B]
    unsigned char* argmin;       // [B][h][w]
    float* loss;
    const float* dloss;
    float* d_tgt;
    float* d_src[TDL_MAX_SRC];
    float* d_disp;
    float* dP;
    // bucketed d_src gather (backward scratch; all null when the atomic scatter kernel is to be used)
    float* dP_acc;               // [B][S][12]   dP accumulator inside the zeroed scratch header; copied to dP by the gather kernel
    float* G;                    // [B][h*w][C]  d loss / d warped value, channel-last
    int* bk_cnt;                 // [S][B][h*w]  taps registered per source pixel
    int2* bk_ent;                // [S][B][h*w][kFeatBucketCap]  (target pixel, weight bits)
    int* ov_cnt;                 // [B]          length of each image's overflow list
    int4* ov_ent;                // [B][4*h*w]   (frame, source pixel, target pixel, weight bits)
};
constexpr int kFeatBucketCap = 8;

// ---- masked image-reconstruction loss (TripleD family) ----------------------------------------------------
struct ReconArgsDev {
    int B, h, w;
    float coef;
    const float* pred;
    const float* tgt;
    const float* mask;
    double* acc;
    float* loss;
    const float* dloss;
    float* d_pred;
};
cudaError_t launch_recon_fwd(const ReconArgsDev& a, cudaStream_t st);
cudaError_t launch_recon_bwd(const ReconArgsDev& a, cudaStream_t st);

// launchers (each returns the cudaError_t of its launch)
cudaError_t launch_photo_fwd(const PhotoDev& p, cudaStream_t st);
bool photo_fwd_can_split(const PhotoDev& p);
cudaError_t launch_photo_warp(const PhotoDev& p, cudaStream_t st);
cudaError_t launch_photo_score(const PhotoDev& p, cudaStream_t st);      // round-1 strip kernel (option photo_v1)
cudaError_t launch_photo_score2(const PhotoDev& p, cudaStream_t st);     // register micro-tile kernel (tdl_photo2.cu)
cudaError_t launch_photo_bwd(const PhotoDev& p, cudaStream_t st);
cudaError_t launch_photo_bwd_list(const PhotoDev& p, cudaStream_t st);   // work-list backward of sparsely selected images
cudaError_t launch_smooth_fwd(const SmoothDev& p, cudaStream_t st);
cudaError_t launch_smooth_bwd(const SmoothDev& p, cudaStream_t st);
cudaError_t launch_area_pyramid(const float* img, int B, int H, int W, float* J, int h, int w, cudaStream_t st);
// every level's area-downsampled image in one launch; also zeroes every level's accumulators (lv[l].acc, 4 doubles per image)
cudaError_t launch_area_pyramid_multi(const float* img, int H, int W, const SmoothDev& p, cudaStream_t st);
cudaError_t launch_edge_finalize_multi(const SmoothDev& p, float* const* loss, cudaStream_t st);
cudaError_t launch_photo_finalize(const PhotoDev& p, const float* photo_coef, const float* smooth_coef,
                                  float* losses, cudaStream_t st);
cudaError_t launch_feat_fwd(const FeatDev& p, cudaStream_t st);
cudaError_t launch_feat_finalize(const FeatDev& p, cudaStream_t st);
cudaError_t launch_feat_bwd(const FeatDev& p, cudaStream_t st);
cudaError_t launch_feat_bwd_gather(const FeatDev& p, cudaStream_t st);      // after launch_feat_bwd when p.G != nullptr
cudaError_t launch_feat_bwd_overflow(const FeatDev& p, cudaStream_t st);    // after the gather
// channel-last (NHWC) feature maps, fp32 or bf16 storage (tdl_feat2.cu)
cudaError_t launch_feat_fwd_nhwc(const FeatDev& p, cudaStream_t st);
cudaError_t launch_feat_bwd_nhwc(const FeatDev& p, cudaStream_t st);
cudaError_t launch_feat_gather_nhwc(const FeatDev& p, cudaStream_t st);
cudaError_t launch_feat_overflow_nhwc(const FeatDev& p, cudaStream_t st);
cudaError_t launch_edge_finalize(const double* acc, int acc_stride, int B, float first_coef, float second_coef,
                                 int h, int w, float* loss, cudaStream_t st);

// input pipeline (tdl_input.cu)
struct InputDev {
    int B, H, W, nframes, erase_count, erase_h, erase_w;
    const unsigned char* frames[TDL_MAX_SRC + 1];
    const float* jitter;
    const int* order;
    const unsigned char* do_aug;
    const unsigned char* do_flip;
    const int* holes;
    float* color[TDL_MAX_SRC + 1];
    float* color_aug[TDL_MAX_SRC + 1];
    float* mask;
    unsigned long long* gsum;    // [B][nframes] grey sums for ImageEnhance.Contrast
};
cudaError_t launch_input_stat(const InputDev& p, cudaStream_t st);
cudaError_t launch_input_apply(const InputDev& p, cudaStream_t st);

cudaError_t launch_proj_fwd(const tdl_proj_args& a, cudaStream_t st);
cudaError_t launch_proj_bwd(const tdl_proj_args& a, cudaStream_t st);
cudaError_t launch_pose_fwd(const float* aa, const float* tr, int B, int invert, float* T, cudaStream_t st);
cudaError_t launch_pose_bwd(const float* aa, const float* tr, const float* dT, int B, int invert, float* d_aa, float* d_tr,
                            cudaStream_t st);

}  // namespace tdl
