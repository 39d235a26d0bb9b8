// extern "C" entry points of libtdl.so (see include/tdl.h): argument validation,
// workspace carving and kernel sequencing.  No torch types, no allocation, no host sync.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "tdl.h"
#include "tdl_internal.h"

using namespace tdl;

namespace {

// ---- process-wide tuning / test switches.  Read from the environment ONCE (first use) and afterwards only changed through
// tdl_set_option(): the entry points below are on the training step's host path and must not call getenv().
enum Opt { kOptNoTma = 0, kOptFusedFwd, kOptSparseMax, kOptFeatAtomic, kOptFeatChunk, kOptPhotoV1, kOptListMax, kOptFeatNoBulk, kOptCount };
const char* const kOptNames[kOptCount] = {"no_tma", "fused_fwd", "photo_sparse_max", "feat_atomic", "feat_chunk", "photo_v1", "photo_list_max", "feat_no_bulk"};
const char* const kOptEnv[kOptCount] = {"TDL_NO_TMA", "TDL_FUSED_FWD", "TDL_PHOTO_SPARSE_MAX", "TDL_FEAT_ATOMIC", "TDL_FEAT_CHUNK", "TDL_PHOTO_V1", "TDL_PHOTO_LIST_MAX", "TDL_FEAT_NO_BULK"};
const int kOptDefault[kOptCount] = {0, 0, 128, 0, 0, 0, kListCap, 0};
std::atomic<int> g_opt[kOptCount];
std::once_flag g_opt_once;

void init_options() {
    for (int i = 0; i < kOptCount; ++i) {
        const char* e = getenv(kOptEnv[i]);
        int v = kOptDefault[i];
        if (e) v = (i == kOptSparseMax || i == kOptFeatChunk || i == kOptListMax) ? atoi(e) : 1;
        g_opt[i].store(v, std::memory_order_relaxed);
    }
}
inline int opt(Opt o) {
    std::call_once(g_opt_once, init_options);
    return g_opt[o].load(std::memory_order_relaxed);
}

// The kernels are sm_100a binaries: any other device (or none) is reported as TDL_ERR_NODEVICE instead of the raw
// cudaErrorNoKernelImageForDevice of the first launch.  The capability is looked up once per device.
int check_device() {
    static std::atomic<int> cc_major[64];
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) {
        cudaGetLastError();
        return TDL_ERR_NODEVICE;
    }
    int major = dev < 64 ? cc_major[dev].load(std::memory_order_relaxed) : 0;
    if (major == 0) {
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
            cudaGetLastError();
            return TDL_ERR_NODEVICE;
        }
        if (dev < 64) cc_major[dev].store(major, std::memory_order_relaxed);
    }
    return major == 10 ? TDL_OK : TDL_ERR_NODEVICE;
}

inline uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

inline bool pow2_factor(int big, int small, int* fac) {
    if (small <= 0 || big % small != 0) return false;
    const int f = big / small;
    if (f < 1 || f > 32 || (f & (f - 1)) != 0) return false;
    *fac = f;
    return true;
}

struct PhotoWsLayout {
    uint64_t acc_off, acc_doubles, j_off[TDL_MAX_SCALES], w_off[TDL_MAX_SCALES], argmin_off, list_off, total;
};

PhotoWsLayout photo_layout(int B, int H, int W, int nscales, const int32_t* dh, const int32_t* dw) {
    PhotoWsLayout L;
    uint64_t off = 0;
    L.acc_off = off;
    // per-image sums, then the work-list header (int count per (scale, image) + the validity word), cleared together
    L.acc_doubles = (uint64_t)nscales * B * 4 + ((uint64_t)nscales * B + 1 + 1) / 2;
    off = align_up(off + L.acc_doubles * sizeof(double), 256);
    for (int s = 0; s < TDL_MAX_SCALES; ++s) {
        L.j_off[s] = off;
        if (s < nscales) off = align_up(off + (uint64_t)B * 3 * dh[s] * dw[s] * sizeof(float), 256);
        L.w_off[s] = off;
        if (s < nscales) off = align_up(off + (uint64_t)B * 6 * dh[s] * dw[s] * sizeof(float), 256);
    }
    L.argmin_off = off;
    off = align_up(off + (uint64_t)nscales * B * H * W, 256);
    L.list_off = off;
    off = align_up(off + (uint64_t)nscales * B * kListCap * sizeof(uint32_t), 256);
    L.total = off;
    return L;
}

int check_photo(const tdl_photo_args* a, bool bwd, PhotoDev* d) {
    if (!a) return TDL_ERR_NULL;
    if (a->S < 1 || a->S > TDL_MAX_SRC || a->nscales < 1 || a->nscales > TDL_MAX_SCALES) return TDL_ERR_COUNT;
    if (a->B < 1 || a->H < 4 || a->W < 4) return TDL_ERR_SHAPE;
    if (!a->target || !a->P || !a->invK || !a->workspace || !a->losses) return TDL_ERR_NULL;
    memset(d, 0, sizeof(*d));
    d->B = a->B; d->H = a->H; d->W = a->W; d->S = a->S; d->nscales = a->nscales;
    for (int f = 0; f < a->S; ++f) {
        if (!a->src[f]) return TDL_ERR_NULL;
        d->src[f] = a->src[f];
    }
    for (int s = 0; s < a->nscales; ++s) {
        if (!a->disp[s]) return TDL_ERR_NULL;
        int fy, fx;
        if (!pow2_factor(a->H, a->disp_h[s], &fy) || !pow2_factor(a->W, a->disp_w[s], &fx) || fy != fx)
            return TDL_ERR_SHAPE;
        d->disp[s] = a->disp[s];
        d->dh[s] = a->disp_h[s]; d->dw[s] = a->disp_w[s]; d->fac[s] = fy;
        d->sy[s] = (float)a->disp_h[s] / (float)a->H;
        d->sx[s] = (float)a->disp_w[s] / (float)a->W;
        d->photo_coef[s] = a->photo_coef[s];
        for (int f = 0; f < a->S; ++f) {
            d->noise[s][f] = a->noise[s][f];
            d->warped[s][f] = a->warped[s][f];
        }
        d->min_index[s] = reinterpret_cast<long long*>(a->min_index[s]);
        if (bwd) {
            if (!a->d_disp[s]) return TDL_ERR_NULL;
            d->d_disp[s] = a->d_disp[s];
        }
    }
    if (bwd && (!a->dlosses || !a->dP)) return TDL_ERR_NULL;
    const PhotoWsLayout L = photo_layout(a->B, a->H, a->W, a->nscales, a->disp_h, a->disp_w);
    if (a->workspace_bytes < L.total) return TDL_ERR_WORKSPACE;
    char* ws = static_cast<char*>(a->workspace);
    d->acc = reinterpret_cast<double*>(ws + L.acc_off);
    for (int s = 0; s < a->nscales; ++s) {
        d->J[s] = reinterpret_cast<float*>(ws + L.j_off[s]);
        d->Wt[s] = reinterpret_cast<float*>(ws + L.w_off[s]);
    }
    d->argmin = reinterpret_cast<unsigned char*>(ws + L.argmin_off);
    d->acc_n = (int)L.acc_doubles;
    d->lcnt = reinterpret_cast<int*>(d->acc + (size_t)a->nscales * a->B * 4);
    d->wlist = reinterpret_cast<uint32_t*>(ws + L.list_off);
    d->automask = a->automask != 0;
    d->use_tma = !opt(kOptNoTma);
    d->split_fwd = !opt(kOptFusedFwd);
    d->v1 = opt(kOptPhotoV1) != 0;
    {
        // tiles with <= 128 selected windows (of 1156 incl. halo, all frames) take the scatter path; 128 is also the hard
        // limit (9 live pixels per window must fit the tile's list).  Tests: 0 forces the dense backward everywhere.
        const int v = opt(kOptSparseMax);
        d->sparse_max = v < 0 ? 0 : (v > 128 ? 128 : v);
    }
    {
        // (image, scale) pairs with <= list_max selected windows (default: the list capacity, 3 % of a 192x640 image) are
        // differentiated from the work list; -1 switches the list path off (tests).  The entries carry the pixel in 28 bits.
        const int v = opt(kOptListMax);
        d->list_max = v < -1 ? -1 : (v > kListCap ? kListCap : v);
        if ((uint64_t)a->H * a->W >= (1ull << 28)) d->list_max = -1;
    }
    d->align_corners = a->align_corners != 0;
    d->min_disp = (float)(1.0 / a->max_depth);
    d->range = (float)(1.0 / a->min_depth - 1.0 / a->max_depth);
    d->seed = a->noise_seed;
    d->target = a->target; d->P = a->P; d->invK = a->invK;
    d->dlosses = a->dlosses; d->dP = a->dP;
    return TDL_OK;
}

void photo_smooth_levels(const tdl_photo_args* a, const PhotoDev& d, bool bwd, SmoothDev* sm) {
    memset(sm, 0, sizeof(*sm));
    sm->B = a->B;
    sm->nlevels = a->nscales;
    for (int s = 0; s < a->nscales; ++s) {
        SmoothLevel& L = sm->lv[s];
        L.C = 1; L.h = d.dh[s]; L.w = d.dw[s];
        L.x = d.disp[s]; L.J = d.J[s]; L.Wt = d.Wt[s];
        L.acc = d.acc + (size_t)s * a->B * 4; L.acc_stride = 4;
        L.norm = a->disp_norm != 0;
        L.alpha = a->smooth_alpha;
        L.first_coef = a->smooth_coef[s]; L.second_coef = a->smooth_coef[s];
        if (bwd) {
            L.dloss = a->dlosses + a->nscales + s;
            L.dx = a->d_disp[s];
        }
    }
}

#define TDL_CUDA(expr)                       \
    do {                                     \
        cudaError_t e__ = (expr);            \
        if (e__ != cudaSuccess) return (int)e__; \
    } while (0)

// ---- optional per-kernel timing (tdl_profile_begin / tdl_profile_end): a cudaEvent pair is recorded on
// the launch stream around every kernel, so bench.py can attribute device time to kernels without ncu.
struct ProfRec {
    const char* name;
    cudaEvent_t a, b;
};
struct Profiler {
    std::atomic<bool> on{false};
    std::mutex mu;                       // guards recs / pool: entry points may be called from several host threads
    std::vector<ProfRec> recs;
    std::vector<cudaEvent_t> pool;
    cudaEvent_t get() {
        if (!pool.empty()) {
            cudaEvent_t e = pool.back();
            pool.pop_back();
            return e;
        }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
} g_prof;

struct ProfScope {
    cudaStream_t st;
    bool active;
    cudaEvent_t end = nullptr;
    ProfScope(const char* name, cudaStream_t s) : st(s), active(g_prof.on.load(std::memory_order_relaxed)) {
        if (active) {
            std::lock_guard<std::mutex> lk(g_prof.mu);
            ProfRec r{name, g_prof.get(), g_prof.get()};
            end = r.b;
            cudaEventRecord(r.a, st);
            g_prof.recs.push_back(r);
        }
    }
    ~ProfScope() {
        if (active) cudaEventRecord(end, st);
    }
};

#define TDL_KERNEL(name, expr)               \
    do {                                     \
        ProfScope scope__(name, st);         \
        cudaError_t e__ = (expr);            \
        if (e__ != cudaSuccess) return (int)e__; \
    } while (0)

}  // namespace

extern "C" {

int tdl_abi_version(void) { return TDL_ABI_VERSION; }

const char* tdl_strerror(int code) {
    switch (code) {
        case TDL_OK: return "ok";
        case TDL_ERR_NULL: return "tdl: a required pointer is NULL";
        case TDL_ERR_SHAPE: return "tdl: unsupported shape (disp size must divide the image size by a power of two <= 32)";
        case TDL_ERR_WORKSPACE: return "tdl: workspace too small (see *_ws_bytes)";
        case TDL_ERR_COUNT: return "tdl: S / nscales / C out of range";
        case TDL_ERR_NODEVICE: return "tdl: no CUDA device, or the current device is not sm_100 (B200): there is no other code path";
        case TDL_ERR_OPTION: return "tdl: unknown option name";
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "tdl: unknown error";
}

int tdl_set_option(const char* name, int value) {
    if (!name) return TDL_ERR_NULL;
    for (int i = 0; i < kOptCount; ++i)
        if (!strcmp(name, kOptNames[i])) {
            opt((Opt)i);                                   // make sure the environment defaults were read first
            g_opt[i].store(value, std::memory_order_relaxed);
            return TDL_OK;
        }
    return TDL_ERR_OPTION;
}

int tdl_get_option(const char* name, int* value) {
    if (!name || !value) return TDL_ERR_NULL;
    for (int i = 0; i < kOptCount; ++i)
        if (!strcmp(name, kOptNames[i])) {
            *value = opt((Opt)i);
            return TDL_OK;
        }
    return TDL_ERR_OPTION;
}

int tdl_launch_count(const char* entry) {
    if (!entry) return 0;
    if (!strcmp(entry, "tdl_photo_fwd")) return 4;        // photo_warp + photo_score (or fused photo_fwd), smooth_fwd, finalize
    if (!strcmp(entry, "tdl_photo_bwd")) return opt(kOptListMax) >= 0 ? 3 : 2;   // smooth_bwd, photo_bwd, photo_bwd_list
    if (!strcmp(entry, "tdl_feat_fwd")) return 2;         // feat_fwd, finalize
    if (!strcmp(entry, "tdl_feat_bwd")) return 1;         // feat_bwd (atomic scatter / frozen features)
    if (!strcmp(entry, "tdl_feat_bwd:gather")) return 3;  // feat_bwd (bucket) + feat_gather + feat_overflow (bwd_scratch given)
    if (!strcmp(entry, "tdl_edge_smooth_fwd")) return 3;  // area pyramid, smooth_fwd, finalize
    if (!strcmp(entry, "tdl_edge_smooth_bwd")) return 1;
    if (!strcmp(entry, "tdl_edge_smooth_multi_fwd")) return 3;   // all levels: area pyramids, smooth_fwd, finalize
    if (!strcmp(entry, "tdl_edge_smooth_multi_bwd")) return 1;
    if (!strcmp(entry, "tdl_recon_fwd")) return 2;        // recon_fwd, finalize
    if (!strcmp(entry, "tdl_recon_bwd")) return 1;
    if (!strcmp(entry, "tdl_pose_fwd") || !strcmp(entry, "tdl_pose_bwd")) return 1;
    if (!strcmp(entry, "tdl_proj_fwd") || !strcmp(entry, "tdl_proj_bwd")) return 1;
    if (!strcmp(entry, "tdl_input_fwd")) return 2;        // input_stat (with jitter), input_apply
    return 0;
}

int tdl_profile_begin(void) {
    g_prof.on = true;
    return TDL_OK;
}

// Synchronises the recorded events, aggregates per kernel name and writes up to `max_entries` rows
// (name, launches, total milliseconds).  Returns the number of rows; profiling is switched off.
int tdl_profile_end(tdl_kernel_time* out, int max_entries) {
    g_prof.on = false;
    std::lock_guard<std::mutex> lk(g_prof.mu);
    std::map<std::string, std::pair<int, double>> agg;
    std::vector<std::string> order;
    for (ProfRec& r : g_prof.recs) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            if (!agg.count(r.name)) order.push_back(r.name);
            agg[r.name].first += 1;
            agg[r.name].second += ms;
        }
        g_prof.pool.push_back(r.a);
        g_prof.pool.push_back(r.b);
    }
    g_prof.recs.clear();
    int n = 0;
    for (const std::string& k : order) {
        if (n >= max_entries || !out) break;
        strncpy(out[n].name, k.c_str(), sizeof(out[n].name) - 1);
        out[n].name[sizeof(out[n].name) - 1] = 0;
        out[n].launches = agg[k].first;
        out[n].total_ms = agg[k].second;
        ++n;
    }
    return n;
}

uint64_t tdl_photo_ws_bytes(int32_t B, int32_t H, int32_t W, int32_t S, int32_t nscales, const int32_t* disp_h,
                            const int32_t* disp_w) {
    (void)S;
    if (B < 1 || H < 1 || W < 1 || nscales < 1 || nscales > TDL_MAX_SCALES || !disp_h || !disp_w) return 0;
    return photo_layout(B, H, W, nscales, disp_h, disp_w).total;
}

int tdl_photo_fwd(const tdl_photo_args* a, tdl_stream_t stream) {
    PhotoDev d;
    int rc = check_photo(a, false, &d);
    if (rc == TDL_OK) rc = check_device();
    if (rc != TDL_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (photo_fwd_can_split(d)) {                             // (photo_warp clears the accumulators itself)
        TDL_KERNEL("photo_warp", launch_photo_warp(d, st));
        TDL_KERNEL("photo_score", d.v1 ? launch_photo_score(d, st) : launch_photo_score2(d, st));
    } else {
        TDL_KERNEL("memset", cudaMemsetAsync(d.acc, 0, (size_t)d.acc_n * sizeof(double), st));
        TDL_KERNEL("photo_fwd", launch_photo_fwd(d, st));
    }
    SmoothDev sm;
    photo_smooth_levels(a, d, false, &sm);
    TDL_KERNEL("smooth_fwd", launch_smooth_fwd(sm, st));
    TDL_KERNEL("photo_finalize", launch_photo_finalize(d, a->photo_coef, a->smooth_coef, a->losses, st));
    return TDL_OK;
}

int tdl_photo_bwd(const tdl_photo_args* a, tdl_stream_t stream) {
    PhotoDev d;
    int rc = check_photo(a, true, &d);
    if (rc == TDL_OK) rc = check_device();
    if (rc != TDL_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SmoothDev sm;
    photo_smooth_levels(a, d, true, &sm);
    sm.zero_ptr = a->dP;                                      // dP (accumulated by photo_bwd) is cleared by smooth_bwd's first CTA
    sm.zero_n = a->B * a->S * 12;
    TDL_KERNEL("smooth_bwd", launch_smooth_bwd(sm, st));       // writes d_disp[s] (=), the photometric kernel adds to it
    TDL_KERNEL("photo_bwd", launch_photo_bwd(d, st));
    if (d.list_max >= 0 && photo_fwd_can_split(d)) TDL_KERNEL("photo_bwd_list", launch_photo_bwd_list(d, st));
    return TDL_OK;
}

// ------------------------------------------------------------------------------------ feature-metric
uint64_t tdl_feat_ws_bytes(int32_t B, int32_t C, int32_t h, int32_t w, int32_t S) {
    (void)C; (void)S;
    if (B < 1 || h < 1 || w < 1) return 0;
    return align_up((uint64_t)B * sizeof(double), 256) + align_up((uint64_t)B * h * w, 256);
}

// backward scratch of the bucketed d_src gather: [ov_cnt[B] | bk_cnt] (zeroed per call) | bk_ent | ov_ent[B][4hw] | G
struct FeatScratchLayout {
    uint64_t cnt_off, hdr_bytes, cnt_bytes, ent_off, ov_off, g_off, total;
};
// The bucketed backward runs in batch chunks whose G array (chunk x h*w x C floats) stays L2-resident between the
// per-pixel kernel that writes it and the gather kernel that reads every row ~4 times: without chunking the step rate
// drops by 10-15 % once G outgrows the 126 MB L2 (batch 64 at 192x640, measured).  It also bounds the scratch size.
static int feat_chunk_images(int B, int C, int h, int w) {
    const uint64_t per_image = (uint64_t)h * w * C * sizeof(float);
    const uint64_t budget = 64ull << 20;
    uint64_t n = per_image ? budget / per_image : 1;
    if (const int forced = opt(kOptFeatChunk)) n = (uint64_t)forced;         // tests: force small chunks
    if (n < 1) n = 1;
    return (int)(n < (uint64_t)B ? n : (uint64_t)B);
}

static FeatScratchLayout feat_scratch_layout(int Bfull, int C, int h, int w, int S) {
    FeatScratchLayout L;
    const int B = feat_chunk_images(Bfull, C, h, w);
    const uint64_t hw = (uint64_t)h * w;
    L.cnt_off = 0;
    L.hdr_bytes = align_up((uint64_t)B * sizeof(int), 256) + align_up((uint64_t)B * S * 12 * sizeof(float), 256);   // ov_cnt[B] | dP_acc[B][S][12]
    L.cnt_bytes = L.hdr_bytes + align_up((uint64_t)S * B * hw * sizeof(int), 256);
    L.ent_off = L.cnt_off + L.cnt_bytes;
    L.ov_off = L.ent_off + align_up((uint64_t)S * B * hw * kFeatBucketCap * sizeof(int2), 256);
    L.g_off = L.ov_off + align_up(4 * (uint64_t)B * hw * sizeof(int4), 256);
    L.total = L.g_off + align_up((uint64_t)B * hw * C * sizeof(float), 256);
    return L;
}

uint64_t tdl_feat_bwd_scratch_bytes(int32_t B, int32_t C, int32_t h, int32_t w, int32_t S) {
    if (B < 1 || C < 1 || h < 1 || w < 1 || S < 1 || S > TDL_MAX_SRC) return 0;
    return feat_scratch_layout(B, C, h, w, S).total;
}

static int check_feat(const tdl_feat_args* a, bool bwd, FeatDev* d) {
    if (!a) return TDL_ERR_NULL;
    if (a->S < 1 || a->S > TDL_MAX_SRC || a->C < 1) return TDL_ERR_COUNT;
    if (a->B < 1 || a->h < 2 || a->w < 2 || a->disp_h < 1 || a->disp_w < 1) return TDL_ERR_SHAPE;
    if (!a->tgt || !a->disp || !a->P || !a->invK || !a->workspace || !a->loss) return TDL_ERR_NULL;
    if (a->workspace_bytes < tdl_feat_ws_bytes(a->B, a->C, a->h, a->w, a->S)) return TDL_ERR_WORKSPACE;
    memset(d, 0, sizeof(*d));
    d->B = a->B; d->C = a->C; d->h = a->h; d->w = a->w; d->S = a->S;
    d->Bnorm = a->B;
    d->dh = a->disp_h; d->dw = a->disp_w;
    d->sy = (float)a->disp_h / (float)a->h;
    d->sx = (float)a->disp_w / (float)a->w;
    d->align_corners = a->align_corners != 0;
    d->min_disp = (float)(1.0 / a->max_depth);
    d->range = (float)(1.0 / a->min_depth - 1.0 / a->max_depth);
    d->coef = a->coef;
    if (a->layout != TDL_LAYOUT_NCHW && a->layout != TDL_LAYOUT_NHWC) return TDL_ERR_SHAPE;
    if (a->dtype != TDL_DTYPE_F32 && a->dtype != TDL_DTYPE_BF16) return TDL_ERR_SHAPE;
    if (a->dtype == TDL_DTYPE_BF16 && a->layout != TDL_LAYOUT_NHWC) return TDL_ERR_SHAPE;          // bf16 rows only
    if (a->layout == TDL_LAYOUT_NHWC && a->C % 4 != 0) return TDL_ERR_COUNT;
    // the channel-last kernels index one image with 32-bit element offsets (pixel index also carries two flag bits)
    if (a->layout == TDL_LAYOUT_NHWC && ((uint64_t)a->h * a->w * a->C >= (1ull << 31) || (uint64_t)a->h * a->w >= (1ull << 29)))
        return TDL_ERR_SHAPE;
    d->layout = a->layout;
    d->dtype = a->dtype;
    d->bulk = !opt(kOptFeatNoBulk);
    d->tgt = static_cast<const float*>(a->tgt); d->disp = a->disp; d->P = a->P; d->invK = a->invK;
    int n_dsrc = 0;
    for (int f = 0; f < a->S; ++f) {
        if (!a->src[f]) return TDL_ERR_NULL;
        d->src[f] = static_cast<const float*>(a->src[f]);
        d->warped[f] = static_cast<float*>(a->warped[f]);
        d->d_src[f] = static_cast<float*>(a->d_src[f]);
        n_dsrc += a->d_src[f] != nullptr;
    }
    if (bwd && n_dsrc != 0 && n_dsrc != a->S) return TDL_ERR_NULL;     // all or none
    d->min_index = reinterpret_cast<long long*>(a->min_index);
    char* ws = static_cast<char*>(a->workspace);
    d->acc = reinterpret_cast<double*>(ws);
    d->argmin = reinterpret_cast<unsigned char*>(ws + align_up((uint64_t)a->B * sizeof(double), 256));
    d->loss = a->loss;
    if (bwd) {
        if (!a->dloss || !a->d_disp || !a->dP) return TDL_ERR_NULL;
        d->dloss = a->dloss; d->d_tgt = static_cast<float*>(a->d_tgt); d->d_disp = a->d_disp; d->dP = a->dP;
        const FeatScratchLayout L = feat_scratch_layout(a->B, a->C, a->h, a->w, a->S);
        const bool scratch_ok = a->bwd_scratch && a->bwd_scratch_bytes >= L.total && (reinterpret_cast<uintptr_t>(a->bwd_scratch) & 15) == 0;
        if (a->layout == TDL_LAYOUT_NHWC && n_dsrc == a->S && !scratch_ok) return TDL_ERR_WORKSPACE;   // no atomic fallback for rows
        if (n_dsrc == a->S && scratch_ok && a->C % 4 == 0 && (uint64_t)a->h * a->w * a->C < (1ull << 28) &&
            (a->layout == TDL_LAYOUT_NHWC || !opt(kOptFeatAtomic))) {
            char* sc = static_cast<char*>(a->bwd_scratch);
            d->ov_cnt = reinterpret_cast<int*>(sc + L.cnt_off);
            d->dP_acc = reinterpret_cast<float*>(sc + L.cnt_off + align_up((uint64_t)feat_chunk_images(a->B, a->C, a->h, a->w) * sizeof(int), 256));
            d->bk_cnt = reinterpret_cast<int*>(sc + L.cnt_off + L.hdr_bytes);
            d->bk_ent = reinterpret_cast<int2*>(sc + L.ent_off);
            d->ov_ent = reinterpret_cast<int4*>(sc + L.ov_off);
            d->G = reinterpret_cast<float*>(sc + L.g_off);
        }
    }
    return TDL_OK;
}

int tdl_feat_fwd(const tdl_feat_args* a, tdl_stream_t stream) {
    FeatDev d;
    int rc = check_feat(a, false, &d);
    if (rc == TDL_OK) rc = check_device();
    if (rc != TDL_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TDL_KERNEL("memset", cudaMemsetAsync(d.acc, 0, (size_t)a->B * sizeof(double), st));
    TDL_KERNEL("feat_fwd", d.layout == TDL_LAYOUT_NHWC ? launch_feat_fwd_nhwc(d, st) : launch_feat_fwd(d, st));
    TDL_KERNEL("feat_finalize", launch_feat_finalize(d, st));
    return TDL_OK;
}

int tdl_feat_bwd(const tdl_feat_args* a, tdl_stream_t stream) {
    FeatDev d;
    int rc = check_feat(a, true, &d);
    if (rc == TDL_OK) rc = check_device();
    if (rc != TDL_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t esz = a->dtype == TDL_DTYPE_BF16 ? 2 : sizeof(float);
    const size_t fbytes = (size_t)a->B * a->C * a->h * a->w * esz;
    const bool nhwc = a->layout == TDL_LAYOUT_NHWC;
    // per-image offset of a feature tensor, in units of the float* the descriptors carry (bf16 images are half as long)
    auto feat_off = [&](const float* base, size_t b0, size_t img) {
        return reinterpret_cast<const float*>(reinterpret_cast<const char*>(base) + b0 * img * esz);
    };
    if (!d.G)        // (bucketed path: dP is accumulated in the zeroed scratch header and written by the gather kernel)
        TDL_KERNEL("memset", cudaMemsetAsync(a->dP, 0, (size_t)a->B * a->S * 12 * sizeof(float), st));
    if (a->disp_h != a->h || a->disp_w != a->w)
        TDL_KERNEL("memset", cudaMemsetAsync(a->d_disp, 0, (size_t)a->B * a->disp_h * a->disp_w * sizeof(float), st));
    if (d.G) {       // bucketed gather: per-pixel kernel registers taps + writes G, gather kernel writes d_src (no memset)
        const FeatScratchLayout L = feat_scratch_layout(a->B, a->C, a->h, a->w, a->S);
        const int Bc = feat_chunk_images(a->B, a->C, a->h, a->w);
        const size_t img = (size_t)a->C * a->h * a->w;
        for (int b0 = 0; b0 < a->B; b0 += Bc) {
            FeatDev c = d;                                    // this chunk's view of every per-image tensor
            c.B = a->B - b0 < Bc ? a->B - b0 : Bc;
            c.tgt = feat_off(d.tgt, b0, img);
            c.disp += (size_t)b0 * a->disp_h * a->disp_w;
            c.P += (size_t)b0 * a->S * 12;
            c.invK += (size_t)b0 * 9;
            c.argmin += (size_t)b0 * a->h * a->w;
            if (c.d_tgt) c.d_tgt = const_cast<float*>(feat_off(d.d_tgt, b0, img));
            c.d_disp += (size_t)b0 * a->disp_h * a->disp_w;
            c.dP += (size_t)b0 * a->S * 12;
            for (int f = 0; f < a->S; ++f) {
                c.src[f] = feat_off(d.src[f], b0, img);
                c.d_src[f] = const_cast<float*>(feat_off(d.d_src[f], b0, img));
            }
            TDL_KERNEL("memset", cudaMemsetAsync(c.ov_cnt, 0, L.cnt_bytes, st));
            TDL_KERNEL("feat_bwd", nhwc ? launch_feat_bwd_nhwc(c, st) : launch_feat_bwd(c, st));
            TDL_KERNEL("feat_gather", nhwc ? launch_feat_gather_nhwc(c, st) : launch_feat_bwd_gather(c, st));
            TDL_KERNEL("feat_overflow", nhwc ? launch_feat_overflow_nhwc(c, st) : launch_feat_bwd_overflow(c, st));
        }
        return TDL_OK;
    }
    if (nhwc) {                                       // frozen extractor: no d_src (checked above), d_tgt optional
        TDL_KERNEL("feat_bwd", launch_feat_bwd_nhwc(d, st));
        return TDL_OK;
    }
    for (int f = 0; f < a->S; ++f)
        if (a->d_src[f]) TDL_KERNEL("memset_dsrc", cudaMemsetAsync(a->d_src[f], 0, fbytes, st));
    TDL_KERNEL("feat_bwd", launch_feat_bwd(d, st));
    return TDL_OK;
}

// ------------------------------------------------------------------------------------ edge-aware smoothness
uint64_t tdl_edge_ws_bytes(int32_t B, int32_t C, int32_t h, int32_t w) {
    (void)C;
    if (B < 1 || h < 1 || w < 1) return 0;
    return align_up((uint64_t)B * 4 * sizeof(double), 256) + align_up((uint64_t)B * 3 * h * w * sizeof(float), 256) +
           align_up((uint64_t)B * 6 * h * w * sizeof(float), 256);
}

static int check_edge(const tdl_edge_args* a, bool bwd, SmoothDev* sm, float** J) {
    if (!a) return TDL_ERR_NULL;
    if (a->B < 1 || a->C < 1) return TDL_ERR_COUNT;
    int fy, fx;
    if (!pow2_factor(a->H, a->h, &fy) || !pow2_factor(a->W, a->w, &fx) || fy != fx) return TDL_ERR_SHAPE;
    if (!a->feature || !a->image || !a->workspace || !a->loss) return TDL_ERR_NULL;
    if (a->workspace_bytes < tdl_edge_ws_bytes(a->B, a->C, a->h, a->w)) return TDL_ERR_WORKSPACE;
    if (bwd && (!a->dloss || !a->d_feature)) return TDL_ERR_NULL;
    char* ws = static_cast<char*>(a->workspace);
    *J = reinterpret_cast<float*>(ws + align_up((uint64_t)a->B * 4 * sizeof(double), 256));
    memset(sm, 0, sizeof(*sm));
    sm->B = a->B;
    sm->nlevels = 1;
    SmoothLevel& L = sm->lv[0];
    L.C = a->C; L.h = a->h; L.w = a->w;
    L.x = a->feature; L.J = *J;
    L.Wt = *J + align_up((uint64_t)a->B * 3 * a->h * a->w * sizeof(float), 256) / sizeof(float);
    L.acc = reinterpret_cast<double*>(ws); L.acc_stride = 4;
    L.norm = 0; L.alpha = a->alpha;
    L.first_coef = a->first_coef; L.second_coef = a->second_coef;
    L.dloss = a->dloss; L.dx = a->d_feature;
    return TDL_OK;
}

int tdl_edge_smooth_fwd(const tdl_edge_args* a, tdl_stream_t stream) {
    SmoothDev sm;
    float* J;
    int rc = check_edge(a, false, &sm, &J);
    if (rc == TDL_OK) rc = check_device();
    if (rc != TDL_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TDL_KERNEL("memset", cudaMemsetAsync(sm.lv[0].acc, 0, (size_t)a->B * 4 * sizeof(double), st));
    TDL_KERNEL("area_pyramid", launch_area_pyramid(a->image, a->B, a->H, a->W, J, a->h, a->w, st));
    TDL_KERNEL("edge_smooth_fwd", launch_smooth_fwd(sm, st));
    TDL_KERNEL("edge_finalize", launch_edge_finalize(sm.lv[0].acc, 4, a->B, a->first_coef, a->second_coef, a->h, a->w, a->loss, st));
    return TDL_OK;
}

int tdl_edge_smooth_bwd(const tdl_edge_args* a, tdl_stream_t stream) {
    SmoothDev sm;
    float* J;
    int rc = check_edge(a, true, &sm, &J);
    if (rc == TDL_OK) rc = check_device();
    if (rc != TDL_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TDL_KERNEL("edge_smooth_bwd", launch_smooth_bwd(sm, st));
    return TDL_OK;
}

static int check_edge_multi(const tdl_edge_multi_args* m, bool bwd, SmoothDev* sm) {
    if (!m) return TDL_ERR_NULL;
    if (m->nlevels < 1 || m->nlevels > TDL_MAX_LEVELS) return TDL_ERR_COUNT;
    memset(sm, 0, sizeof(*sm));
    for (int l = 0; l < m->nlevels; ++l) {
        SmoothDev one;
        float* J;
        const int rc = check_edge(&m->level[l], bwd, &one, &J);
        if (rc != TDL_OK) return rc;
        const tdl_edge_args& a = m->level[l];
        if (a.B != m->level[0].B || a.H != m->level[0].H || a.W != m->level[0].W || a.image != m->level[0].image) return TDL_ERR_SHAPE;
        sm->lv[l] = one.lv[0];
    }
    sm->B = m->level[0].B;
    sm->nlevels = m->nlevels;
    return TDL_OK;
}

int tdl_edge_smooth_multi_fwd(const tdl_edge_multi_args* m, tdl_stream_t stream) {
    SmoothDev sm;
    int rc = check_edge_multi(m, false, &sm);
    if (rc == TDL_OK) rc = check_device();
    if (rc != TDL_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* loss[TDL_MAX_LEVELS];
    for (int l = 0; l < m->nlevels; ++l) loss[l] = m->level[l].loss;
    TDL_KERNEL("area_pyramid", launch_area_pyramid_multi(m->level[0].image, m->level[0].H, m->level[0].W, sm, st));
    TDL_KERNEL("edge_smooth_fwd", launch_smooth_fwd(sm, st));
    TDL_KERNEL("edge_finalize", launch_edge_finalize_multi(sm, loss, st));
    return TDL_OK;
}

int tdl_edge_smooth_multi_bwd(const tdl_edge_multi_args* m, tdl_stream_t stream) {
    SmoothDev sm;
    int rc = check_edge_multi(m, true, &sm);
    if (rc == TDL_OK) rc = check_device();
    if (rc != TDL_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TDL_KERNEL("edge_smooth_bwd", launch_smooth_bwd(sm, st));
    return TDL_OK;
}

// ------------------------------------------------------------------------------------ masked reconstruction
uint64_t tdl_recon_ws_bytes(void) { return 256; }

static int check_recon(const tdl_recon_args* a, bool bwd, ReconArgsDev* d) {
    if (!a) return TDL_ERR_NULL;
    if (a->B < 1 || a->h < 4 || a->w < 4) return TDL_ERR_SHAPE;
    if (!a->pred || !a->target || !a->workspace || !a->loss) return TDL_ERR_NULL;
    if (a->workspace_bytes < tdl_recon_ws_bytes()) return TDL_ERR_WORKSPACE;
    if (bwd && (!a->dloss || !a->d_pred)) return TDL_ERR_NULL;
    *d = ReconArgsDev{a->B, a->h, a->w, a->coef, a->pred, a->target, a->mask,
                      reinterpret_cast<double*>(a->workspace), a->loss, a->dloss, a->d_pred};
    return TDL_OK;
}

int tdl_recon_fwd(const tdl_recon_args* a, tdl_stream_t stream) {
    ReconArgsDev d;
    int rc = check_recon(a, false, &d);
    if (rc == TDL_OK) rc = check_device();
    if (rc != TDL_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TDL_KERNEL("memset", cudaMemsetAsync(d.acc, 0, 2 * sizeof(double), st));
    TDL_KERNEL("recon_fwd", launch_recon_fwd(d, st));
    return TDL_OK;
}

int tdl_recon_bwd(const tdl_recon_args* a, tdl_stream_t stream) {
    ReconArgsDev d;
    int rc = check_recon(a, true, &d);
    if (rc == TDL_OK) rc = check_device();
    if (rc != TDL_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TDL_KERNEL("recon_bwd", launch_recon_bwd(d, st));
    return TDL_OK;
}

// ------------------------------------------------------------------------------------ pose prologue
int tdl_pose_fwd(const tdl_pose_args* a, tdl_stream_t stream) {
    if (!a || !a->axisangle || !a->translation || !a->T) return TDL_ERR_NULL;
    if (a->B < 1) return TDL_ERR_SHAPE;
    if (const int rc = check_device()) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TDL_KERNEL("pose_fwd", launch_pose_fwd(a->axisangle, a->translation, a->B, a->invert, a->T, st));
    return TDL_OK;
}

int tdl_pose_bwd(const tdl_pose_args* a, tdl_stream_t stream) {
    if (!a || !a->axisangle || !a->translation || !a->dT || !a->d_axisangle || !a->d_translation) return TDL_ERR_NULL;
    if (a->B < 1) return TDL_ERR_SHAPE;
    if (const int rc = check_device()) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TDL_KERNEL("pose_bwd", launch_pose_bwd(a->axisangle, a->translation, a->dT, a->B, a->invert, a->d_axisangle,
                                           a->d_translation, st));
    return TDL_OK;
}

// ------------------------------------------------------------------------------------ projection prologue
int tdl_proj_fwd(const tdl_proj_args* a, tdl_stream_t stream) {
    if (!a || !a->K || !a->inv_K || !a->P_full || !a->P_half || !a->invK3 || !a->invKh3) return TDL_ERR_NULL;
    if (a->S < 1 || a->S > TDL_MAX_SRC) return TDL_ERR_COUNT;
    if (a->B < 1) return TDL_ERR_SHAPE;
    for (int f = 0; f < a->S; ++f)
        if (!a->T[f]) return TDL_ERR_NULL;
    if (const int rc = check_device()) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TDL_KERNEL("proj_fwd", launch_proj_fwd(*a, st));
    return TDL_OK;
}

int tdl_proj_bwd(const tdl_proj_args* a, tdl_stream_t stream) {
    if (!a || !a->K) return TDL_ERR_NULL;
    if (a->S < 1 || a->S > TDL_MAX_SRC) return TDL_ERR_COUNT;
    if (a->B < 1) return TDL_ERR_SHAPE;
    if (const int rc = check_device()) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TDL_KERNEL("proj_bwd", launch_proj_bwd(*a, st));
    return TDL_OK;
}

// ------------------------------------------------------------------------------------ input pipeline
uint64_t tdl_input_ws_bytes(int32_t B, int32_t nframes) {
    if (B < 1 || nframes < 1) return 0;
    return align_up((uint64_t)B * nframes * sizeof(unsigned long long), 256);
}

int tdl_input_fwd(const tdl_input_args* a, tdl_stream_t stream) {
    if (!a) return TDL_ERR_NULL;
    if (a->nframes < 1 || a->nframes > TDL_MAX_SRC + 1) return TDL_ERR_COUNT;
    if (a->B < 1 || a->H < 1 || a->W < 1 || (int64_t)a->H * a->W > (int64_t)1 << 30) return TDL_ERR_SHAPE;
    if (a->erase_count < 0 || (a->erase_count > 0 && (a->erase_h < 1 || a->erase_w < 1))) return TDL_ERR_SHAPE;
    if (a->erase_count > 64) return TDL_ERR_COUNT;
    for (int f = 0; f < a->nframes; ++f)
        if (!a->frames[f]) return TDL_ERR_NULL;
    if (a->jitter && (!a->order || !a->do_aug || !a->workspace)) return TDL_ERR_NULL;
    if (a->mask && a->erase_count > 0 && !a->holes) return TDL_ERR_NULL;
    if (a->jitter && a->workspace_bytes < tdl_input_ws_bytes(a->B, a->nframes)) return TDL_ERR_WORKSPACE;
    if (const int rc = check_device()) return rc;
    InputDev d{};
    d.B = a->B; d.H = a->H; d.W = a->W; d.nframes = a->nframes;
    d.erase_count = a->mask ? a->erase_count : 0; d.erase_h = a->erase_h; d.erase_w = a->erase_w;
    for (int f = 0; f < a->nframes; ++f) {
        d.frames[f] = a->frames[f];
        d.color[f] = a->color[f];
        d.color_aug[f] = a->color_aug[f];
    }
    d.jitter = a->jitter; d.order = a->order; d.do_aug = a->do_aug; d.do_flip = a->do_flip; d.holes = a->holes;
    d.mask = a->mask;
    d.gsum = static_cast<unsigned long long*>(a->workspace);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (d.jitter) {
        TDL_KERNEL("memset", cudaMemsetAsync(d.gsum, 0, (size_t)d.B * d.nframes * sizeof(unsigned long long), st));
        TDL_KERNEL("input_stat", launch_input_stat(d, st));
    }
    TDL_KERNEL("input_apply", launch_input_apply(d, st));
    return TDL_OK;
}

}  // extern "C"
