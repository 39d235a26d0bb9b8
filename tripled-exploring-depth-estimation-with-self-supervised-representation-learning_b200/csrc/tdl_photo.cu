// Fused photometric view-synthesis loss, forward and backward (sm_100a).
//
// Forward  (one CTA = one 32x32 full-resolution tile of one image, ALL scales):
//   target + source tiles (halo 1, reflect) -> shared memory once; target window
//   statistics and the identity (automask) reprojection errors are computed once
//   and reused by every scale; per scale the tile (+halo) is re-projected through
//   disp_s / P_f, the sources are sampled bilinearly (border) into shared memory,
//   3x3 SSIM + robust L1 are evaluated with a register sliding window, the
//   per-pixel minimum / first-argmin over [identity..., warped...] is taken and the
//   per-image sums are reduced with warp shuffles.  Side products written from the
//   tile already in shared memory: the warped images, min_index, the u8 argmin mask
//   (workspace), the area-downsampled target pyramid and the per-image disparity sums
//   that the smoothness kernel needs.
// Backward (one CTA = one tile of one image at ONE scale): recomputes the warp with
//   halo 2, evaluates the SSIM window adjoint, the bilinear-sampling / projection /
//   back-projection adjoints, reduces dP with shuffles and folds the bilinear
//   up-sampling adjoint of disp_s inside the tile before touching global memory.
//
// Reference behaviour: mono/model/mono_fm/net.py:63-67,90-106,157-170 and
// mono/model/mono_fm/layers.py:57-107 (see include/tdl.h).
#include "tdl_common.cuh"
#include "tdl_internal.h"
#include "tdl_ssim.cuh"
#include "tdl_tma.cuh"

namespace tdl {

constexpr int kTW = 32;        // tile width  (one warp lane per column)
constexpr int kTH = 32;        // tile height
constexpr int kNT = 256;       // threads per CTA
constexpr int kR = 4;          // rows per thread (vertical strip)
static_assert(kTH == (kNT / 32) * kR, "tile / thread mapping");

// ------------------------------------------------------------------------------------------------
// SSIM + robust-L1 over a vertical strip of kR pixels of one channel.
// xs / ys point at the channel plane in shared memory with pitch `pitch`; (r, c) is the
// top-left corner of the 3x3 window of the strip's first pixel.
// nn.AvgPool2d(3,1) of ATen (CPU and CUDA alike) accumulates the window in row-major order and
// divides by 9; the same order is used here so that, for identical inputs, the statistics -- and with
// them the rounding noise that torch.clamp(.., 0, 1) rectifies when SSIM ~ 0 -- are bit-identical.
// The strip is streamed row by row: row j is the first row of window j, the second of window j-1 and the
// third of window j-2, so each window still sees its nine values in row-major order while only three
// windows are in flight (keeps the register footprint small).
TDL_DEV void acc_row(float& a, bool first, float v0, float v1, float v2) {
    a = first ? v0 : __fadd_rn(a, v0);
    a = __fadd_rn(a, v1);
    a = __fadd_rn(a, v2);
}

template <int PITCH>
TDL_DEV void strip_target_stats(const float* __restrict__ ys, int r, int c, float mu_y[kR], float sg_y[kR]) {
    float ay[kR], ayy[kR];
#pragma unroll
    for (int j = 0; j < kR + 2; ++j) {
        const float* yr = ys + (r + j) * PITCH + c;
        const float y0 = yr[0], y1 = yr[1], y2 = yr[2];
        const float q0 = __fmul_rn(y0, y0), q1 = __fmul_rn(y1, y1), q2 = __fmul_rn(y2, y2);
#pragma unroll
        for (int i = 0; i < kR; ++i) {
            if (i <= j && j <= i + 2) {
                acc_row(ay[i], j == i, y0, y1, y2);
                acc_row(ayy[i], j == i, q0, q1, q2);
            }
        }
        if (j >= 2) {
            const int i = j - 2;
            const float m = div9(ay[i]);
            mu_y[i] = m;
            sg_y[i] = __fsub_rn(div9(ayy[i]), __fmul_rn(m, m));                           // layers.py:102
        }
    }
}

TDL_DEV float ssim_value(float mu_x, float mu_y, float sg_x, float sg_y, float sg_xy) {
    const float n = __fmul_rn(__fadd_rn(__fmul_rn(__fmul_rn(2.f, mu_x), mu_y), kSsimC1),
                              __fadd_rn(__fmul_rn(2.f, sg_xy), kSsimC2));                  // layers.py:104
    const float d = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(mu_x, mu_x), __fmul_rn(mu_y, mu_y)), kSsimC1),
                              __fadd_rn(__fadd_rn(sg_x, sg_y), kSsimC2));                  // layers.py:105
    const float v = __fmul_rn(__fsub_rn(1.f, div_rn(n, d)), 0.5f);                         // layers.py:106
    return fminf(fmaxf(v, 0.f), 1.f);
}

template <int PITCH>
TDL_DEV void strip_ssim_l1(const float* __restrict__ xs, const float* __restrict__ ys, int r, int c,
                           const float* __restrict__ mu_y, const float* __restrict__ sg_y, float ssim_acc[kR],
                           float l1_acc[kR]) {
    float ax[kR], axx[kR], axy[kR];
#pragma unroll
    for (int j = 0; j < kR + 2; ++j) {
        const float* xr = xs + (r + j) * PITCH + c;
        const float* yr = ys + (r + j) * PITCH + c;
        const float x0 = xr[0], x1 = xr[1], x2 = xr[2];
        const float y0 = yr[0], y1 = yr[1], y2 = yr[2];
        const float q0 = __fmul_rn(x0, x0), q1 = __fmul_rn(x1, x1), q2 = __fmul_rn(x2, x2);
        const float p0 = __fmul_rn(x0, y0), p1 = __fmul_rn(x1, y1), p2 = __fmul_rn(x2, y2);
#pragma unroll
        for (int i = 0; i < kR; ++i) {
            if (i <= j && j <= i + 2) {
                acc_row(ax[i], j == i, x0, x1, x2);
                acc_row(axx[i], j == i, q0, q1, q2);
                acc_row(axy[i], j == i, p0, p1, p2);
            }
        }
        if (j >= 1 && j <= kR) {                                                           // centre pixel of window j-1
            const float df = __fsub_rn(y1, x1);                                            // net.py:57
            l1_acc[j - 1] += sqrt_fast(__fadd_rn(__fmul_rn(df, df), kL1Eps2));
        }
        if (j >= 2) {
            const int i = j - 2;
            const float mu_x = div9(ax[i]);
            const float my = mu_y[i * kNT], sy = sg_y[i * kNT];                            // [i][thread] layout
            const float sg_x = __fsub_rn(div9(axx[i]), __fmul_rn(mu_x, mu_x));
            const float sg_xy = __fsub_rn(div9(axy[i]), __fmul_rn(mu_x, my));
            ssim_acc[i] += ssim_value(mu_x, my, sg_x, sy, sg_xy);
        }
    }
}

// rho = 0.85 * mean_c SSIM + 0.15 * mean_c L1 for the strip (net.py:63-67)
template <int PITCH>
TDL_DEV void strip_reprojection(const float* __restrict__ pred3, const float* __restrict__ tgt3, int plane,
                                int r, int c, const float* __restrict__ stats, float rho[kR]) {
    float sa[kR], la[kR];
#pragma unroll
    for (int i = 0; i < kR; ++i) sa[i] = la[i] = 0.f;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
        strip_ssim_l1<PITCH>(pred3 + ch * plane, tgt3 + ch * plane, r, c, stats + (ch * 2) * kR * kNT,
                             stats + (ch * 2 + 1) * kR * kNT, sa, la);
#pragma unroll
    for (int i = 0; i < kR; ++i)
        rho[i] = __fadd_rn(__fmul_rn(0.85f, div3(sa[i])), __fmul_rn(0.15f, div3(la[i])));
}

// N(0,1) draws for the S identity channels of one pixel at one scale: the caller's tensors when given
// (reference parity), else ONE Philox4x32-10 call -> two Box-Muller pairs -> up to four normals.
template <int S>
TDL_DEV void automask_noise(const PhotoDev& p, int s, int b, size_t pix, size_t HW, float out[S]) {
    if (p.noise[s][0]) {
#pragma unroll
        for (int f = 0; f < S; ++f) out[f] = __ldg(p.noise[s][f] + (size_t)b * HW + pix);
        return;
    }
    const unsigned long long idx = (unsigned long long)b * HW + pix;
    const uint4 r = philox4x32(make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)s, 0u),
                               make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32)));
    const float2 n01 = box_muller(r.x, r.y);
    out[0] = n01.x;
    if (S > 1) out[1] = n01.y;
    if (S > 2) {
        const float2 n23 = box_muller(r.z, r.w);
        out[2] = n23.x;
        if (S > 3) out[S - 1] = n23.y;
    }
}

// ------------------------------------------------------------------------------------------------
template <int S>
__global__ void __launch_bounds__(kNT, 2) photo_fwd_kernel(const PhotoDev p) {
    constexpr int PW = kTW + 2, PH = kTH + 2, PLANE = PW * PH;
    extern __shared__ float smem[];
    float* s_tgt = smem;                  // [3][PH][PW]
    float* s_buf = smem + 3 * PLANE;      // [S][3][PH][PW]: sources (identity), then warped per scale
    __shared__ float s_red[32];
    __shared__ float s_cam[TDL_MAX_SRC * 12 + 9];

    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int b = blockIdx.z, tx0 = blockIdx.x * kTW, ty0 = blockIdx.y * kTH;
    const int H = p.H, W = p.W;
    const size_t HW = (size_t)H * W;
    const ProjConst pcst = make_proj_const(H, W, p.align_corners);      // (IEEE divisions: once per kernel, not per loop)

    if (tid < S * 12) s_cam[tid] = __ldg(p.P + (size_t)b * S * 12 + tid);
    if (tid >= 64 && tid < 64 + 9) s_cam[TDL_MAX_SRC * 12 + tid - 64] = __ldg(p.invK + (size_t)b * 9 + tid - 64);
    const float* s_iK = s_cam + TDL_MAX_SRC * 12;

    // ---- phase 0: target and source tiles with a reflected halo of 1
    {
        const float* tb = p.target + (size_t)b * 3 * HW;
        for (int i = tid; i < PLANE; i += kNT) {
            const int r = i / PW, c = i - r * PW;
            const size_t o = (size_t)reflect1(ty0 - 1 + r, H) * W + reflect1(tx0 - 1 + c, W);
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) s_tgt[ch * PLANE + i] = __ldg(tb + ch * HW + o);
#pragma unroll
            for (int f = 0; f < S; ++f) {
                const float* sb = p.src[f] + (size_t)b * 3 * HW;
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) s_buf[(f * 3 + ch) * PLANE + i] = __ldg(sb + ch * HW + o);
            }
        }
    }
    __syncthreads();

    // ---- area-downsampled target pyramid (F.interpolate(mode='area'), net.py:259) and disparity sums
    for (int s = 0; s < p.nscales; ++s) {
        const int fac = p.fac[s], cells = kTW / fac, h = p.dh[s], w = p.dw[s];
        const int cj0 = ty0 / fac, ci0 = tx0 / fac;
        const float inv = 1.f / (float)(fac * fac);
        float dsum = 0.f;
        for (int i = tid; i < cells * cells * 3; i += kNT) {
            const int ch = i / (cells * cells), rem = i - ch * cells * cells;
            const int cj = rem / cells, ci = rem - cj * cells;
            if (cj0 + cj < h && ci0 + ci < w) {
                const float* base = s_tgt + ch * PLANE + (1 + cj * fac) * PW + 1 + ci * fac;
                float acc = 0.f;
                for (int dy = 0; dy < fac; ++dy)
                    for (int dx = 0; dx < fac; ++dx) acc += base[dy * PW + dx];
                const size_t o = (((size_t)b * 3 + ch) * h + cj0 + cj) * w + ci0 + ci;
                p.J[s][o] = acc * inv;
                if (ch == 0) dsum += __ldg(p.disp[s] + ((size_t)b * h + cj0 + cj) * w + ci0 + ci);
            }
        }
        dsum = block_sum(dsum, s_red);
        if (tid == 0) atomicAdd(p.acc + ((size_t)s * p.B + b) * 4 + 1, (double)dsum);
    }

    // ---- per-thread strip: column `lane`, rows wrp*kR .. wrp*kR+kR-1 of the tile
    const int r0 = wrp * kR;
    const int gx = tx0 + lane;
    float* s_stats = s_buf + S * 3 * PLANE + tid;      // [ch][mu|sigma][i][thread]
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        float mu_y[kR], sg_y[kR];
        strip_target_stats<PW>(s_tgt + ch * PLANE, r0, lane, mu_y, sg_y);
#pragma unroll
        for (int i = 0; i < kR; ++i) {
            s_stats[((ch * 2) * kR + i) * kNT] = mu_y[i];
            s_stats[((ch * 2 + 1) * kR + i) * kNT] = sg_y[i];
        }
    }

    float rho_id[S][kR];
    if (p.automask) {
#pragma unroll
        for (int f = 0; f < S; ++f)
            strip_reprojection<PW>(s_buf + f * 3 * PLANE, s_tgt, PLANE, r0, lane, s_stats, rho_id[f]);
    }
    __syncthreads();

    const DepthParams dp{p.min_disp, p.range};

    for (int s = 0; s < p.nscales; ++s) {
        // ---- phase A: re-project the tile (+halo) through disp_s and sample every source
        {
            const int h = p.dh[s], w = p.dw[s];
            const float* db = p.disp[s] + (size_t)b * h * w;
            for (int i = tid; i < PLANE; i += kNT) {
                const int r = i / PW, c = i - r * PW;
                const int py = reflect1(ty0 - 1 + r, H), px = reflect1(tx0 - 1 + c, W);
                const UpTap ut = up_tap(py, px, p.sy[s], p.sx[s], h, w);
                const Geo g = backproject(up_value(db, w, ut), dp, s_iK, px, py);
#pragma unroll
                for (int f = 0; f < S; ++f) {
                    const Proj pr = project<false>(g, s_cam + f * 12, pcst);
                    const Bilin bt = bilin_taps(pr.ix, pr.iy, H, W);
                    const float* sb = p.src[f] + (size_t)b * 3 * HW;
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) s_buf[(f * 3 + ch) * PLANE + i] = bilin_sample(sb + ch * HW, W, bt);
                }
            }
        }
        __syncthreads();

        // ---- phase B: reprojection errors, automask, minimum over frames
        float best[kR];
        int arg[kR];
#pragma unroll
        for (int i = 0; i < kR; ++i) {
            best[i] = 0.f;
            arg[i] = -1;
        }
        int chan = 0;
        if (p.automask) {
#pragma unroll
            for (int i = 0; i < kR; ++i) {
                const int gy = ty0 + r0 + i;
                float nz[S];
#pragma unroll
                for (int f = 0; f < S; ++f) nz[f] = 0.f;
                if (gx < W && gy < H) automask_noise<S>(p, s, b, (size_t)gy * W + gx, HW, nz);
#pragma unroll
                for (int f = 0; f < S; ++f) {
                    const float v = __fadd_rn(rho_id[f][i], __fmul_rn(nz[f], 1e-5f));     // net.py:94
                    if (f == 0 || v < best[i]) {
                        best[i] = v;
                        arg[i] = f;
                    }
                }
            }
            chan = S;
        }
#pragma unroll
        for (int f = 0; f < S; ++f) {
            float rho[kR];
            strip_reprojection<PW>(s_buf + f * 3 * PLANE, s_tgt, PLANE, r0, lane, s_stats, rho);
#pragma unroll
            for (int i = 0; i < kR; ++i) {
                if (arg[i] < 0 || rho[i] < best[i]) {
                    best[i] = rho[i];
                    arg[i] = chan;
                }
            }
            ++chan;
        }

        // ---- outputs of this scale
        float lsum = 0.f;
#pragma unroll
        for (int i = 0; i < kR; ++i) {
            const int gy = ty0 + r0 + i;
            if (gx < W && gy < H) {
                const size_t pix = (size_t)gy * W + gx;
                lsum += best[i];
                p.argmin[((size_t)s * p.B + b) * HW + pix] = (unsigned char)arg[i];
                if (p.min_index[s]) p.min_index[s][(size_t)b * HW + pix] = arg[i];
#pragma unroll
                for (int f = 0; f < S; ++f) {
                    float* wo = p.warped[s][f];
                    if (wo) {
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch)
                            wo[((size_t)b * 3 + ch) * HW + pix] =
                                s_buf[(f * 3 + ch) * PLANE + (r0 + 1 + i) * PW + lane + 1];
                    }
                }
            }
        }
        lsum = block_sum(lsum, s_red);     // also the barrier that frees s_buf for the next scale
        if (tid == 0) atomicAdd(p.acc + ((size_t)s * p.B + b) * 4 + 0, (double)lsum);
        __syncthreads();
    }
}


// ================================================================================================
// Split forward (used when the warped images are materialised and TMA can describe the tensors):
//   photo_warp_kernel   one thread per pixel and scale: up-sample disp -> depth -> back-project -> project ->
//                       bilinear border sampling of every source; writes outputs[("color", f, s)].  No halo,
//                       no shared memory, high occupancy.
//   photo_score_kernel  one CTA per 32x32 tile, all scales: stages the target, the sources (identity terms)
//                       and, per scale, the S warped tiles with 3-D TMA box copies (halo 1, reflect cells
//                       patched), then runs the SSIM + L1 + automask + min pipeline of photo_fwd_kernel.
// Compared with the fused photo_fwd_kernel the projection is not recomputed for the tile halo, and the
// scoring kernel carries no geometry code (fewer live registers, latency of the gathers hidden elsewhere).
// ================================================================================================
template <int S>
// (min-blocks 1: lets ptxas keep the 24 taps of a scale in flight -- 90 -> 85 us, measured)
__global__ void __launch_bounds__(256, 1) photo_warp_kernel(const PhotoDev p) {
    __shared__ float s_cam[TDL_MAX_SRC * 12 + 9];
    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const int H = p.H, W = p.W;
    const size_t HW = (size_t)H * W;
    const ProjConst pcst = make_proj_const(H, W, p.align_corners);      // (IEEE divisions: once per kernel, not per loop)
    if (tid < S * 12) s_cam[tid] = __ldg(p.P + (size_t)b * S * 12 + tid);
    if (tid >= 64 && tid < 64 + 9) s_cam[TDL_MAX_SRC * 12 + tid - 64] = __ldg(p.invK + (size_t)b * 9 + tid - 64);
    __syncthreads();
    // the scoring / smoothness kernels that follow accumulate their per-image sums into p.acc: cleared here (one launch less)
    if (blockIdx.x == 0 && b == 0)
        for (int i = tid; i < p.acc_n; i += 256) p.acc[i] = 0.0;          // (+ the work-list header)
    const int pix = blockIdx.x * 256 + tid;
    if (pix >= (int)HW) return;
    const int y = pix / W, x = pix - y * W;
    const DepthParams dp{p.min_disp, p.range};
    // all scales of one pixel in one thread: the camera matrices are loaded once and the independent scales
    // give the scheduler several gather rounds to overlap
    float dsp[TDL_MAX_SCALES];
#pragma unroll
    for (int s = 0; s < TDL_MAX_SCALES; ++s) {
        if (s < p.nscales) {
            const int h = p.dh[s], w = p.dw[s];
            const UpTap ut = up_tap(y, x, p.sy[s], p.sx[s], h, w);
            dsp[s] = up_value(p.disp[s] + (size_t)b * h * w, w, ut);
        }
    }
#pragma unroll
    for (int s = 0; s < TDL_MAX_SCALES; ++s) {
        if (s < p.nscales) {
            const Geo g = backproject(dsp[s], dp, s_cam + TDL_MAX_SRC * 12, x, y);
            Bilin bt[S];
            float v[S][3][4];
#pragma unroll
            for (int f = 0; f < S; ++f) {
                const Proj pr = project<false>(g, s_cam + f * 12, pcst);
                bt[f] = bilin_taps(pr.ix, pr.iy, H, W);
                // block-uniform 64-bit base + 32-bit per-thread offsets: one integer add per access
                const float* sb = opaque(p.src[f] + (size_t)b * 3 * HW);
                const unsigned o00 = (unsigned)(bt[f].y0 * W + bt[f].x0);
                const unsigned o01 = o00 + (bt[f].vx ? 1u : 0u), o10 = o00 + (bt[f].vy ? (unsigned)W : 0u);
                const unsigned o11 = o10 + (bt[f].vx ? 1u : 0u);
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const unsigned co = (unsigned)ch * (unsigned)HW;
                    v[f][ch][0] = __ldg(sb + (co + o00));
                    v[f][ch][1] = __ldg(sb + (co + o01));
                    v[f][ch][2] = __ldg(sb + (co + o10));
                    v[f][ch][3] = __ldg(sb + (co + o11));
                }
            }
#pragma unroll
            for (int f = 0; f < S; ++f) {
                float* wo = opaque(p.warped[s][f] + (size_t)b * 3 * HW);
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    float acc = v[f][ch][0] * bt[f].nw;
                    acc += v[f][ch][1] * bt[f].ne;
                    acc += v[f][ch][2] * bt[f].sw;
                    acc += v[f][ch][3] * bt[f].se;
                    wo[(unsigned)ch * (unsigned)HW + (unsigned)pix] = acc;
                }
            }
        }
    }
}

constexpr int kFX0 = 4, kFY0 = 1;                       // tile origin inside the staged forward box
constexpr int kFW = kTW + 8, kFH = kTH + 2, kFPLANE = kFW * kFH;
constexpr int kFGROUP = (3 * kFPLANE + 31) / 32 * 32;   // 3 planes, 128-byte multiple (TMA destination)

// patch the reflect cells (one pixel outside the image) of `ngroups` staged 3-plane groups
TDL_DEV void reflect_fix(float* groups, int ngroups, int tx0, int ty0, int H, int W, int tid) {
    const bool bl = tx0 == 0, br = tx0 + kTW >= W, bt_ = ty0 == 0, bb = ty0 + kTH >= H;
    if (bl || br) {
        for (int e = tid; e < ngroups * 3 * kFH; e += kNT) {
            const int g = e / (3 * kFH), rem = e - g * 3 * kFH, ch = rem / kFH, r = rem - ch * kFH;
            float* row = groups + g * kFGROUP + ch * kFPLANE + r * kFW;
            if (bl) row[kFX0 - 1] = row[kFX0 + 1];
            if (br) row[W - tx0 + kFX0] = row[W - tx0 + kFX0 - 2];
        }
    }
    if (bt_ || bb) {
        __syncthreads();
        for (int e = tid; e < ngroups * 3 * kFW; e += kNT) {
            const int g = e / (3 * kFW), rem = e - g * 3 * kFW, ch = rem / kFW, c = rem - ch * kFW;
            float* col = groups + g * kFGROUP + ch * kFPLANE + c;
            if (bt_) col[(kFY0 - 1) * kFW] = col[(kFY0 + 1) * kFW];
            if (bb) col[(H - ty0 + kFY0) * kFW] = col[(H - ty0 + kFY0 - 2) * kFW];
        }
    }
}

template <int S>
__global__ void __launch_bounds__(kNT, 2) photo_score_kernel(const PhotoDev p, const __grid_constant__ PhotoMaps maps,
                                                           const __grid_constant__ PhotoMaps wmaps) {
    constexpr int FW = kFW, FPLANE = kFPLANE, FGROUP = kFGROUP;
    extern __shared__ __align__(128) float smem_score[];
    // S <= 2: the warped tiles are double-buffered (still two CTAs per SM) -- the TMA copies of scale s+1 fly while scale s
    // is scored, and those of scale 0 travel with the target / source tiles, so the CTA waits for memory once, not 1 + nscales times
    constexpr bool kDB = S <= 2;
    float* s_tgt = smem_score;                 // group 0: target
    float* s_img = smem_score + FGROUP;        // groups 1..S: sources (identity terms), then warped per scale (buffer A)
    float* s_alt = s_img + S * FGROUP + 6 * kR * kNT;      // kDB: S more groups (buffer B), behind the statistics
    __shared__ float s_red[32];
    __shared__ uint64_t s_bar, s_bar2;

    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int b = blockIdx.z, tx0 = blockIdx.x * kTW, ty0 = blockIdx.y * kTH;
    const int H = p.H, W = p.W;
    const size_t HW = (size_t)H * W;
    constexpr uint32_t kGroupBytes = 3 * FPLANE * sizeof(float);

    if (tid == 0) {
        mbar_init(&s_bar, 1);
        mbar_init(&s_bar2, 1);
        mbar_arrive_expect_tx(&s_bar, (1 + S) * kGroupBytes);
        tma_load_3d(s_tgt, &maps.tgt, &s_bar, tx0 - kFX0, ty0 - kFY0, b * 3);
#pragma unroll
        for (int f = 0; f < S; ++f) tma_load_3d(s_img + f * FGROUP, &maps.img[f], &s_bar, tx0 - kFX0, ty0 - kFY0, b * 3);
        if (kDB) {                             // warped tiles of scale 0 -> buffer B
            mbar_arrive_expect_tx(&s_bar2, S * kGroupBytes);
#pragma unroll
            for (int f = 0; f < S; ++f)
                tma_load_3d(s_alt + f * FGROUP, &wmaps.img[f], &s_bar2, tx0 - kFX0, ty0 - kFY0, b * 3);
        }
    }
    __syncthreads();                           // barrier initialisation visible to the waiting threads
    uint32_t parity = 0, parity2 = 0;
    mbar_wait(&s_bar, parity);
    parity ^= 1;
    reflect_fix(smem_score, 1 + S, tx0, ty0, H, W, tid);
    __syncthreads();

    // ---- area-downsampled target pyramid (F.interpolate(mode='area'), net.py:259) and disparity sums; the per-scale
    //      sums stay in registers and are reduced over the CTA once, after the scale loop (no barriers in between)
    float dsum[TDL_MAX_SCALES], lsum[TDL_MAX_SCALES];
#pragma unroll
    for (int s = 0; s < TDL_MAX_SCALES; ++s) dsum[s] = lsum[s] = 0.f;
#pragma unroll
    for (int s = 0; s < TDL_MAX_SCALES; ++s) {
        if (s < p.nscales) {
            const int fac = p.fac[s], cells = kTW / fac, h = p.dh[s], w = p.dw[s];
            const int cj0 = ty0 / fac, ci0 = tx0 / fac;
            const float inv = 1.f / (float)(fac * fac);
            for (int i = tid; i < cells * cells * 3; i += kNT) {
                const int ch = i / (cells * cells), rem = i - ch * cells * cells;
                const int cj = rem / cells, ci = rem - cj * cells;
                if (cj0 + cj < h && ci0 + ci < w) {
                    const float* base = s_tgt + ch * FPLANE + (kFY0 + cj * fac) * FW + kFX0 + ci * fac;
                    float acc = 0.f;
                    for (int dy = 0; dy < fac; ++dy)
                        for (int dx = 0; dx < fac; ++dx) acc += base[dy * FW + dx];
                    const size_t o = (((size_t)b * 3 + ch) * h + cj0 + cj) * w + ci0 + ci;
                    p.J[s][o] = acc * inv;
                    if (ch == 0) dsum[s] += __ldg(p.disp[s] + ((size_t)b * h + cj0 + cj) * w + ci0 + ci);
                }
            }
        }
    }

    // ---- per-thread strip: column `lane`, rows wrp*kR .. of the tile; window corner (r0, lane + kFX0 - 1)
    const int r0 = wrp * kR, c0 = lane + kFX0 - 1;
    const int gx = tx0 + lane;
    float* s_stats = s_img + S * FGROUP + tid;      // [ch][mu|sigma][i][thread]
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        float mu_y[kR], sg_y[kR];
        strip_target_stats<FW>(s_tgt + ch * FPLANE, r0, c0, mu_y, sg_y);
#pragma unroll
        for (int i = 0; i < kR; ++i) {
            s_stats[((ch * 2) * kR + i) * kNT] = mu_y[i];
            s_stats[((ch * 2 + 1) * kR + i) * kNT] = sg_y[i];
        }
    }
    float rho_id[S][kR];
    if (p.automask) {
#pragma unroll
        for (int f = 0; f < S; ++f)
            strip_reprojection<FW>(s_img + f * FGROUP, s_tgt, FPLANE, r0, c0, s_stats, rho_id[f]);
    }
    __syncthreads();

    for (int s = 0; s < p.nscales; ++s) {
        // ---- stage the S warped tiles of this scale (written by photo_warp_kernel)
        float* s_cur = s_img;
        if (kDB) {
            // scale s was requested one step ago (scale 0: with the prologue); request scale s + 1 into the other buffer,
            // which every thread has finished reading (the identity terms, or scale s - 1, ended with a CTA barrier)
            const bool curB = (s & 1) == 0;
            s_cur = curB ? s_alt : s_img;
            if (tid == 0 && s + 1 < p.nscales) {
                float* s_nxt = curB ? s_img : s_alt;
                uint64_t* bar_nxt = curB ? &s_bar : &s_bar2;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive_expect_tx(bar_nxt, S * kGroupBytes);
#pragma unroll
                for (int f = 0; f < S; ++f)
                    tma_load_3d(s_nxt + f * FGROUP, &wmaps.img[(s + 1) * S + f], bar_nxt, tx0 - kFX0, ty0 - kFY0, b * 3);
            }
            if (curB) {
                mbar_wait(&s_bar2, parity2);
                parity2 ^= 1;
            } else {
                mbar_wait(&s_bar, parity);
                parity ^= 1;
            }
        } else {
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive_expect_tx(&s_bar, S * kGroupBytes);
#pragma unroll
                for (int f = 0; f < S; ++f)
                    tma_load_3d(s_img + f * FGROUP, &wmaps.img[s * S + f], &s_bar, tx0 - kFX0, ty0 - kFY0, b * 3);
            }
            mbar_wait(&s_bar, parity);
            parity ^= 1;
        }
        reflect_fix(s_cur, S, tx0, ty0, H, W, tid);
        __syncthreads();

        // ---- reprojection errors, automask, minimum over frames
        float best[kR];
        int arg[kR];
#pragma unroll
        for (int i = 0; i < kR; ++i) {
            best[i] = 0.f;
            arg[i] = -1;
        }
        int chan = 0;
        if (p.automask) {
#pragma unroll
            for (int i = 0; i < kR; ++i) {
                const int gy = ty0 + r0 + i;
                float nz[S];
#pragma unroll
                for (int f = 0; f < S; ++f) nz[f] = 0.f;
                if (gx < W && gy < H) automask_noise<S>(p, s, b, (size_t)gy * W + gx, HW, nz);
#pragma unroll
                for (int f = 0; f < S; ++f) {
                    const float v = __fadd_rn(rho_id[f][i], __fmul_rn(nz[f], 1e-5f));     // net.py:94
                    if (f == 0 || v < best[i]) {
                        best[i] = v;
                        arg[i] = f;
                    }
                }
            }
            chan = S;
        }
#pragma unroll
        for (int f = 0; f < S; ++f) {
            float rho[kR];
            strip_reprojection<FW>(s_cur + f * FGROUP, s_tgt, FPLANE, r0, c0, s_stats, rho);
#pragma unroll
            for (int i = 0; i < kR; ++i) {
                if (arg[i] < 0 || rho[i] < best[i]) {
                    best[i] = rho[i];
                    arg[i] = chan;
                }
            }
            ++chan;
        }
        float ls = 0.f;
#pragma unroll
        for (int i = 0; i < kR; ++i) {
            const int gy = ty0 + r0 + i;
            if (gx < W && gy < H) {
                const size_t pix = (size_t)gy * W + gx;
                ls += best[i];
                p.argmin[((size_t)s * p.B + b) * HW + pix] = (unsigned char)arg[i];
                if (p.min_index[s]) p.min_index[s][(size_t)b * HW + pix] = arg[i];
            }
        }
#pragma unroll
        for (int k = 0; k < TDL_MAX_SCALES; ++k)
            if (k == s) lsum[k] = ls;
        __syncthreads();           // every thread is done with this scale's tiles before a later TMA overwrites them
    }
    // ---- one CTA reduction for the 2 * nscales partial sums: warp shuffles, then warp 0 adds the kNT/32 warp totals
    {
        float* s_part = s_img;                 // [2 * TDL_MAX_SCALES][kNT / 32]; the tiles are no longer needed
#pragma unroll
        for (int k = 0; k < TDL_MAX_SCALES; ++k) {
            const float a = warp_sum(lsum[k]), d = warp_sum(dsum[k]);
            if (lane == 0) {
                s_part[k * (kNT / 32) + wrp] = a;
                s_part[(TDL_MAX_SCALES + k) * (kNT / 32) + wrp] = d;
            }
        }
        __syncthreads();
        if (tid < 2 * TDL_MAX_SCALES) {
            const int k = tid % TDL_MAX_SCALES;
            if (k < p.nscales) {
                float t = 0.f;
#pragma unroll
                for (int j = 0; j < kNT / 32; ++j) t += s_part[tid * (kNT / 32) + j];
                atomicAdd(p.acc + ((size_t)k * p.B + b) * 4 + (tid < TDL_MAX_SCALES ? 0 : 1), (double)t);
            }
        }
    }
}


// SSIM window adjoint: for the 3x3 window centred at xs / ys (pitch PITCH) returns the coefficients (cA, cB, cC) with
//     d (k_in * 9 * v) / d x_q = cA + 2 x_q cB + y_q cC          for every pixel q of the window,
// v = clamp((1 - SSIM)/2, 0, 1) (G_mu, G_a, G_c), zero where torch.clamp blocks the gradient.
// Round 2: the five window sums are accumulated with fused multiply-adds and SSIM is evaluated on the raw 9-sums
// (tdl_ssim.cuh: numerator and denominator scaled by 81^2, no divisions by 9, K1 / K2 only ever added to exact
// products): ~100 instructions per window and channel instead of ~145 for the ATen-ordered sums of round 1, which were
// 36 % of this kernel's 292 M warp-instructions on the scene workload (profiles/r2_*).  The callers pass k = g / 9 (the
// factor of the mean over the window, kept from round 1): with the scaled sums Sx = 9 mu_x ... the chain rule gives
//     d v / d x_q = [ q Sx (d2 - d1) - Sy (n2 - n1) ] / d  +  x_q * 9 q d1 / d  +  y_q * (-9 n1 / d),      q = n / d,
// so cA = 9k (...)/d, 2 cB = 9k * 9 q d1 / d, cC = -9k * 9 n1 / d.
TDL_DEV void window_coefs(float sx, float sy, float sxx, float syy, float sxy, float k, float& cA, float& cB, float& cC) {
    const float n1 = fmaf(sx, sy, fmaf(sx, sy, kK1));
    const float d1 = fmaf(sx, sx, fmaf(sy, sy, kK1));
    const float n2 = fmaf(-sx, sy, fmaf(9.f, sxy, fmaf(9.f, sxy, fmaf(-sx, sy, kK2))));
    const float d2 = fmaf(-sx, sx, fmaf(9.f, sxx, fmaf(9.f, syy, fmaf(-sy, sy, kK2))));
    const float n = __fmul_rn(n1, n2), d = __fmul_rn(d1, d2);
    const float e = __fsub_rn(d, n);                           // v = e / (2 d): the clamp passes gradient for 0 <= v <= 1
    cA = cB = cC = 0.f;
    if (e >= 0.f && e <= 2.f * d) {
        const float rd = rcp_newton(d);
        const float q = n * rd;
        const float k9 = 9.f * k * rd;
        cA = k9 * fmaf(q * sx, d2 - d1, -sy * (n2 - n1));
        cB = 4.5f * k9 * q * d1;
        cC = -9.f * k9 * n1;
    }
}

template <int PITCH>
TDL_DEV void window_adjoint(const float* __restrict__ xs, const float* __restrict__ ys, float k, float& cA, float& cB,
                            float& cC) {
    float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            const float xv = xs[dy * PITCH + dx], yv = ys[dy * PITCH + dx];
            sx += xv;
            sy += yv;
            sxx = fmaf(xv, xv, sxx);
            syy = fmaf(yv, yv, syy);
            sxy = fmaf(xv, yv, sxy);
        }
    }
    window_coefs(sx, sy, sxx, syy, sxy, k, cA, cB, cC);
}

// ------------------------------------------------------------------------------------------------
// Backward.
// Backward tile: halo 2 in y, and in x a left margin of 4 so that the TMA box starts on a 16-byte boundary
// (the innermost box coordinate must keep 16-byte alignment -- tx0 - 2 raised "illegal instruction" on B200).
constexpr int kQX0 = 4, kQY0 = 2;                                     // tile origin inside the staged box
constexpr int kQW = kTW + 8, kQH = kTH + 4, kQPLANE = kQW * kQH;
constexpr int kQGROUP = (3 * kQPLANE + 31) / 32 * 32;                // 3 planes, 128-byte multiple (TMA destination)

template <int S, bool kTMA>
__global__ void __launch_bounds__(kNT, 2) photo_bwd_kernel(const PhotoDev p, const __grid_constant__ PhotoMaps maps) {
    constexpr int QW = kQW, QH = kQH, QPLANE = kQPLANE, QGROUP = kQGROUP;
    constexpr int PW = kTW + 2, PH = kTH + 2;                           // window centres: halo 1
    extern __shared__ __align__(128) float smem[];
    float* s_tgt = smem;                         // [3][QPLANE]
    float* s_wrp = s_tgt + QGROUP;               // [S] groups of [3][QPLANE]
    float* s_coef = s_wrp + S * QGROUP;          // [3 ch][3 coef][QPLANE]
    unsigned char* s_mask = reinterpret_cast<unsigned char*>(s_coef + 9 * QPLANE);   // [QPLANE]
    __shared__ uint64_t s_bar;
    __shared__ float s_cam[TDL_MAX_SRC * 12 + 9];
    __shared__ float s_dP[TDL_MAX_SRC * 12];
    __shared__ int s_cnt[TDL_MAX_SRC];
    __shared__ int s_nlive;             // sparse path: live (frame, pixel) pairs (its own counter: zeroed before the first barrier)
    constexpr int kLCAP = PH * PW;                       // capacity of one frame's window list
    __shared__ unsigned short s_list[S * kLCAP];         // selected windows (cell index), one list per frame
    __shared__ unsigned short s_live[9 * 128];           // sparse path: live (pixel | frame << 11) pairs of <= 128 windows
    __shared__ int s_tx0[kTW], s_tx1[kTW], s_ty0[kTH], s_ty1[kTH];
    __shared__ float s_tlx[kTW], s_tly[kTH];

    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int b = blockIdx.z / p.nscales, s = blockIdx.z - b * p.nscales;
    const int tx0 = blockIdx.x * kTW, ty0 = blockIdx.y * kTH;
    const int H = p.H, W = p.W;
    const size_t HW = (size_t)H * W;
    const ProjConst pcst = make_proj_const(H, W, p.align_corners);      // (IEEE divisions: once per kernel, not per loop)
    const int h = p.dh[s], w = p.dw[s];
    const float up = __ldg(p.dlosses + s) * p.photo_coef[s] / ((float)p.B * (float)H * (float)W);

    // Everything the CTA needs from global memory is requested up front -- the TMA box copies, the arg-min bytes and
    // the disparity taps below do not depend on shared memory -- so one memory latency sits on the critical path instead
    // of three (cam/tables -> barrier -> gathers -> barrier -> tiles): 280 -> 2xx us on the bench shape.
    // (image, scale) pairs whose selected windows fit the work list the scoring kernel emitted are differentiated by
    // photo_bwd_list_kernel: their tiles leave before anything is requested (static scenes: ~1 % of the windows selected)
    if (p.list_max >= 0 && p.lcnt[p.nscales * p.B] == kListMagic && p.lcnt[s * p.B + b] <= p.list_max) return;   // CTA-uniform
    if (kTMA && tid == 0) {
        // one elected thread stages the target and the S warped tiles (halo 2) with 3-D TMA box copies
        mbar_init(&s_bar, 1);
        mbar_arrive_expect_tx(&s_bar, (uint32_t)((1 + S) * 3 * QPLANE * sizeof(float)));
        tma_load_3d(s_tgt, &maps.tgt, &s_bar, tx0 - kQX0, ty0 - kQY0, b * 3);
#pragma unroll
        for (int f = 0; f < S; ++f) tma_load_3d(s_wrp + f * QGROUP, &maps.img[s * S + f], &s_bar, tx0 - kQX0, ty0 - kQY0, b * 3);
    }
    if (tid < S * 12) {
        s_cam[tid] = __ldg(p.P + (size_t)b * S * 12 + tid);
        s_dP[tid] = 0.f;
    }
    // the coefficient planes start from zero (16-byte stores, issued while the TMA copies fly): the dense path then only
    // ever clears the windows of the frame it has finished, and the sparse path finds its accumulators / bit map cleared
    {
        float4* z = reinterpret_cast<float4*>(s_coef);
        for (int i = tid; i < 9 * QPLANE / 4; i += kNT) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        static_assert((9 * QPLANE) % 4 == 0 && QGROUP % 4 == 0, "16-byte stores over the coefficient planes");
    }
    if (tid >= 128 && tid < 128 + TDL_MAX_SRC) s_cnt[tid - 128] = 0;
    if (tid == 254) s_nlive = 0;
    if (tid >= 64 && tid < 64 + 9) s_cam[TDL_MAX_SRC * 12 + tid - 64] = __ldg(p.invK + (size_t)b * 9 + tid - 64);
    const float* s_iK = s_cam + TDL_MAX_SRC * 12;
    if (tid < kTW) up_index(tx0 + tid, p.sx[s], w, s_tx0[tid], s_tx1[tid], s_tlx[tid]);
    if (tid >= 32 && tid < 32 + kTH) up_index(ty0 + tid - 32, p.sy[s], h, s_ty0[tid - 32], s_ty1[tid - 32], s_tly[tid - 32]);
    const DepthParams dp{p.min_disp, p.range};
    const float* db = p.disp[s] + (size_t)b * h * w;
    const unsigned char* am = p.argmin + ((size_t)s * p.B + b) * HW;

    // depth / ray / camera point of this thread's pixels: independent of the source frame AND of the staged
    // tiles, so it is computed here, while the TMA copies issued above are in flight
    const int r0 = wrp * kR;
    const int gx = tx0 + lane;
    // the arg-min bytes of the tile (TMA path) are requested first so that they travel together with the disparity
    // gathers below instead of after them (one global-memory latency less on the CTA's critical path)
    constexpr int NITM = (QPLANE + kNT - 1) / kNT;
    unsigned char mreg[NITM];
    if (kTMA) {
#pragma unroll
        for (int it = 0; it < NITM; ++it) {
            const int i = min(tid + it * kNT, QPLANE - 1);
            const int r = i / QW, c = i - r * QW;
            const int ry = ty0 - 2 + r, rx = tx0 - kQX0 + c;
            const bool inside = ry >= 0 && ry < H && rx >= 0 && rx < W;
            mreg[it] = inside ? am[(size_t)ry * W + rx] : (unsigned char)255;
        }
    }
    float dval[kR];                       // up-sampled disparity of this thread's pixels (taps recomputed here, not read
#pragma unroll                            // from the shared tables, so that the gathers start before the first barrier)
    for (int i = 0; i < kR; ++i) dval[i] = up_value(db, w, up_tap(ty0 + r0 + i, gx, p.sy[s], p.sx[s], h, w));
    __syncthreads();

    // ---- phase 1: target, argmin mask and the warped sources over the tile with halo 2.  When the forward
    //      materialised outputs[("color",f,s)] they are re-read (coalesced, bit-identical to what the forward
    //      scored); otherwise the warp is recomputed.  All loads of a thread are issued before the first
    //      shared-memory store (the kernel is latency-bound here: ~10 loads x 5 cells per thread).
    {
        const float* tb = p.target + (size_t)b * 3 * HW;
        const bool have_warped = p.warped[s][0] != nullptr;
        constexpr int NIT = (QPLANE + kNT - 1) / kNT;
        if (kTMA) {
#pragma unroll
            for (int it = 0; it < NITM; ++it)
                if (tid + it * kNT < QPLANE) s_mask[tid + it * kNT] = mreg[it];
            mbar_wait(&s_bar, 0);
            // nn.ReflectionPad2d(1): the zero-filled cells one pixel outside the image take their mirror value
            const bool bl = tx0 == 0, br = tx0 + kTW >= W, bt_ = ty0 == 0, bb = ty0 + kTH >= H;
            if (bl || br) {
                for (int e = tid; e < (1 + S) * 3 * QH; e += kNT) {
                    const int g = e / (3 * QH), rem = e - g * 3 * QH, ch = rem / QH, r = rem - ch * QH;
                    float* row = smem + g * QGROUP + ch * QPLANE + r * QW;
                    if (bl) row[kQX0 - 1] = row[kQX0 + 1];
                    if (br) row[W - tx0 + kQX0] = row[W - tx0 + kQX0 - 2];
                }
            }
            if (bt_ || bb) {
                __syncthreads();
                for (int e = tid; e < (1 + S) * 3 * QW; e += kNT) {
                    const int g = e / (3 * QW), rem = e - g * 3 * QW, ch = rem / QW, c = rem - ch * QW;
                    float* col = smem + g * QGROUP + ch * QPLANE + c;
                    if (bt_) col[1 * QW] = col[3 * QW];
                    if (bb) col[(H - ty0 + 2) * QW] = col[(H - ty0) * QW];
                }
            }
        } else if (have_warped) {
            float tv[NIT][3], wv[NIT][S][3];
            unsigned char mv[NIT];
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const int i = min(tid + it * kNT, QPLANE - 1);
                const int r = i / QW, c = i - r * QW;
                const int ry = ty0 - 2 + r, rx = tx0 - kQX0 + c;
                const bool inside = ry >= 0 && ry < H && rx >= 0 && rx < W;
                const size_t o = (size_t)reflect1(ry, H) * W + reflect1(rx, W);
                mv[it] = inside ? am[o] : (unsigned char)255;
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) tv[it][ch] = __ldg(tb + ch * HW + o);
#pragma unroll
                for (int f = 0; f < S; ++f) {
                    const float* wb = p.warped[s][f] + (size_t)b * 3 * HW + o;
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) wv[it][f][ch] = __ldg(wb + ch * HW);
                }
            }
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const int i = tid + it * kNT;
                if (i < QPLANE) {
                    s_mask[i] = mv[it];
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) s_tgt[ch * QPLANE + i] = tv[it][ch];
#pragma unroll
                    for (int f = 0; f < S; ++f)
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) s_wrp[f * QGROUP + ch * QPLANE + i] = wv[it][f][ch];
                }
            }
        } else {
            for (int i = tid; i < QPLANE; i += kNT) {
                const int r = i / QW, c = i - r * QW;
                const int ry = ty0 - 2 + r, rx = tx0 - kQX0 + c;
                const bool inside = ry >= 0 && ry < H && rx >= 0 && rx < W;
                const int py = reflect1(ry, H), px = reflect1(rx, W);
                const size_t o = (size_t)py * W + px;
                s_mask[i] = inside ? am[o] : (unsigned char)255;
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) s_tgt[ch * QPLANE + i] = __ldg(tb + ch * HW + o);
                const UpTap ut = up_tap(py, px, p.sy[s], p.sx[s], h, w);
                const Geo g = backproject(up_value(db, w, ut), dp, s_iK, px, py);
#pragma unroll
                for (int f = 0; f < S; ++f) {
                    const Proj pr = project<false>(g, s_cam + f * 12, pcst);
                    const Bilin bt = bilin_taps(pr.ix, pr.iy, H, W);
                    const float* sb = p.src[f] + (size_t)b * 3 * HW;
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch)
                        s_wrp[f * QGROUP + ch * QPLANE + i] = bilin_sample(sb + ch * HW, W, bt);
                }
            }
        }
    }
    __syncthreads();

    float gd[kR];                       // d loss / d (up-sampled disparity) of this thread's pixels
#pragma unroll
    for (int i = 0; i < kR; ++i) gd[i] = 0.f;
    const float g_ssim = up * 0.85f / 3.f, g_l1 = up * 0.15f / 3.f;

    // ---- phase 2a: the windows (tile + halo 1) whose arg-min is a WARPED frame carry gradient; they are compacted into
    //      ONE LIST PER FRAME (window cell; frame k's list starts at k * kLCAP), so that the heavy statistics run with
    //      full warps: a warp of the per-frame loop below then holds 32 windows of the same frame (one mixed list left
    //      half of the lanes idle in each of the S passes wherever the frames mix at pixel level -- the rule on a moving
    //      scene: 236 M -> 207 M warp-instructions on the bench's scene workload).
    const int chan0 = p.automask ? S : 0;
    for (int i = tid; i < PH * PW; i += kNT) {
        const int r = i / PW, c = i - r * PW;
        const int q = (r + 1) * QW + c + kQX0 - 1;
        const int f = (int)s_mask[q] - chan0;
        if (f >= 0 && f < S) s_list[f * kLCAP + atomicAdd(&s_cnt[f], 1)] = (unsigned short)q;
    }
    __syncthreads();
    int n_all = 0;
#pragma unroll
    for (int k = 0; k < S; ++k) n_all += s_cnt[k];
    if (n_all == 0) return;                       // automasking: nothing selected in this tile (CTA-uniform)

    if (n_all <= p.sparse_max) {
        // ---- sparse tile (auto-masked / static regions: a few selected windows).  The dense path below would spend its
        //      time on box sums of zeros and on strips with one live pixel; here the selected windows SCATTER their adjoint
        //      into per-frame tile accumulators, the (frame, pixel) pairs that received something are compacted, and one
        //      thread per live pair runs the sampling / projection chain.  Same arithmetic, different summation order.
        constexpr int TP = kTH * kTW;
        float* s_G = s_coef;                                  // [S][3][TP]
        unsigned* s_bits = reinterpret_cast<unsigned*>(s_coef + S * 3 * TP);   // [S][TP / 32]: pixel already in the live list
        static_assert(S * 3 * TP + S * TP / 32 <= 9 * QPLANE, "accumulators + bit map inside the (pre-zeroed) coefficient planes");
        for (int e = tid; e < 3 * n_all; e += kNT) {          // one thread per (selected window, channel)
            int wi = e / 3, f = 0;
            const int ch = e - wi * 3;
#pragma unroll
            for (int k = 0; k < S - 1; ++k)
                if (f == k && wi >= s_cnt[k]) {
                    wi -= s_cnt[k];
                    f = k + 1;
                }
            const int q = s_list[f * kLCAP + wi];
            const int r = q / QW, c = q - r * QW;
            const int wy = ty0 - kQY0 + r, wx = tx0 - kQX0 + c;           // image coordinates of the window centre
            const float* xs = s_wrp + f * QGROUP + ch * QPLANE + q;
            const float* ys = s_tgt + ch * QPLANE + q;
            float* Gf = s_G + (f * 3 + ch) * TP;
            float cA, cB, cC;
            window_adjoint<QW>(xs, ys, g_ssim * (1.f / 9.f), cA, cB, cC);
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy) {
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    // nn.ReflectionPad2d(1): a tap outside the image IS its mirror pixel (the staged cell holds the
                    // mirror value), which gives the border multiplicities of the dense path for free
                    const int ly = reflect1(wy + dy, H) - ty0, lx = reflect1(wx + dx, W) - tx0;
                    if (ly >= 0 && ly < kTH && lx >= 0 && lx < kTW) {
                        const int px = ly * kTW + lx;
                        atomicAdd(&Gf[px], cA + 2.f * xs[dy * QW + dx] * cB + ys[dy * QW + dx] * cC);
                        // the channel-0 thread of the window enters the pixel into the live list, once per (frame, pixel)
                        if (ch == 0 && tx0 + lx < W && ty0 + ly < H &&
                            !(atomicOr(&s_bits[f * (TP / 32) + (px >> 5)], 1u << (px & 31)) & (1u << (px & 31))))
                            s_live[atomicAdd(&s_nlive, 1)] = (unsigned short)(px | (f << 11));
                    }
                }
            }
            const int ly = wy - ty0, lx = wx - tx0;
            if (ly >= 0 && ly < kTH && lx >= 0 && lx < kTW) {                 // robust-L1 term of the centre pixel
                const float df = xs[0] - ys[0];
                atomicAdd(&Gf[ly * kTW + lx], g_l1 * df * rsqrt_approx(df * df + kL1Eps2));
            }
        }
        __syncthreads();
        const int n_act = s_nlive;
        float* dd = p.d_disp[s] + (size_t)b * h * w;
        for (int e0 = wrp * 32; e0 < n_act; e0 += kNT) {      // warp-uniform trip count: the dP reduction below is warp-wide
            const int e = e0 + lane;
            const bool on = e < n_act;
            const int ent = on ? s_live[e] : 0, i = ent & 2047, f = ent >> 11;
            const int ly = i / kTW, lx = i - ly * kTW;
            const int py = ty0 + ly, px = tx0 + lx;
            const float* Pf = s_cam + f * 12;
            const float* sbase = p.src[0];
#pragma unroll
            for (int k = 1; k < S; ++k)
                if (k == f) sbase = p.src[k];
            sbase += (size_t)b * 3 * HW;
            const UpTap ut = up_tap(py, px, p.sy[s], p.sx[s], h, w);
            const Geo g = backproject(up_value(db, w, ut), dp, s_iK, px, py);
            const Proj pr = project<true>(g, Pf, pcst);
            const Bilin bt = bilin_taps(pr.ix, pr.iy, H, W);
            const float* q = sbase + (size_t)bt.y0 * W + bt.x0;
            const int dx = bt.vx ? 1 : 0, dy = bt.vy ? W : 0;
            float gix = 0.f, giy = 0.f;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const float v00 = __ldg(q + ch * HW), v01 = bt.vx ? __ldg(q + ch * HW + dx) : 0.f;
                const float v10 = bt.vy ? __ldg(q + ch * HW + dy) : 0.f;
                const float v11 = (bt.vx && bt.vy) ? __ldg(q + ch * HW + dy + dx) : 0.f;
                const float dix = -v00 * bt.ey + v01 * bt.ey - v10 * bt.ay + v11 * bt.ay;
                const float diy = -v00 * bt.ex - v01 * bt.ax + v10 * bt.ex + v11 * bt.ax;
                const float Gc = on ? s_G[(f * 3 + ch) * TP + i] : 0.f;
                gix += Gc * dix;
                giy += Gc * diy;
            }
            const float gu = gix * pr.mx, gv = giy * pr.my;
            const float rz = rcp_newton(pr.z);
            const float gp0 = gu * rz, gp1 = gv * rz, gp2 = -(gu * pr.u + gv * pr.v) * rz;
            float aP[12];
            aP[0] = gp0 * g.X0; aP[1] = gp0 * g.X1; aP[2] = gp0 * g.X2; aP[3] = gp0;
            aP[4] = gp1 * g.X0; aP[5] = gp1 * g.X1; aP[6] = gp1 * g.X2; aP[7] = gp1;
            aP[8] = gp2 * g.X0; aP[9] = gp2 * g.X1; aP[10] = gp2 * g.X2; aP[11] = gp2;
#pragma unroll 1
            for (int ff = 0; ff < S; ++ff) {                  // 12-value warp reduction per frame, one shared atomic per value
                float a2[12];
#pragma unroll
                for (int k = 0; k < 12; ++k) a2[k] = (on && f == ff) ? aP[k] : 0.f;
                float tot;
                const int slot = warp_sum12(a2, tot);
                if (slot >= 0 && tot != 0.f) atomicAdd(&s_dP[ff * 12 + slot], tot);
            }
            if (on && (gp0 != 0.f || gp1 != 0.f || gp2 != 0.f)) {
                const float gX0 = Pf[0] * gp0 + Pf[4] * gp1 + Pf[8] * gp2;
                const float gX1 = Pf[1] * gp0 + Pf[5] * gp1 + Pf[9] * gp2;
                const float gX2 = Pf[2] * gp0 + Pf[6] * gp1 + Pf[10] * gp2;
                const float gD = gX0 * g.r0 + gX1 * g.r1 + gX2 * g.r2;
                const float gdisp = -p.range * g.D * g.D * gD;
                const float hy = 1.f - ut.ly, hx = 1.f - ut.lx;   // adjoint of the bilinear up-sampling, straight to d_disp_s
                atomicAdd(dd + (size_t)ut.y0 * w + ut.x0, hy * hx * gdisp);
                atomicAdd(dd + (size_t)ut.y0 * w + ut.x1, hy * ut.lx * gdisp);
                atomicAdd(dd + (size_t)ut.y1 * w + ut.x0, ut.ly * hx * gdisp);
                atomicAdd(dd + (size_t)ut.y1 * w + ut.x1, ut.ly * ut.lx * gdisp);
            }
        }
        __syncthreads();
        if (tid < S * 12 && s_dP[tid] != 0.f) atomicAdd(p.dP + ((size_t)b * S) * 12 + tid, s_dP[tid]);
        return;
    }

    // ---- dense tile: depth / ray / camera point of this thread's strip, then per frame the coefficient planes
    Geo geo[kR];
#pragma unroll
    for (int i = 0; i < kR; ++i) geo[i] = backproject(dval[i], dp, s_iK, gx, ty0 + r0 + i);

    int prev_f = -1;
#pragma unroll 1
    for (int f = 0; f < S; ++f) {
        const int chan = chan0 + f;
        const int n_f = s_cnt[f];                 // windows that selected this frame
        if (n_f == 0) continue;                   // CTA-uniform
        const unsigned short* lst = s_list + f * kLCAP;
        // ---- phase 2b: SSIM adjoint coefficients of the windows that selected this frame; every other cell of the planes
        //      is zero: they start zeroed and the windows of the previous frame (other cells than this frame's: a window
        //      selects one frame) are cleared here through that frame's list
        if (prev_f >= 0) {                        // CTA-uniform
            const unsigned short* pl = s_list + prev_f * kLCAP;
            const int n_p = s_cnt[prev_f];
            for (int e = tid; e < n_p; e += kNT) {
                const int q = pl[e];
#pragma unroll
                for (int k = 0; k < 9; ++k) s_coef[k * QPLANE + q] = 0.f;
            }
        }
        prev_f = f;
        for (int e = tid; e < n_f; e += kNT) {
            const int q = lst[e];
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                float cA, cB, cC;
                window_adjoint<QW>(s_wrp + f * QGROUP + ch * QPLANE + q, s_tgt + ch * QPLANE + q,
                                   g_ssim * (1.f / 9.f), cA, cB, cC);
                s_coef[(ch * 3 + 0) * QPLANE + q] = cA;
                s_coef[(ch * 3 + 1) * QPLANE + q] = cB;
                s_coef[(ch * 3 + 2) * QPLANE + q] = cC;
            }
        }
        __syncthreads();

        // ---- phase 3: gather the window adjoint (3x3 box sums of the coefficient planes, shared along the
        //      thread's vertical strip), then chain through sampling and projection for the pixels that
        //      actually receive gradient from this source frame.
        float G[3][kR];
        {
            const int qc = lane + kQX0;                               // Q column of this thread's pixels
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                float box[3][kR];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float* pl = s_coef + (ch * 3 + k) * QPLANE + (r0 + 1) * QW + qc;
                    float hsum[kR + 2];
#pragma unroll
                    for (int j = 0; j < kR + 2; ++j) hsum[j] = pl[j * QW - 1] + pl[j * QW] + pl[j * QW + 1];
#pragma unroll
                    for (int i = 0; i < kR; ++i) box[k][i] = hsum[i] + hsum[i + 1] + hsum[i + 2];
                }
#pragma unroll
                for (int i = 0; i < kR; ++i) {
                    const int gy = ty0 + r0 + i;
                    const int q = (r0 + i + 2) * QW + qc;
                    // reflect padding: the border windows see their inner neighbour twice (multiplicity 2, 4 in
                    // the corners) -- add the extra copies for the pixels next to the image border
                    const bool bx0 = gx == 1, bx1 = gx == W - 2, by0 = gy == 1, by1 = gy == H - 2;
                    if (bx0 || bx1 || by0 || by1) {
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            const float* c = s_coef + (ch * 3 + k) * QPLANE + q;
                            float e = 0.f;
                            if (bx0) e += c[-QW - 1] + c[-1] + c[QW - 1];
                            if (bx1) e += c[-QW + 1] + c[1] + c[QW + 1];
                            if (by0) e += c[-QW - 1] + c[-QW] + c[-QW + 1];
                            if (by1) e += c[QW - 1] + c[QW] + c[QW + 1];
                            if (bx0 && by0) e += c[-QW - 1];
                            if (bx1 && by0) e += c[-QW + 1];
                            if (bx0 && by1) e += c[QW - 1];
                            if (bx1 && by1) e += c[QW + 1];
                            box[k][i] += e;
                        }
                    }
                    const float xv = s_wrp[f * QGROUP + ch * QPLANE + q], yv = s_tgt[ch * QPLANE + q];
                    float g = box[0][i] + 2.f * xv * box[1][i] + yv * box[2][i];
                    if (s_mask[q] == chan) {
                        const float df = xv - yv;
                        g += g_l1 * df * rsqrt_approx(df * df + kL1Eps2);
                    }
                    G[ch][i] = g;
                }
            }
        }
        float aP[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) aP[k] = 0.f;
        const float* Pf = s_cam + f * 12;
        const float* sbase = opaque(p.src[f] + (size_t)b * 3 * HW);     // (+ 32-bit tap offsets: one multiply-add per address)
#pragma unroll
        for (int i0 = 0; i0 < kR; i0 += 2) {
            bool act[2];
            Proj pr[2];
            Bilin bt[2];
            float v[2][3][4];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int i = i0 + u, gy = ty0 + r0 + i;
                act[u] = gx < W && gy < H && (G[0][i] != 0.f || G[1][i] != 0.f || G[2][i] != 0.f);
                pr[u] = project<true>(geo[i], Pf, pcst);
                bt[u] = bilin_taps(pr[u].ix, pr[u].iy, H, W);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const unsigned o00 = (unsigned)(bt[u].y0 * W + bt[u].x0);
                const unsigned o01 = o00 + (bt[u].vx ? 1u : 0u), o10 = o00 + (bt[u].vy ? (unsigned)W : 0u);
                const unsigned o11 = o10 + (bt[u].vx ? 1u : 0u);
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const unsigned co = (unsigned)ch * (unsigned)HW;
                    v[u][ch][0] = act[u] ? __ldg(sbase + (co + o00)) : 0.f;
                    v[u][ch][1] = act[u] ? __ldg(sbase + (co + o01)) : 0.f;
                    v[u][ch][2] = act[u] ? __ldg(sbase + (co + o10)) : 0.f;
                    v[u][ch][3] = act[u] ? __ldg(sbase + (co + o11)) : 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (act[u]) {
                    const int i = i0 + u;
                    const Geo& g = geo[i];
                    float gix = 0.f, giy = 0.f;
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        // ATen grid_sampler_2d_backward with the out-of-range taps skipped (weight-0 taps)
                        const float v00 = v[u][ch][0], v01 = bt[u].vx ? v[u][ch][1] : 0.f;
                        const float v10 = bt[u].vy ? v[u][ch][2] : 0.f;
                        const float v11 = (bt[u].vx && bt[u].vy) ? v[u][ch][3] : 0.f;
                        const float dix = -v00 * bt[u].ey + v01 * bt[u].ey - v10 * bt[u].ay + v11 * bt[u].ay;
                        const float diy = -v00 * bt[u].ex - v01 * bt[u].ax + v10 * bt[u].ex + v11 * bt[u].ax;
                        gix += G[ch][i] * dix;
                        giy += G[ch][i] * diy;
                    }
                    const float gu = gix * pr[u].mx, gv = giy * pr[u].my;
                    const float rz = rcp_newton(pr[u].z);
                    const float gp0 = gu * rz, gp1 = gv * rz, gp2 = -(gu * pr[u].u + gv * pr[u].v) * rz;
                    aP[0] += gp0 * g.X0; aP[1] += gp0 * g.X1; aP[2] += gp0 * g.X2; aP[3] += gp0;
                    aP[4] += gp1 * g.X0; aP[5] += gp1 * g.X1; aP[6] += gp1 * g.X2; aP[7] += gp1;
                    aP[8] += gp2 * g.X0; aP[9] += gp2 * g.X1; aP[10] += gp2 * g.X2; aP[11] += gp2;
                    const float gX0 = Pf[0] * gp0 + Pf[4] * gp1 + Pf[8] * gp2;
                    const float gX1 = Pf[1] * gp0 + Pf[5] * gp1 + Pf[9] * gp2;
                    const float gX2 = Pf[2] * gp0 + Pf[6] * gp1 + Pf[10] * gp2;
                    const float gD = gX0 * g.r0 + gX1 * g.r1 + gX2 * g.r2;
                    gd[i] += -p.range * g.D * g.D * gD;
                }
            }
        }
        {
            float tot;
            const int slot = warp_sum12(aP, tot);
            if (slot >= 0) atomicAdd(&s_dP[f * 12 + slot], tot);
        }
        __syncthreads();          // s_coef is rewritten by the next source frame
    }

    // ---- phase 4: adjoint of the bilinear up-sampling of disp_s, folded inside the tile
    float* s_g = s_coef;                        // [kTH][kTW]
    float* s_t = s_coef + kTH * kTW;            // [kTH][kTW + 2]
    if (tid < S * 12 && s_dP[tid] != 0.f) atomicAdd(p.dP + ((size_t)b * S) * 12 + tid, s_dP[tid]);
#pragma unroll
    for (int i = 0; i < kR; ++i) s_g[(r0 + i) * kTW + lane] = gd[i];
    __syncthreads();
    const int ilo = s_tx0[0], ihi = s_tx1[kTW - 1], ni = ihi - ilo + 1;       // <= kTW + 1
    const int jlo = s_ty0[0], jhi = s_ty1[kTH - 1], nj = jhi - jlo + 1;
    constexpr int TP = kTW + 2;
    for (int e = tid; e < kTH * ni; e += kNT) {
        const int y = e / ni, ii = e - y * ni, gi = ilo + ii;
        float acc = 0.f;
        const int fac = p.fac[s];
        const int xlo = max(0, (gi - 1) * fac + fac / 2 - 1 - tx0), xhi = min(kTW - 1, (gi + 1) * fac + fac / 2 - tx0);
        for (int x = xlo; x <= xhi; ++x) {
            const float gval = s_g[y * kTW + x];
            if (s_tx0[x] == gi) acc += (1.f - s_tlx[x]) * gval;
            if (s_tx1[x] == gi) acc += s_tlx[x] * gval;
        }
        s_t[y * TP + ii] = acc;
    }
    __syncthreads();
    float* dd = p.d_disp[s] + (size_t)b * h * w;
    for (int e = tid; e < nj * ni; e += kNT) {
        const int jj = e / ni, ii = e - jj * ni, gj = jlo + jj;
        float acc = 0.f;
        const int fac = p.fac[s];
        const int ylo = max(0, (gj - 1) * fac + fac / 2 - 1 - ty0), yhi = min(kTH - 1, (gj + 1) * fac + fac / 2 - ty0);
        for (int y = ylo; y <= yhi; ++y) {
            const float tv = s_t[y * TP + ii];
            if (s_ty0[y] == gj) acc += (1.f - s_tly[y]) * tv;
            if (s_ty1[y] == gj) acc += s_tly[y] * tv;
        }
        if (acc != 0.f) atomicAdd(dd + (size_t)gj * w + ilo + ii, acc);
    }
}

// ------------------------------------------------------------------------------------------------
template <int S>
static cudaError_t launch_fwd_fused(const PhotoDev& p, cudaStream_t st) {
    constexpr int PLANE = (kTW + 2) * (kTH + 2);
    const size_t smem = (size_t)(3 + 3 * S) * PLANE * sizeof(float) + (size_t)6 * kR * kNT * sizeof(float);
    static SmemOptIn opt_in;
    if (cudaError_t e = opt_in(photo_fwd_kernel<S>, smem)) return e;
    dim3 grid((p.W + kTW - 1) / kTW, (p.H + kTH - 1) / kTH, p.B);
    photo_fwd_kernel<S><<<grid, kNT, smem, st>>>(p);
    return cudaGetLastError();
}

template <int S>
static bool encode_fwd_maps(const PhotoDev& p, PhotoMaps* maps, PhotoMaps* wmaps) {
    // split path: every warped image is materialised and TMA can describe target / sources / warps
    bool ok = p.use_tma && p.split_fwd && encode_image_map(&maps->tgt, p.target, p.B * 3, p.H, p.W, kFW, kFH, 3);
    for (int f = 0; ok && f < S; ++f) ok = encode_image_map(&maps->img[f], p.src[f], p.B * 3, p.H, p.W, kFW, kFH, 3);
    for (int s = 0; ok && s < p.nscales; ++s)
        for (int f = 0; ok && f < S; ++f)
            ok = p.warped[s][f] && encode_image_map(&wmaps->img[s * S + f], p.warped[s][f], p.B * 3, p.H, p.W, kFW, kFH, 3);
    return ok;
}

template <int S>
static cudaError_t launch_warp_t(const PhotoDev& p, cudaStream_t st) {
    dim3 wgrid((unsigned)(((size_t)p.H * p.W + 255) / 256), p.B);
    photo_warp_kernel<S><<<wgrid, 256, 0, st>>>(p);
    return cudaGetLastError();
}

template <int S>
static cudaError_t launch_score_t(const PhotoDev& p, cudaStream_t st) {
    PhotoMaps maps, wmaps;
    if (!encode_fwd_maps<S>(p, &maps, &wmaps)) return cudaErrorInvalidValue;
    const size_t smem = (size_t)(1 + S + (S <= 2 ? S : 0)) * kFGROUP * sizeof(float) + (size_t)6 * kR * kNT * sizeof(float);
    static SmemOptIn opt_in;
    if (cudaError_t e = opt_in(photo_score_kernel<S>, smem)) return e;
    dim3 grid((p.W + kTW - 1) / kTW, (p.H + kTH - 1) / kTH, p.B);
    photo_score_kernel<S><<<grid, kNT, smem, st>>>(p, maps, wmaps);
    return cudaGetLastError();
}

#define TDL_DISPATCH_S(fn, ...)                  \
    switch (p.S) {                               \
        case 1: return fn<1>(__VA_ARGS__);       \
        case 2: return fn<2>(__VA_ARGS__);       \
        case 3: return fn<3>(__VA_ARGS__);       \
        case 4: return fn<4>(__VA_ARGS__);       \
    }

bool photo_fwd_can_split(const PhotoDev& p) {
    PhotoMaps maps, wmaps;
    switch (p.S) {
        case 1: return encode_fwd_maps<1>(p, &maps, &wmaps);
        case 2: return encode_fwd_maps<2>(p, &maps, &wmaps);
        case 3: return encode_fwd_maps<3>(p, &maps, &wmaps);
        case 4: return encode_fwd_maps<4>(p, &maps, &wmaps);
    }
    return false;
}

cudaError_t launch_photo_warp(const PhotoDev& p, cudaStream_t st) {
    TDL_DISPATCH_S(launch_warp_t, p, st)
    return cudaErrorInvalidValue;
}

cudaError_t launch_photo_score(const PhotoDev& p, cudaStream_t st) {
    TDL_DISPATCH_S(launch_score_t, p, st)
    return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------------------------------------
// Work-list backward.  For an (image, scale) pair in which auto-masking left only a few windows selected (a static scene:
// ~1 % of the pixels), staging every tile to find a dozen windows each is what costs -- photo_bwd_kernel spent 110-130 us
// on a batch of such images.  photo_score2_kernel appends every selected window (pixel | frame << 28) to a per-(scale,
// image) list while it writes the arg-min; here NINE THREADS per list entry work straight from global memory.  Read as
// (channel, window row) each loads 3 taps of the warped and the target image into shared memory (27 x 2 values per
// window, once); one thread per channel forms the window sums and the SSIM adjoint coefficients; read as TAP each thread then
// takes d loss / d warped value of its pixel and runs the sampling and projection adjoint for it -- the chain is linear
// in the incoming gradient, so the contributions of overlapping windows need not be combined first.  Taps outside the
// image are their mirror pixel (nn.ReflectionPad2d(1)).  4 atomics per thread into d_disp_s; the pose adjoint leaves
// through a warp reduction per frame + shared-memory atomics + one global atomic per CTA.  The kernel is a chain of three
// dependent memory latencies with ~800 instructions in between, so what matters is threads in flight (<= 64 registers).
// Pairs with more than list_max windows are left to photo_bwd_kernel, which in turn skips the listed ones.
constexpr int kListWin = 28;             // windows per CTA sweep: 9 threads each (252 of 256 threads)
constexpr int kListCtas = 74;            // CTAs per (scale, image): 2072 windows per sweep of the grid

template <int S>
__global__ void __launch_bounds__(256, 4) photo_bwd_list_kernel(const PhotoDev p) {
    __shared__ float s_cam[TDL_MAX_SRC * 12 + 9];
    __shared__ float s_dP[TDL_MAX_SRC * 12];
    __shared__ float2 s_xy[kListWin][3][9];          // (warped, target) value of every tap of the CTA's windows
    __shared__ float s_cf[kListWin][3][3];           // per channel: cA, cB, cC (window_coefs)
    const int tid = threadIdx.x;
    const int bs = blockIdx.y, s = bs / p.B, b = bs - s * p.B;
    const int n = p.lcnt[bs];
    if (p.lcnt[p.nscales * p.B] != kListMagic || n > p.list_max || (int)(blockIdx.x * kListWin) >= n) return;   // CTA-uniform
    const int H = p.H, W = p.W, h = p.dh[s], w = p.dw[s];
    const size_t HW = (size_t)H * W;
    const ProjConst pcst = make_proj_const(H, W, p.align_corners);
    if (tid < S * 12) {
        s_cam[tid] = __ldg(p.P + (size_t)b * S * 12 + tid);
        s_dP[tid] = 0.f;
    }
    if (tid >= 64 && tid < 64 + 9) s_cam[TDL_MAX_SRC * 12 + tid - 64] = __ldg(p.invK + (size_t)b * 9 + tid - 64);
    const float* s_iK = s_cam + TDL_MAX_SRC * 12;
    const DepthParams dp{p.min_disp, p.range};
    const float up = __ldg(p.dlosses + s) * p.photo_coef[s] / ((float)p.B * (float)H * (float)W);
    const float g_ssim = up * 0.85f / 3.f, g_l1 = up * 0.15f / 3.f;
    const float* db = p.disp[s] + (size_t)b * h * w;
    float* dd = p.d_disp[s] + (size_t)b * h * w;
    const uint32_t* wl = p.wlist + (size_t)bs * kListCap;
    const float* tb = p.target + (size_t)b * 3 * HW;
    // thread = (window j of the sweep, r9): r9 is the TAP whose pixel the thread differentiates, and -- read as
    // (channel, window row) -- the part of the window sums it contributes
    const int j = tid / 9, r9 = tid - 9 * j;
    const int ch_a = r9 / 3, row_a = r9 - 3 * ch_a;

    for (int e0 = blockIdx.x * kListWin; e0 < n; e0 += gridDim.x * kListWin) {             // CTA-uniform trip count
        const bool on = j < kListWin && e0 + j < n;
        const uint32_t ent = on ? __ldg(wl + e0 + j) : 0u;
        const int f = (int)(ent >> 28), pix = (int)(ent & 0x0fffffffu);
        const int wy = pix / W, wx = pix - wy * W;
        const float* wbase = p.warped[s][0];
        const float* sbase = p.src[0];
#pragma unroll
        for (int k = 1; k < S; ++k)
            if (k == f) {
                wbase = p.warped[s][k];
                sbase = p.src[k];
            }
        wbase += (size_t)b * 3 * HW;
        sbase += (size_t)b * 3 * HW;
        // taps outside the image are their mirror pixel (nn.ReflectionPad2d(1))
        const int ty = r9 / 3, tx = r9 - 3 * ty;
        const int py = reflect1(wy + ty - 1, H), px = reflect1(wx + tx - 1, W);             // this thread's pixel
        // (the disparity taps of the pixel do not depend on the window values: requested first)
        const UpTap ut = up_tap(py, px, p.sy[s], p.sx[s], h, w);
        const float dv = up_value(db, w, ut);
        if (on) {                                                // one window row of one channel: 3 taps of warped + target
            const size_t ro = (size_t)ch_a * HW + (size_t)reflect1(wy + row_a - 1, H) * W;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int cx = reflect1(wx + c - 1, W);
                s_xy[j][ch_a][row_a * 3 + c] = make_float2(__ldg(wbase + ro + cx), __ldg(tb + ro + cx));
            }
        }
        __syncthreads();                                         // (also: s_cam / s_dP of the prologue)
        if (on && r9 < 3) {                                      // one thread per (window, channel): the coefficients, with
            float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;      // the sums in the tap order of window_adjoint
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const float2 v = s_xy[j][r9][k];
                sx += v.x;
                sy += v.y;
                sxx = fmaf(v.x, v.x, sxx);
                syy = fmaf(v.y, v.y, syy);
                sxy = fmaf(v.x, v.y, sxy);
            }
            float cA, cB, cC;
            window_coefs(sx, sy, sxx, syy, sxy, g_ssim * (1.f / 9.f), cA, cB, cC);
            s_cf[j][r9][0] = cA;
            s_cf[j][r9][1] = cB;
            s_cf[j][r9][2] = cC;
        }
        __syncthreads();
        float G[3] = {0.f, 0.f, 0.f};
        if (on) {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const float2 v = s_xy[j][ch][r9];
                G[ch] = s_cf[j][ch][0] + 2.f * v.x * s_cf[j][ch][1] + v.y * s_cf[j][ch][2];
                if (r9 == 4) {                                                              // robust-L1 term of the centre
                    const float df = v.x - v.y;
                    G[ch] += g_l1 * df * rsqrt_approx(df * df + kL1Eps2);
                }
            }
        }
        const float* Pf = s_cam + f * 12;
        const Geo g = backproject(dv, dp, s_iK, px, py);
        const Proj pr = project<true>(g, Pf, pcst);
        const Bilin bt = bilin_taps(pr.ix, pr.iy, H, W);
        const float* q = sbase + (size_t)bt.y0 * W + bt.x0;
        const int dx = bt.vx ? 1 : 0, dy = bt.vy ? W : 0;
        float gix = 0.f, giy = 0.f;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const float v00 = __ldg(q + ch * HW), v01 = bt.vx ? __ldg(q + ch * HW + dx) : 0.f;
            const float v10 = bt.vy ? __ldg(q + ch * HW + dy) : 0.f;
            const float v11 = (bt.vx && bt.vy) ? __ldg(q + ch * HW + dy + dx) : 0.f;
            const float dix = -v00 * bt.ey + v01 * bt.ey - v10 * bt.ay + v11 * bt.ay;
            const float diy = -v00 * bt.ex - v01 * bt.ax + v10 * bt.ex + v11 * bt.ax;
            gix += G[ch] * dix;
            giy += G[ch] * diy;
        }
        const float gu = gix * pr.mx, gv = giy * pr.my;
        const float rz = rcp_newton(pr.z);
        const float gp0 = on ? gu * rz : 0.f, gp1 = on ? gv * rz : 0.f, gp2 = on ? -(gu * pr.u + gv * pr.v) * rz : 0.f;
        float aP[12];
        aP[0] = gp0 * g.X0; aP[1] = gp0 * g.X1; aP[2] = gp0 * g.X2; aP[3] = gp0;
        aP[4] = gp1 * g.X0; aP[5] = gp1 * g.X1; aP[6] = gp1 * g.X2; aP[7] = gp1;
        aP[8] = gp2 * g.X0; aP[9] = gp2 * g.X1; aP[10] = gp2 * g.X2; aP[11] = gp2;
#pragma unroll 1
        for (int ff = 0; ff < S; ++ff) {                      // 12-value warp reduction per frame, one shared atomic per value
            float a2[12];
#pragma unroll
            for (int k = 0; k < 12; ++k) a2[k] = (f == ff) ? aP[k] : 0.f;
            float tot;
            const int slot = warp_sum12(a2, tot);
            if (slot >= 0 && tot != 0.f) atomicAdd(&s_dP[ff * 12 + slot], tot);
        }
        if (gp0 != 0.f || gp1 != 0.f || gp2 != 0.f) {
            const float gX0 = Pf[0] * gp0 + Pf[4] * gp1 + Pf[8] * gp2;
            const float gX1 = Pf[1] * gp0 + Pf[5] * gp1 + Pf[9] * gp2;
            const float gX2 = Pf[2] * gp0 + Pf[6] * gp1 + Pf[10] * gp2;
            const float gD = gX0 * g.r0 + gX1 * g.r1 + gX2 * g.r2;
            const float gdisp = -p.range * g.D * g.D * gD;
            const float hy = 1.f - ut.ly, hx = 1.f - ut.lx;           // adjoint of the bilinear up-sampling, straight to d_disp_s
            atomicAdd(dd + (size_t)ut.y0 * w + ut.x0, hy * hx * gdisp);
            atomicAdd(dd + (size_t)ut.y0 * w + ut.x1, hy * ut.lx * gdisp);
            atomicAdd(dd + (size_t)ut.y1 * w + ut.x0, ut.ly * hx * gdisp);
            atomicAdd(dd + (size_t)ut.y1 * w + ut.x1, ut.ly * ut.lx * gdisp);
        }
        __syncthreads();                                         // the shared window tables are rewritten by the next sweep
    }
    if (tid < S * 12 && s_dP[tid] != 0.f) atomicAdd(p.dP + ((size_t)b * S) * 12 + tid, s_dP[tid]);
}

template <int S>
static cudaError_t launch_bwd_list_t(const PhotoDev& p, cudaStream_t st) {
    dim3 grid(kListCtas, p.nscales * p.B);
    photo_bwd_list_kernel<S><<<grid, 256, 0, st>>>(p);
    return cudaGetLastError();
}

template <int S, bool kTMA>
static cudaError_t launch_bwd_k(const PhotoDev& p, const PhotoMaps& maps, cudaStream_t st) {
    const size_t smem = (size_t)((1 + S) * kQGROUP + 9 * kQPLANE) * sizeof(float) + kQPLANE;
    static SmemOptIn opt_in;
    if (cudaError_t e = opt_in(photo_bwd_kernel<S, kTMA>, smem)) return e;
    dim3 grid((p.W + kTW - 1) / kTW, (p.H + kTH - 1) / kTH, p.B * p.nscales);
    photo_bwd_kernel<S, kTMA><<<grid, kNT, smem, st>>>(p, maps);
    return cudaGetLastError();
}

template <int S>
static cudaError_t launch_bwd_t(const PhotoDev& p, cudaStream_t st) {
    // TMA path: the forward materialised the warps, every tensor is 16-byte aligned and W % 4 == 0
    PhotoMaps maps;
    bool tma = p.use_tma && encode_image_map(&maps.tgt, p.target, p.B * 3, p.H, p.W, kQW, kQH, 3);
    for (int s = 0; tma && s < p.nscales; ++s)
        for (int f = 0; tma && f < S; ++f)
            tma = p.warped[s][f] && encode_image_map(&maps.img[s * S + f], p.warped[s][f], p.B * 3, p.H, p.W, kQW, kQH, 3);
    return tma ? launch_bwd_k<S, true>(p, maps, st) : launch_bwd_k<S, false>(p, maps, st);
}

cudaError_t launch_photo_fwd(const PhotoDev& p, cudaStream_t st) {       // fused single-kernel forward
    TDL_DISPATCH_S(launch_fwd_fused, p, st)
    return cudaErrorInvalidValue;
}

cudaError_t launch_photo_bwd(const PhotoDev& p, cudaStream_t st) {
    switch (p.S) {
        case 1: return launch_bwd_t<1>(p, st);
        case 2: return launch_bwd_t<2>(p, st);
        case 3: return launch_bwd_t<3>(p, st);
        case 4: return launch_bwd_t<4>(p, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_photo_bwd_list(const PhotoDev& p, cudaStream_t st) {
    TDL_DISPATCH_S(launch_bwd_list_t, p, st)
    return cudaErrorInvalidValue;
}


// ================================================================================================
// Masked image-reconstruction loss of the TripleD family (mono/model/mono_fm_joint_inpaint/net.py:80-91):
//     rho = 0.85 * mean_c SSIM(pred, tgt) + 0.15 * mean_c robust_l1(pred, tgt)            (B,1,h,w)
//     loss = coef * sum(rho * (1 - mask)) / sum(1 - mask)            mask (B,3,h,w) broadcasts over rho
// Forward: 32x32 tiles, reflect halo 1, the strip SSIM of the photometric path.  Backward: halo 2, window adjoint
// coefficients for every window, box-sum gather, dense store of d_pred (no atomics).
// ================================================================================================
struct ReconDev {
    int B, h, w;
    float coef;
    const float* pred;
    const float* tgt;
    const float* mask;       // may be null: plain mean over (B,1,h,w)
    double* acc;             // [2]: numerator, denominator
    float* loss;
    const float* dloss;
    float* d_pred;
};

TDL_DEV float recon_weight(const ReconDev& p, int b, size_t hw, size_t pix) {
    if (!p.mask) return 1.f;
    const float* m = p.mask + (size_t)b * 3 * hw + pix;
    return (1.f - __ldg(m)) + (1.f - __ldg(m + hw)) + (1.f - __ldg(m + 2 * hw));
}

__global__ void __launch_bounds__(kNT, 2) recon_fwd_kernel(const ReconDev p) {
    constexpr int PW = kTW + 2, PH = kTH + 2, PLANE = PW * PH;
    extern __shared__ float smem_recon[];
    float* s_tgt = smem_recon;                    // [3][PLANE]
    float* s_prd = smem_recon + 3 * PLANE;        // [3][PLANE]
    float* s_stats = smem_recon + 6 * PLANE + threadIdx.x;
    __shared__ float s_red[32];
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int b = blockIdx.z, tx0 = blockIdx.x * kTW, ty0 = blockIdx.y * kTH;
    const int H = p.h, W = p.w;
    const size_t HW = (size_t)H * W;
    for (int i = tid; i < PLANE; i += kNT) {
        const int r = i / PW, c = i - r * PW;
        const size_t o = (size_t)b * 3 * HW + (size_t)reflect1(ty0 - 1 + r, H) * W + reflect1(tx0 - 1 + c, W);
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            s_tgt[ch * PLANE + i] = __ldg(p.tgt + o + ch * HW);
            s_prd[ch * PLANE + i] = __ldg(p.pred + o + ch * HW);
        }
    }
    __syncthreads();
    const int r0 = wrp * kR, gx = tx0 + lane;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        float mu_y[kR], sg_y[kR];
        strip_target_stats<PW>(s_tgt + ch * PLANE, r0, lane, mu_y, sg_y);
#pragma unroll
        for (int i = 0; i < kR; ++i) {
            s_stats[((ch * 2) * kR + i) * kNT] = mu_y[i];
            s_stats[((ch * 2 + 1) * kR + i) * kNT] = sg_y[i];
        }
    }
    float rho[kR];
    strip_reprojection<PW>(s_prd, s_tgt, PLANE, r0, lane, s_stats, rho);
    float num = 0.f, den = 0.f;
#pragma unroll
    for (int i = 0; i < kR; ++i) {
        const int gy = ty0 + r0 + i;
        if (gx < W && gy < H) {
            const float wq = recon_weight(p, b, HW, (size_t)gy * W + gx);
            num += rho[i] * wq;
            den += wq;
        }
    }
    num = block_sum(num, s_red);
    if (tid == 0) atomicAdd(p.acc, (double)num);
    den = block_sum(den, s_red);
    if (tid == 0) atomicAdd(p.acc + 1, (double)den);
}

__global__ void recon_finalize_kernel(const double* acc, float coef, float* loss) {
    if (threadIdx.x == 0) loss[0] = __fmul_rn(coef, (float)(acc[0] / acc[1]));
}

__global__ void __launch_bounds__(kNT, 2) recon_bwd_kernel(const ReconDev p) {
    constexpr int QW = kTW + 4, QH = kTH + 4, QPLANE = QW * QH;
    constexpr int PW = kTW + 2, PH = kTH + 2;
    extern __shared__ float smem_recon[];
    float* s_tgt = smem_recon;                    // [3][QPLANE]
    float* s_prd = s_tgt + 3 * QPLANE;            // [3][QPLANE]
    float* s_coef = s_prd + 3 * QPLANE;           // [3 ch][3 coef][QPLANE]
    float* s_wq = s_coef + 9 * QPLANE;            // [QPLANE] window weight sum_c (1 - mask), 0 outside the image
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int b = blockIdx.z, tx0 = blockIdx.x * kTW, ty0 = blockIdx.y * kTH;
    const int H = p.h, W = p.w;
    const size_t HW = (size_t)H * W;
    const float up = __ldg(p.dloss) * p.coef / (float)p.acc[1];
    for (int i = tid; i < QPLANE; i += kNT) {
        const int r = i / QW, c = i - r * QW;
        const int ry = ty0 - 2 + r, rx = tx0 - 2 + c;
        const bool inside = ry >= 0 && ry < H && rx >= 0 && rx < W;
        const size_t pix = (size_t)reflect1(ry, H) * W + reflect1(rx, W);
        const size_t o = (size_t)b * 3 * HW + pix;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            s_tgt[ch * QPLANE + i] = __ldg(p.tgt + o + ch * HW);
            s_prd[ch * QPLANE + i] = __ldg(p.pred + o + ch * HW);
        }
        s_wq[i] = inside ? recon_weight(p, b, HW, pix) : 0.f;
    }
    __syncthreads();
    for (int i = tid; i < PH * PW; i += kNT) {
        const int r = i / PW, c = i - r * PW;
        const int q = (r + 1) * QW + (c + 1);
        const float k = up * s_wq[q] * (0.85f / 3.f) * (1.f / 9.f);
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            float cA = 0.f, cB = 0.f, cC = 0.f;
            if (k != 0.f) window_adjoint<QW>(s_prd + ch * QPLANE + q, s_tgt + ch * QPLANE + q, k, cA, cB, cC);
            s_coef[(ch * 3 + 0) * QPLANE + q] = cA;
            s_coef[(ch * 3 + 1) * QPLANE + q] = cB;
            s_coef[(ch * 3 + 2) * QPLANE + q] = cC;
        }
    }
    __syncthreads();
    const int r0 = wrp * kR, gx = tx0 + lane, qc = lane + 2;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        float box[3][kR];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float* pl = s_coef + (ch * 3 + k) * QPLANE + (r0 + 1) * QW + qc;
            float hsum[kR + 2];
#pragma unroll
            for (int j = 0; j < kR + 2; ++j) hsum[j] = pl[j * QW - 1] + pl[j * QW] + pl[j * QW + 1];
#pragma unroll
            for (int i = 0; i < kR; ++i) box[k][i] = hsum[i] + hsum[i + 1] + hsum[i + 2];
        }
#pragma unroll
        for (int i = 0; i < kR; ++i) {
            const int gy = ty0 + r0 + i;
            if (gx >= W || gy >= H) continue;
            const int q = (r0 + i + 2) * QW + qc;
            const bool bx0 = gx == 1, bx1 = gx == W - 2, by0 = gy == 1, by1 = gy == H - 2;
            if (bx0 || bx1 || by0 || by1) {                       // reflect-padding multiplicity (see photo_bwd_kernel)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float* c = s_coef + (ch * 3 + k) * QPLANE + q;
                    float e = 0.f;
                    if (bx0) e += c[-QW - 1] + c[-1] + c[QW - 1];
                    if (bx1) e += c[-QW + 1] + c[1] + c[QW + 1];
                    if (by0) e += c[-QW - 1] + c[-QW] + c[-QW + 1];
                    if (by1) e += c[QW - 1] + c[QW] + c[QW + 1];
                    if (bx0 && by0) e += c[-QW - 1];
                    if (bx1 && by0) e += c[-QW + 1];
                    if (bx0 && by1) e += c[QW - 1];
                    if (bx1 && by1) e += c[QW + 1];
                    box[k][i] += e;
                }
            }
            const float xv = s_prd[ch * QPLANE + q], yv = s_tgt[ch * QPLANE + q];
            const float df = xv - yv;
            const float g = box[0][i] + 2.f * xv * box[1][i] + yv * box[2][i] +
                            up * s_wq[q] * (0.15f / 3.f) * df * rsqrt_approx(df * df + kL1Eps2);
            p.d_pred[((size_t)b * 3 + ch) * HW + (size_t)gy * W + gx] = g;
        }
    }
}

cudaError_t launch_recon_fwd(const ReconArgsDev& a, cudaStream_t st) {
    ReconDev p{a.B, a.h, a.w, a.coef, a.pred, a.tgt, a.mask, a.acc, a.loss, a.dloss, a.d_pred};
    constexpr int PLANE = (kTW + 2) * (kTH + 2);
    const size_t smem = (size_t)6 * PLANE * sizeof(float) + (size_t)6 * kR * kNT * sizeof(float);
    static SmemOptIn opt_in;
    if (cudaError_t e = opt_in(recon_fwd_kernel, smem)) return e;
    dim3 grid((a.w + kTW - 1) / kTW, (a.h + kTH - 1) / kTH, a.B);
    recon_fwd_kernel<<<grid, kNT, smem, st>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    recon_finalize_kernel<<<1, 32, 0, st>>>(a.acc, a.coef, a.loss);
    return cudaGetLastError();
}

cudaError_t launch_recon_bwd(const ReconArgsDev& a, cudaStream_t st) {
    ReconDev p{a.B, a.h, a.w, a.coef, a.pred, a.tgt, a.mask, a.acc, a.loss, a.dloss, a.d_pred};
    constexpr int QPLANE = (kTW + 4) * (kTH + 4);
    const size_t smem = (size_t)16 * QPLANE * sizeof(float);
    static SmemOptIn opt_in;
    if (cudaError_t e = opt_in(recon_bwd_kernel, smem)) return e;
    dim3 grid((a.w + kTW - 1) / kTW, (a.h + kTH - 1) / kTH, a.B);
    recon_bwd_kernel<<<grid, kNT, smem, st>>>(p);
    return cudaGetLastError();
}

}  // namespace tdl
