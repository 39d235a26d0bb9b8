// Edge-aware first + second order smoothness (forward and backward) on (B,C,h,w) maps.
//
// Serves both
//   get_smooth_loss on the mean-normalised disparity   mono/model/mono_fm/net.py:123-131,255-283
//   get_feature_regularization_loss on feature maps    mono/model/mono_fm_joint/net.py:309-330
// One thread owns one low-resolution pixel: it evaluates the six difference stencils that
// START at that pixel (dx, dy, dxx, dxy, dyx, dyy), weights them with
// exp(-alpha * mean_c |same stencil of the area-downsampled image|), divides each by the
// element count of its own difference map (torch.mean) and reduces per image with
// warp shuffles + one double atomic per CTA.  All levels (scales) run in ONE launch
// (blockIdx.y = level).
//
// Mean-normalisation adjoint: with dhat = d / (m + eps) the loss L_b of image b is
// homogeneous of degree one in dhat, so sum_x dhat(x) * dL/ddhat(x) = L_b (Euler) and
//     dL/dd(y) = [ dL/ddhat(y) - L_b / (h*w) ] / (m + eps)
// needs no second reduction pass: L_b was stored by the forward.
#include "tdl_common.cuh"
#include "tdl_internal.h"

namespace tdl {

constexpr int kSmoothNT = 256;

struct Stencil6 {           // values of the six stencils anchored at one pixel
    float v[6];             // dx, dy, dxx, dxy, dyx, dyy
};

// stencils of plane `a` (pitch w) anchored at (j,i); out-of-range ones are left at 0
template <typename F>
TDL_DEV Stencil6 stencils(F a, int j, int i, int h, int w) {
    Stencil6 s;
#pragma unroll
    for (int k = 0; k < 6; ++k) s.v[k] = 0.f;
    const bool x1 = i + 1 < w, x2 = i + 2 < w, y1 = j + 1 < h, y2 = j + 2 < h;
    const float a00 = a(j, i);
    float dx0 = 0.f, dy0 = 0.f;
    if (x1) {
        dx0 = __fsub_rn(a(j, i + 1), a00);
        s.v[0] = dx0;
    }
    if (y1) {
        dy0 = __fsub_rn(a(j + 1, i), a00);
        s.v[1] = dy0;
    }
    if (x2) s.v[2] = __fsub_rn(__fsub_rn(a(j, i + 2), a(j, i + 1)), dx0);                   // d/dx (d/dx)
    if (x1 && y1) {
        const float a11 = a(j + 1, i + 1);
        s.v[3] = __fsub_rn(__fsub_rn(a11, a(j + 1, i)), dx0);                               // d/dy (d/dx)
        s.v[4] = __fsub_rn(__fsub_rn(a11, a(j, i + 1)), dy0);                               // d/dx (d/dy)
    }
    if (y2) s.v[5] = __fsub_rn(__fsub_rn(a(j + 2, i), a(j + 1, i)), dy0);                   // d/dy (d/dy)
    return s;
}

// 1 / (element count of each difference map) for a (B,C,h,w) tensor (0 for an empty map)
struct Counts {
    float inv[6];
};

TDL_DEV Counts inv_counts(int B, int C, int h, int w) {
    Counts c;
    const float bc = (float)B * (float)C;
    const float n[6] = {bc * h * (w - 1), bc * (h - 1) * w, bc * h * (w - 2),
                        bc * (h - 1) * (w - 1), bc * (h - 1) * (w - 1), bc * (h - 2) * w};
#pragma unroll
    for (int k = 0; k < 6; ++k) c.inv[k] = n[k] > 0.f ? 1.f / n[k] : 0.f;      // empty map: handled by the caller
    return c;
}

// edge weights exp(-alpha * mean_c |stencil(J)|) anchored at (j,i); 0 where the stencil does not fit
TDL_DEV Stencil6 edge_weights(const float* __restrict__ Jb, int j, int i, int h, int w, float alpha) {
    Stencil6 m;
#pragma unroll
    for (int k = 0; k < 6; ++k) m.v[k] = 0.f;
    const size_t hw = (size_t)h * w;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const float* pl = opaque(Jb + ch * hw);             // (32-bit pixel offsets on an opaque plane base: see tdl_common.cuh)
        const Stencil6 s = stencils([&](int jj, int ii) { return __ldg(pl + (unsigned)(jj * w + ii)); }, j, i, h, w);
#pragma unroll
        for (int k = 0; k < 6; ++k) m.v[k] += fabsf(s.v[k]);
    }
    const bool x1 = i + 1 < w, x2 = i + 2 < w, y1 = j + 1 < h, y2 = j + 2 < h;
    const bool ok[6] = {x1, y1, x2, x1 && y1, x1 && y1, y2};
#pragma unroll
    for (int k = 0; k < 6; ++k) m.v[k] = ok[k] ? expf(-alpha * div3(m.v[k])) : 0.f;
    return m;
}

TDL_DEV float norm_denominator(const SmoothLevel& L, int b) {
    // disp.mean(2).mean(3) + 1e-7   (net.py:124-125)
    const double sum = L.acc[(size_t)b * L.acc_stride + 1];
    return __fadd_rn((float)(sum / ((double)L.h * (double)L.w)), 1e-7f);
}

__global__ void __launch_bounds__(kSmoothNT) smooth_fwd_kernel(const SmoothDev p) {
    __shared__ float s_red[32];
    const SmoothLevel& L = p.lv[blockIdx.y];
    const int b = blockIdx.z;
    const int h = L.h, w = L.w, C = L.C;
    const int pix = blockIdx.x * kSmoothNT + threadIdx.x;
    if ((int)(blockIdx.x * kSmoothNT) >= h * w) return;       // whole CTA outside this level
    float first = 0.f, second = 0.f;
    if (pix < h * w) {
        const int j = pix / w, i = pix - j * w;
        const Stencil6 wt = edge_weights(L.J + (size_t)b * 3 * h * w, j, i, h, w, L.alpha);
#pragma unroll
        for (int k = 0; k < 6; ++k) L.Wt[((size_t)b * 6 + k) * h * w + pix] = wt.v[k];       // for the backward
        const Counts cn = inv_counts(p.B, C, h, w);
        const float den = L.norm ? norm_denominator(L, b) : 1.f;
        for (int c = 0; c < C; ++c) {
            const float* pl = opaque(L.x + ((size_t)b * C + c) * h * w);
            const Stencil6 s = stencils(
                [&](int jj, int ii) {
                    const float v = __ldg(pl + (unsigned)(jj * w + ii));
                    return L.norm ? div_rn(v, den) : v;
                },
                j, i, h, w);
            first += fabsf(s.v[0]) * wt.v[0] * cn.inv[0] + fabsf(s.v[1]) * wt.v[1] * cn.inv[1];
            second += fabsf(s.v[2]) * wt.v[2] * cn.inv[2] + fabsf(s.v[3]) * wt.v[3] * cn.inv[3] +
                      fabsf(s.v[4]) * wt.v[4] * cn.inv[4] + fabsf(s.v[5]) * wt.v[5] * cn.inv[5];
        }
    }
    first = block_sum(first, s_red);
    if (threadIdx.x == 0) atomicAdd(L.acc + (size_t)b * L.acc_stride + 2, (double)first);
    second = block_sum(second, s_red);
    if (threadIdx.x == 0) atomicAdd(L.acc + (size_t)b * L.acc_stride + 3, (double)second);
}

// torch.mean of an EMPTY difference map is NaN (e.g. d_dyy of a 2-row map).  The accumulators stay
// finite (the backward needs them); the NaN is injected when the final scalar is formed.
TDL_DEV float nan_if_empty(float v, bool empty) { return empty ? __int_as_float(0x7fc00000) : v; }

TDL_DEV float sgn(float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); }

__global__ void __launch_bounds__(kSmoothNT) smooth_bwd_kernel(const SmoothDev p) {
    if (p.zero_ptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)
        for (int i = threadIdx.x; i < p.zero_n; i += kSmoothNT) p.zero_ptr[i] = 0.f;
    const SmoothLevel& L = p.lv[blockIdx.y];
    const int b = blockIdx.z;
    const int h = L.h, w = L.w, C = L.C;
    const int pix = blockIdx.x * kSmoothNT + threadIdx.x;
    if (pix >= h * w) return;
    const int j = pix / w, i = pix - j * w;
    const Counts cn = inv_counts(p.B, C, h, w);
    const float up = __ldg(L.dloss);
    const float c1 = up * L.first_coef, c2 = up * L.second_coef;
    const float kcoef[6] = {c1 * cn.inv[0], c1 * cn.inv[1], c2 * cn.inv[2], c2 * cn.inv[3], c2 * cn.inv[4], c2 * cn.inv[5]};
    const float den = L.norm ? norm_denominator(L, b) : 1.f;

    // anchors whose stencils touch (j,i): (0,0) (0,-1) (0,-2) (-1,0) (-2,0) (-1,-1)
    const int aj[6] = {0, 0, 0, -1, -2, -1};
    const int ai[6] = {0, -1, -2, 0, 0, -1};
    // coefficient of d(j,i) inside stencil k anchored at anchor a (0 = not touched)
    //                      dx   dy   dxx  dxy  dyx  dyy
    const float cf[6][6] = {{-1.f, -1.f, 1.f, 1.f, 1.f, 1.f},      // anchor (0,0)
                            {1.f, 0.f, -2.f, -1.f, -1.f, 0.f},     // anchor (0,-1)
                            {0.f, 0.f, 1.f, 0.f, 0.f, 0.f},        // anchor (0,-2)
                            {0.f, 1.f, 0.f, -1.f, -1.f, -2.f},     // anchor (-1,0)
                            {0.f, 0.f, 0.f, 0.f, 0.f, 1.f},        // anchor (-2,0)
                            {0.f, 0.f, 0.f, 1.f, 1.f, 0.f}};       // anchor (-1,-1)
    Stencil6 wt[6];
    const float* wtb = opaque(L.Wt + (size_t)b * 6 * h * w);
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        const int jj = j + aj[a], ii = i + ai[a];
        if (jj >= 0 && ii >= 0) {
#pragma unroll
            for (int k = 0; k < 6; ++k)
                wt[a].v[k] = (cf[a][k] != 0.f) ? __ldg(wtb + (unsigned)(k * h * w + jj * w + ii)) : 0.f;
        } else {
#pragma unroll
            for (int k = 0; k < 6; ++k) wt[a].v[k] = 0.f;
        }
    }
    // Euler term of the mean-normalisation adjoint (see file header)
    float euler = 0.f;
    if (L.norm) {
        const double Lb = (double)L.first_coef * L.acc[(size_t)b * L.acc_stride + 2] +
                          (double)L.second_coef * L.acc[(size_t)b * L.acc_stride + 3];
        euler = up * (float)(Lb / ((double)h * (double)w));
    }
    for (int c = 0; c < C; ++c) {
        const float* pl = opaque(L.x + ((size_t)b * C + c) * h * w);
        auto ld = [&](int jj, int ii) {
            const float v = __ldg(pl + (unsigned)(jj * w + ii));
            return L.norm ? div_rn(v, den) : v;
        };
        float g = 0.f;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            const int jj = j + aj[a], ii = i + ai[a];
            if (jj >= 0 && ii >= 0) {
                const Stencil6 s = stencils(ld, jj, ii, h, w);
#pragma unroll
                for (int k = 0; k < 6; ++k)
                    if (cf[a][k] != 0.f) g += cf[a][k] * sgn(s.v[k]) * wt[a].v[k] * kcoef[k];
            }
        }
        if (L.norm) g = (g - euler) / den;
        L.dx[((size_t)b * C + c) * h * w + pix] = g;
    }
}

// F.interpolate(img, (h,w), mode='area') for an integer factor: mean over fac x fac blocks.
__global__ void __launch_bounds__(256) area_pyramid_kernel(const float* __restrict__ img, int H, int W,
                                                          float* __restrict__ J, int h, int w, int fac, size_t total) {
    const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= total) return;
    const int i = idx % w;
    const int j = (idx / w) % h;
    const size_t bc = idx / ((size_t)w * h);
    const float* base = img + bc * (size_t)H * W + (size_t)j * fac * W + (size_t)i * fac;
    float acc = 0.f;
    for (int dy = 0; dy < fac; ++dy)
        for (int dx = 0; dx < fac; ++dx) acc += __ldg(base + (size_t)dy * W + dx);
    J[idx] = acc * (1.f / (float)(fac * fac));
}

// all levels in one launch (blockIdx.y = level); the first CTA of a level also clears that level's accumulators
struct PyrLevels {
    int n, B, H, W;
    float* J[TDL_MAX_LEVELS];
    double* acc[TDL_MAX_LEVELS];
    int h[TDL_MAX_LEVELS], w[TDL_MAX_LEVELS];
};

__global__ void __launch_bounds__(256) area_pyramid_multi_kernel(const float* __restrict__ img, const PyrLevels p) {
    const int l = blockIdx.y;
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < p.B * 4; i += 256) p.acc[l][i] = 0.0;
    const int h = p.h[l], w = p.w[l], fac = p.H / h;
    const size_t total = (size_t)p.B * 3 * h * w;
    for (size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (size_t)gridDim.x * 256) {
        const int i = idx % w;
        const int j = (idx / w) % h;
        const size_t bc = idx / ((size_t)w * h);
        const float* base = img + bc * (size_t)p.H * p.W + (size_t)j * fac * p.W + (size_t)i * fac;
        float acc = 0.f;
        for (int dy = 0; dy < fac; ++dy)
            for (int dx = 0; dx < fac; ++dx) acc += __ldg(base + (size_t)dy * p.W + dx);
        p.J[l][idx] = acc * (1.f / (float)(fac * fac));
    }
}

// ---------------------------------------------------------------------------------------------
// final scalars
__device__ void photo_finalize_body(const double* __restrict__ acc, int B, int nscales, double inv_bhw,
                                    const float* photo_coef4, const float* smooth_coef4, const int* hh, const int* ww,
                                    float* __restrict__ losses) {
    // one warp; lane s < nscales -> photometric, lane nscales+s -> smoothness
    const int t = threadIdx.x;
    if (t < nscales) {
        double sum = 0.0;
        for (int b = 0; b < B; ++b) sum += acc[((size_t)t * B + b) * 4 + 0];
        const float mean = (float)(sum * inv_bhw);
        losses[t] = __fmul_rn(photo_coef4[t], mean);                 // .mean() / len(scales)
    } else if (t < 2 * nscales) {
        const int s = t - nscales;
        double f1 = 0.0, f2 = 0.0;
        for (int b = 0; b < B; ++b) {
            f1 += acc[((size_t)s * B + b) * 4 + 2];
            f2 += acc[((size_t)s * B + b) * 4 + 3];
        }
        const float sm = __fadd_rn(nan_if_empty((float)f1, hh[s] < 2 || ww[s] < 2),
                                   nan_if_empty((float)f2, hh[s] < 3 || ww[s] < 3));   // smooth1 + smooth2 (net.py:278)
        losses[t] = __fmul_rn(smooth_coef4[s], sm);
    }
}

struct Coef8 {
    float photo[TDL_MAX_SCALES], smooth[TDL_MAX_SCALES];
    int h[TDL_MAX_SCALES], w[TDL_MAX_SCALES];
};

__global__ void photo_finalize_kernel(const double* __restrict__ acc, int B, int nscales, double inv_bhw, Coef8 c,
                                      float* __restrict__ losses) {
    photo_finalize_body(acc, B, nscales, inv_bhw, c.photo, c.smooth, c.h, c.w, losses);
}

__global__ void edge_finalize_kernel(const double* __restrict__ acc, int stride, int B, float c1, float c2, int h, int w,
                                     float* __restrict__ loss) {
    if (threadIdx.x == 0) {
        double f1 = 0.0, f2 = 0.0;
        for (int b = 0; b < B; ++b) {
            f1 += acc[(size_t)b * stride + 2];
            f2 += acc[(size_t)b * stride + 3];
        }
        loss[0] = __fadd_rn(__fmul_rn(c1, nan_if_empty((float)f1, h < 2 || w < 2)),
                            __fmul_rn(c2, nan_if_empty((float)f2, h < 3 || w < 3)));
    }
}

struct FinLevels {
    int n, B;
    const double* acc[TDL_MAX_LEVELS];
    float c1[TDL_MAX_LEVELS], c2[TDL_MAX_LEVELS];
    int h[TDL_MAX_LEVELS], w[TDL_MAX_LEVELS];
    float* loss[TDL_MAX_LEVELS];
};

__global__ void edge_finalize_multi_kernel(const FinLevels p) {
    const int l = threadIdx.x;
    if (l < p.n) {
        double f1 = 0.0, f2 = 0.0;
        for (int b = 0; b < p.B; ++b) {
            f1 += p.acc[l][(size_t)b * 4 + 2];
            f2 += p.acc[l][(size_t)b * 4 + 3];
        }
        p.loss[l][0] = __fadd_rn(__fmul_rn(p.c1[l], nan_if_empty((float)f1, p.h[l] < 2 || p.w[l] < 2)),
                                 __fmul_rn(p.c2[l], nan_if_empty((float)f2, p.h[l] < 3 || p.w[l] < 3)));
    }
}

// ---------------------------------------------------------------------------------------------
static void smooth_grid(const SmoothDev& p, dim3& grid) {
    int maxpix = 1;
    for (int l = 0; l < p.nlevels; ++l) maxpix = max(maxpix, p.lv[l].h * p.lv[l].w);
    grid = dim3((maxpix + kSmoothNT - 1) / kSmoothNT, p.nlevels, p.B);
}

cudaError_t launch_smooth_fwd(const SmoothDev& p, cudaStream_t st) {
    dim3 grid;
    smooth_grid(p, grid);
    smooth_fwd_kernel<<<grid, kSmoothNT, 0, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_smooth_bwd(const SmoothDev& p, cudaStream_t st) {
    dim3 grid;
    smooth_grid(p, grid);
    smooth_bwd_kernel<<<grid, kSmoothNT, 0, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_area_pyramid(const float* img, int B, int H, int W, float* J, int h, int w, cudaStream_t st) {
    const size_t total = (size_t)B * 3 * h * w;
    area_pyramid_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(img, H, W, J, h, w, H / h, total);
    return cudaGetLastError();
}

cudaError_t launch_area_pyramid_multi(const float* img, int H, int W, const SmoothDev& p, cudaStream_t st) {
    PyrLevels a;
    a.n = p.nlevels; a.B = p.B; a.H = H; a.W = W;
    size_t most = 1;
    for (int l = 0; l < p.nlevels; ++l) {
        a.J[l] = const_cast<float*>(p.lv[l].J);
        a.acc[l] = p.lv[l].acc;
        a.h[l] = p.lv[l].h;
        a.w[l] = p.lv[l].w;
        most = max(most, (size_t)p.B * 3 * p.lv[l].h * p.lv[l].w);
    }
    area_pyramid_multi_kernel<<<dim3((unsigned)((most + 255) / 256), p.nlevels), 256, 0, st>>>(img, a);
    return cudaGetLastError();
}

cudaError_t launch_edge_finalize_multi(const SmoothDev& p, float* const* loss, cudaStream_t st) {
    FinLevels a;
    a.n = p.nlevels; a.B = p.B;
    for (int l = 0; l < p.nlevels; ++l) {
        a.acc[l] = p.lv[l].acc;
        a.c1[l] = p.lv[l].first_coef;
        a.c2[l] = p.lv[l].second_coef;
        a.h[l] = p.lv[l].h;
        a.w[l] = p.lv[l].w;
        a.loss[l] = loss[l];
    }
    edge_finalize_multi_kernel<<<1, 32, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_photo_finalize(const PhotoDev& p, const float* photo_coef, const float* smooth_coef, float* losses,
                                  cudaStream_t st) {
    Coef8 c;
    for (int s = 0; s < TDL_MAX_SCALES; ++s) {
        c.photo[s] = photo_coef[s];
        c.smooth[s] = smooth_coef[s];
        c.h[s] = p.dh[s];
        c.w[s] = p.dw[s];
    }
    const double inv_bhw = 1.0 / ((double)p.B * p.H * p.W);
    photo_finalize_kernel<<<1, 32, 0, st>>>(p.acc, p.B, p.nscales, inv_bhw, c, losses);
    return cudaGetLastError();
}

cudaError_t launch_edge_finalize(const double* acc, int acc_stride, int B, float first_coef, float second_coef,
                                 int h, int w, float* loss, cudaStream_t st) {
    edge_finalize_kernel<<<1, 32, 0, st>>>(acc, acc_stride, B, first_coef, second_coef, h, w, loss);
    return cudaGetLastError();
}

}  // namespace tdl
