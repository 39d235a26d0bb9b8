// Feature-metric (FeatDepth) loss, forward and backward (sm_100a).
//
//   generate_features_pred + compute_perceptional_loss + min over source frames
//   mono/model/mono_fm/net.py:59-61,111-118,172-199; mono/model/mono_fm_joint_inpaint/net.py:58-70
//
// One thread = one feature-map pixel.  The warp of a pixel is computed once per source
// frame (same projection code as the photometric path, half-resolution intrinsics), then
// the C channels are streamed: lanes hold consecutive x, so the target reads and the
// warped-feature writes are fully coalesced and the 4-tap gathers hit neighbouring lines.
// HBM-bound: (1 + S) * C * 4 B read + S * C * 4 B written per pixel in the forward.
#include "tdl_common.cuh"
#include "tdl_internal.h"

namespace tdl {

constexpr int kFeatNT = 128;

template <int S>
__global__ void __launch_bounds__(kFeatNT) feat_fwd_kernel(const FeatDev p) {
    __shared__ float s_red[32];
    __shared__ float s_cam[TDL_MAX_SRC * 12 + 9];
    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const int h = p.h, w = p.w, C = p.C;
    const size_t hw = (size_t)h * w;
    if (tid < S * 12) s_cam[tid] = __ldg(p.P + (size_t)b * S * 12 + tid);
    if (tid >= 64 && tid < 64 + 9) s_cam[TDL_MAX_SRC * 12 + tid - 64] = __ldg(p.invK + (size_t)b * 9 + tid - 64);
    __syncthreads();
    const int pix = blockIdx.x * kFeatNT + tid;
    float best = 0.f;
    if (pix < (int)hw) {
        const int y = pix / w, x = pix - y * w;
        const DepthParams dp{p.min_disp, p.range};
        const UpTap ut = up_tap(y, x, p.sy, p.sx, p.dh, p.dw);
        const Geo g = backproject(up_value(p.disp + (size_t)b * p.dh * p.dw, p.dw, ut), dp,
                                  s_cam + TDL_MAX_SRC * 12, x, y);
        Bilin bt[S];
#pragma unroll
        for (int f = 0; f < S; ++f) {
            const Proj pr = project<false>(g, s_cam + f * 12, h, w, p.align_corners);
            bt[f] = bilin_taps(pr.ix, pr.iy, h, w);
        }
        float acc[S];
#pragma unroll
        for (int f = 0; f < S; ++f) acc[f] = 0.f;
        const float* tb = p.tgt + (size_t)b * C * hw + pix;
#pragma unroll 2
        for (int c = 0; c < C; ++c) {
            const float t = __ldg(tb + (size_t)c * hw);
#pragma unroll
            for (int f = 0; f < S; ++f) {
                const float v = bilin_sample(p.src[f] + ((size_t)b * C + c) * hw, w, bt[f]);
                if (p.warped[f]) p.warped[f][((size_t)b * C + c) * hw + pix] = v;
                const float df = __fsub_rn(v, t);                                  // robust_l1(tgt_f, src_f)
                acc[f] += __fsqrt_rn(__fadd_rn(__fmul_rn(df, df), kL1Eps2));
            }
        }
        int arg = 0;
        const float fc = (float)C;
        best = __fdiv_rn(acc[0], fc);
#pragma unroll
        for (int f = 1; f < S; ++f) {
            const float v = __fdiv_rn(acc[f], fc);
            if (v < best) {
                best = v;
                arg = f;
            }
        }
        p.argmin[(size_t)b * hw + pix] = (unsigned char)arg;
        if (p.min_index) p.min_index[(size_t)b * hw + pix] = arg;
    }
    best = block_sum(best, s_red);
    if (tid == 0) atomicAdd(p.acc + b, (double)best);
}

__global__ void feat_finalize_kernel(const double* __restrict__ acc, int B, double inv_n, float coef, float* loss) {
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int b = 0; b < B; ++b) s += acc[b];
        loss[0] = __fmul_rn(coef, (float)(s * inv_n));
    }
}

// Backward: only the arg-min source of each pixel receives gradient (torch.min backward).
template <bool kGradFeat>
__global__ void __launch_bounds__(kFeatNT) feat_bwd_kernel(const FeatDev p) {
    __shared__ float s_cam[TDL_MAX_SRC * 12 + 9];
    __shared__ float s_dP[TDL_MAX_SRC * 12];
    const int tid = threadIdx.x, lane = tid & 31;
    const int b = blockIdx.y;
    const int h = p.h, w = p.w, C = p.C, S = p.S;
    const size_t hw = (size_t)h * w;
    if (tid < S * 12) {
        s_cam[tid] = __ldg(p.P + (size_t)b * S * 12 + tid);
        s_dP[tid] = 0.f;
    }
    if (tid >= 64 && tid < 64 + 9) s_cam[TDL_MAX_SRC * 12 + tid - 64] = __ldg(p.invK + (size_t)b * 9 + tid - 64);
    __syncthreads();
    const int pix = blockIdx.x * kFeatNT + tid;
    const bool active = pix < (int)hw;
    float aP[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) aP[k] = 0.f;
    int fsel = 0;
    if (active) {
        const int y = pix / w, x = pix - y * w;
        fsel = p.argmin[(size_t)b * hw + pix];
        const DepthParams dp{p.min_disp, p.range};
        const UpTap ut = up_tap(y, x, p.sy, p.sx, p.dh, p.dw);
        const Geo g = backproject(up_value(p.disp + (size_t)b * p.dh * p.dw, p.dw, ut), dp,
                                  s_cam + TDL_MAX_SRC * 12, x, y);
        const float* Pf = s_cam + fsel * 12;
        const Proj pr = project<true>(g, Pf, h, w, p.align_corners);
        const Bilin bt = bilin_taps(pr.ix, pr.iy, h, w);
        const float up = __ldg(p.dloss) * p.coef / ((float)p.B * (float)h * (float)w) / (float)C;
        const float* sb = p.src[0];
#pragma unroll
        for (int f = 1; f < TDL_MAX_SRC; ++f)
            if (f == fsel) sb = p.src[f];
        float* dsb = nullptr;
        if (kGradFeat) {
            dsb = p.d_src[0];
#pragma unroll
            for (int f = 1; f < TDL_MAX_SRC; ++f)
                if (f == fsel) dsb = p.d_src[f];
        }
        const float* tb = p.tgt + (size_t)b * C * hw + pix;
        float gix = 0.f, giy = 0.f;
#pragma unroll 2
        for (int c = 0; c < C; ++c) {
            const size_t plane = ((size_t)b * C + c) * hw;
            const float t = __ldg(tb + (size_t)c * hw);
            float dix, diy;
            const float v = bilin_sample_grad(sb + plane, w, bt, dix, diy);
            const float df = v - t;
            const float gvv = up * df / sqrtf(df * df + kL1Eps2);      // d loss / d warped value
            gix += gvv * dix;
            giy += gvv * diy;
            if (kGradFeat) {
                if (p.d_tgt) p.d_tgt[plane + pix] = -gvv;
                if (dsb) {
                    float* q = dsb + plane + (size_t)bt.y0 * w + bt.x0;
                    atomicAdd(q, gvv * bt.nw);
                    if (bt.vx) atomicAdd(q + 1, gvv * bt.ne);
                    if (bt.vy) atomicAdd(q + w, gvv * bt.sw);
                    if (bt.vx && bt.vy) atomicAdd(q + w + 1, gvv * bt.se);
                }
            }
        }
        const float gu = gix * pr.mx, gv = giy * pr.my;
        const float rz = 1.f / pr.z;
        const float gp0 = gu * rz, gp1 = gv * rz, gp2 = -(gu * pr.u + gv * pr.v) * rz;
        aP[0] = gp0 * g.X0; aP[1] = gp0 * g.X1; aP[2] = gp0 * g.X2; aP[3] = gp0;
        aP[4] = gp1 * g.X0; aP[5] = gp1 * g.X1; aP[6] = gp1 * g.X2; aP[7] = gp1;
        aP[8] = gp2 * g.X0; aP[9] = gp2 * g.X1; aP[10] = gp2 * g.X2; aP[11] = gp2;
        const float gX0 = Pf[0] * gp0 + Pf[4] * gp1 + Pf[8] * gp2;
        const float gX1 = Pf[1] * gp0 + Pf[5] * gp1 + Pf[9] * gp2;
        const float gX2 = Pf[2] * gp0 + Pf[6] * gp1 + Pf[10] * gp2;
        const float gD = gX0 * g.r0 + gX1 * g.r1 + gX2 * g.r2;
        const float gdisp = -p.range * g.D * g.D * gD;
        float* dd = p.d_disp + (size_t)b * p.dh * p.dw;
        if (p.dh == h && p.dw == w) {
            dd[pix] = gdisp;                         // identity resize: plain store, no atomics
        } else {                                     // adjoint of the bilinear resize (d_disp was zeroed)
            const float hy = 1.f - ut.ly, hx = 1.f - ut.lx;
            atomicAdd(dd + (size_t)ut.y0 * p.dw + ut.x0, hy * hx * gdisp);
            atomicAdd(dd + (size_t)ut.y0 * p.dw + ut.x1, hy * ut.lx * gdisp);
            atomicAdd(dd + (size_t)ut.y1 * p.dw + ut.x0, ut.ly * hx * gdisp);
            atomicAdd(dd + (size_t)ut.y1 * p.dw + ut.x1, ut.ly * ut.lx * gdisp);
        }
    }
    // dP: per source frame, reduce the 12 partials over the warp, then one shared atomic per warp
    for (int f = 0; f < S; ++f) {
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            const float v = warp_sum((active && f == fsel) ? aP[k] : 0.f);
            if (lane == 0 && v != 0.f) atomicAdd(&s_dP[f * 12 + k], v);
        }
    }
    __syncthreads();
    if (tid < S * 12 && s_dP[tid] != 0.f) atomicAdd(p.dP + (size_t)b * S * 12 + tid, s_dP[tid]);
}

cudaError_t launch_feat_fwd(const FeatDev& p, cudaStream_t st) {
    dim3 grid((unsigned)(((size_t)p.h * p.w + kFeatNT - 1) / kFeatNT), p.B);
    switch (p.S) {
        case 1: feat_fwd_kernel<1><<<grid, kFeatNT, 0, st>>>(p); break;
        case 2: feat_fwd_kernel<2><<<grid, kFeatNT, 0, st>>>(p); break;
        case 3: feat_fwd_kernel<3><<<grid, kFeatNT, 0, st>>>(p); break;
        case 4: feat_fwd_kernel<4><<<grid, kFeatNT, 0, st>>>(p); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_feat_finalize(const FeatDev& p, cudaStream_t st) {
    feat_finalize_kernel<<<1, 32, 0, st>>>(p.acc, p.B, 1.0 / ((double)p.B * p.h * p.w), p.coef, p.loss);
    return cudaGetLastError();
}

cudaError_t launch_feat_bwd(const FeatDev& p, cudaStream_t st) {
    dim3 grid((unsigned)(((size_t)p.h * p.w + kFeatNT - 1) / kFeatNT), p.B);
    if (p.d_tgt || p.d_src[0])
        feat_bwd_kernel<true><<<grid, kFeatNT, 0, st>>>(p);
    else
        feat_bwd_kernel<false><<<grid, kFeatNT, 0, st>>>(p);
    return cudaGetLastError();
}

}  // namespace tdl
