// Feature-metric (FeatDepth) loss, forward and backward (sm_100a).
//
//   generate_features_pred + compute_perceptional_loss + min over source frames
//   mono/model/mono_fm/net.py:59-61,111-118,172-199; mono/model/mono_fm_joint_inpaint/net.py:58-70
//
// Forward / per-pixel backward: one thread = one feature-map pixel.  The warp of a pixel is computed once per source
// frame (same projection code as the photometric path, half-resolution intrinsics), then the C channels are streamed:
// lanes hold consecutive x, so the target reads and the warped-feature writes are fully coalesced; the 4-tap gathers
// hit neighbouring lines for a smooth flow and one line per lane for the bench's pixel-level flow, which makes these
// kernels L1/TEX-wavefront bound there (ncu: 85 % forward, 55 % backward).
// Algorithmic traffic: (1 + S) * C * 4 B read + S * C * 4 B written per pixel in the forward.
// Backward with trainable features: grid_sample's d_src scatter runs as a bucketed GATHER (feat_bwd_bucket_kernel +
// feat_gather_kernel + feat_overflow_kernel below); feat_bwd_kernel<true> is the atomic-scatter fallback.
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
        // corner offsets with clamped east / south taps: a clamped tap always has weight 0 (ix == w-1 or
        // iy == h-1 after border clipping), so loading it unconditionally adds exactly 0 -- and every load
        // of a channel batch can be issued before the first use (memory-level parallelism).
        int o00[S], o01[S], o10[S], o11[S];
#pragma unroll
        for (int f = 0; f < S; ++f) {
            const int x1 = min(bt[f].x0 + 1, w - 1), y1 = min(bt[f].y0 + 1, h - 1);
            o00[f] = bt[f].y0 * w + bt[f].x0;
            o01[f] = bt[f].y0 * w + x1;
            o10[f] = y1 * w + bt[f].x0;
            o11[f] = y1 * w + x1;
        }
        constexpr int CB = (S <= 2) ? 8 : 4;                       // channels per batch
        // block-uniform 64-bit bases + 32-bit per-thread offsets: one integer add per load / store
        const unsigned uhw = (unsigned)hw;
        const float* tbase = p.tgt + (size_t)b * C * hw;
        const float* sbase[S];
        float* wbase[S];
#pragma unroll
        for (int f = 0; f < S; ++f) {
            sbase[f] = p.src[f] + (size_t)b * C * hw;
            wbase[f] = p.warped[f] ? p.warped[f] + (size_t)b * C * hw : nullptr;
        }
        for (int c0 = 0; c0 < C; c0 += CB) {
            float t[CB], v[S][CB][4];
#pragma unroll
            for (int j = 0; j < CB; ++j) {
                const unsigned co = (unsigned)min(c0 + j, C - 1) * uhw;
                t[j] = __ldcs(tbase + (co + (unsigned)pix));          // read once: keep L1 for the gathers
#pragma unroll
                for (int f = 0; f < S; ++f) {
                    v[f][j][0] = __ldg(sbase[f] + (co + (unsigned)o00[f]));
                    v[f][j][1] = __ldg(sbase[f] + (co + (unsigned)o01[f]));
                    v[f][j][2] = __ldg(sbase[f] + (co + (unsigned)o10[f]));
                    v[f][j][3] = __ldg(sbase[f] + (co + (unsigned)o11[f]));
                }
            }
#pragma unroll
            for (int j = 0; j < CB; ++j) {
                if (c0 + j < C) {
                    const unsigned co = (unsigned)(c0 + j) * uhw + (unsigned)pix;
#pragma unroll
                    for (int f = 0; f < S; ++f) {
                        float val = v[f][j][0] * bt[f].nw;
                        val += v[f][j][1] * bt[f].ne;
                        val += v[f][j][2] * bt[f].sw;
                        val += v[f][j][3] * bt[f].se;
                        if (wbase[f]) __stcs(wbase[f] + co, val);
                        const float df = __fsub_rn(val, t[j]);                     // robust_l1(tgt_f, src_f)
                        acc[f] += sqrt_fast(__fadd_rn(__fmul_rn(df, df), kL1Eps2));
                    }
                }
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
//
// d_src scatter: a pixel adds g*w to the four corners of its sampling point.  Global reductions (RED) issue
// at ~1.3 cycles per lane per SM on B200, so 4 atomics x C channels per pixel would bound the kernel.  Along
// a row the sampling points of neighbouring pixels are (almost always) one source pixel apart, i.e. the
// north-east / south-east corner of lane i IS the north-west / south-west corner of lane i+1: lane i+1 takes
// those two contributions over with a warp shuffle and lane i skips its two atomics ("absorbed").
template <bool kGradFeat>
__global__ void __launch_bounds__(kFeatNT) feat_bwd_kernel(const FeatDev p) {
    __shared__ float s_cam[TDL_MAX_SRC * 12 + 9];
    __shared__ float s_dP[TDL_MAX_SRC * 12];
    // CTA = 32 columns x 4 consecutive rows (one warp per row): the south taps of row r are the north taps of
    // row r + 1 wherever the flow is regular, so they are handed down through shared memory and only the last
    // row of the CTA sends its south taps to L2 (5 instead of 8 reductions per element column and channel).
    __shared__ int s_o00[kFeatNT], s_fs[kFeatNT];
    __shared__ float s_bot[4][kFeatNT];
    const int tid = threadIdx.x, lane = tid & 31, wr = tid >> 5;
    const int b = blockIdx.z;
    const int h = p.h, w = p.w, C = p.C, S = p.S;
    const size_t hw = (size_t)h * w;
    if (tid < S * 12) {
        s_cam[tid] = __ldg(p.P + (size_t)b * S * 12 + tid);
        s_dP[tid] = 0.f;
    }
    if (tid >= 64 && tid < 64 + 9) s_cam[TDL_MAX_SRC * 12 + tid - 64] = __ldg(p.invK + (size_t)b * 9 + tid - 64);
    __syncthreads();
    const int xq = blockIdx.x * 32 + lane, yq = blockIdx.y * 4 + wr;
    const bool active = xq < w && yq < h;
    const int y = active ? yq : 0, x = active ? xq : 0;
    const int pix = y * w + x;
    const int fsel = active ? p.argmin[(size_t)b * hw + pix] : -1;
    const DepthParams dp{p.min_disp, p.range};
    const UpTap ut = up_tap(y, x, p.sy, p.sx, p.dh, p.dw);
    const Geo g = backproject(up_value(p.disp + (size_t)b * p.dh * p.dw, p.dw, ut), dp, s_cam + TDL_MAX_SRC * 12, x, y);
    const float* Pf = s_cam + max(fsel, 0) * 12;
    const Proj pr = project<true>(g, Pf, h, w, p.align_corners);
    const Bilin bt = bilin_taps(pr.ix, pr.iy, h, w);
    const float up = active ? __ldg(p.dloss) * p.coef / ((float)p.Bnorm * (float)h * (float)w) / (float)C : 0.f;
    const float* sb = p.src[0];
    float* dsb = kGradFeat ? p.d_src[0] : nullptr;
#pragma unroll
    for (int f = 1; f < TDL_MAX_SRC; ++f)
        if (f == fsel) {
            sb = p.src[f];
            if (kGradFeat) dsb = p.d_src[f];
        }
    const int x1 = min(bt.x0 + 1, w - 1), y1 = min(bt.y0 + 1, h - 1);
    const int o00 = bt.y0 * w + bt.x0, o01 = bt.y0 * w + x1, o10 = y1 * w + bt.x0, o11 = y1 * w + x1;

    // absorption pattern (channel independent)
    bool absorbed = false, absorbs = false;          // my east taps are taken over by lane+1 / I take lane-1's
    if (kGradFeat && dsb) {
        const int nx0 = __shfl_down_sync(0xffffffffu, bt.x0, 1), ny0 = __shfl_down_sync(0xffffffffu, bt.y0, 1);
        const int nf = __shfl_down_sync(0xffffffffu, fsel, 1);
        absorbed = active && lane < 31 && nf == fsel && ny0 == bt.y0 && nx0 == bt.x0 + 1 && bt.vx;
        absorbs = __shfl_up_sync(0xffffffffu, (int)absorbed, 1) != 0 && lane > 0;
    }

    // vertical pattern: the row below samples one source row further down in the same source column
    bool gives_down = false, takes_up = false;
    if (kGradFeat && dsb) {
        s_o00[tid] = active ? o00 : -1;
        s_fs[tid] = fsel;
        __syncthreads();
        if (wr < 3) gives_down = active && bt.vy && s_fs[tid + 32] == fsel && s_o00[tid + 32] == o00 + w;
        if (wr > 0) takes_up = active && s_fs[tid - 32] == fsel && s_fs[tid - 32] >= 0 && s_o00[tid - 32] + w == o00 &&
                               s_o00[tid - 32] >= 0;
    }

    constexpr int CB = 4;
    // block-uniform 64-bit bases + 32-bit per-thread offsets (one integer add per access)
    const unsigned uhw = (unsigned)hw, upix = (unsigned)pix;
    const float* tbase = p.tgt + (size_t)b * C * hw;
    const float* sbb = sb + (size_t)b * C * hw;
    float* dtb = (kGradFeat && p.d_tgt) ? p.d_tgt + (size_t)b * C * hw : nullptr;
    float* dsbb = (kGradFeat && dsb) ? dsb + (size_t)b * C * hw : nullptr;
    float gix = 0.f, giy = 0.f;
    for (int c0 = 0; c0 < C; c0 += CB) {
        float t[CB], v[CB][4];
#pragma unroll
        for (int j = 0; j < CB; ++j) {
            const unsigned co = (unsigned)min(c0 + j, C - 1) * uhw;
            t[j] = active ? __ldg(tbase + (co + upix)) : 0.f;
            v[j][0] = active ? __ldg(sbb + (co + (unsigned)o00)) : 0.f;
            v[j][1] = active ? __ldg(sbb + (co + (unsigned)o01)) : 0.f;
            v[j][2] = active ? __ldg(sbb + (co + (unsigned)o10)) : 0.f;
            v[j][3] = active ? __ldg(sbb + (co + (unsigned)o11)) : 0.f;
        }
        float topv[CB], botv[CB], etop[CB], ebot[CB];
#pragma unroll
        for (int j = 0; j < CB; ++j) {
            topv[j] = botv[j] = etop[j] = ebot[j] = 0.f;
            if (c0 + j < C) {                                             // uniform across the warp
                const unsigned co = (unsigned)(c0 + j) * uhw;
                // zero weight <=> clamped tap: same arithmetic as bilin_sample_grad with the tap skipped
                const float v00 = v[j][0], v01 = bt.vx ? v[j][1] : 0.f, v10 = bt.vy ? v[j][2] : 0.f;
                const float v11 = (bt.vx && bt.vy) ? v[j][3] : 0.f;
                const float val = v00 * bt.nw + v01 * bt.ne + v10 * bt.sw + v11 * bt.se;
                const float dix = -v00 * bt.ey + v01 * bt.ey - v10 * bt.ay + v11 * bt.ay;
                const float diy = -v00 * bt.ex - v01 * bt.ax + v10 * bt.ex + v11 * bt.ax;
                const float df = val - t[j];
                const float gvv = up * df * rsqrt_approx(df * df + kL1Eps2);      // d loss / d warped value
                gix += gvv * dix;
                giy += gvv * diy;
                if (kGradFeat) {
                    if (dtb && active) dtb[co + upix] = -gvv;
                    if (dsbb) {
                        float top = gvv * bt.nw, bot = gvv * bt.sw;
                        etop[j] = gvv * bt.ne;
                        ebot[j] = gvv * bt.se;
                        const float in_top = __shfl_up_sync(0xffffffffu, etop[j], 1);
                        const float in_bot = __shfl_up_sync(0xffffffffu, ebot[j], 1);
                        if (absorbs) {
                            top += in_top;
                            bot += in_bot;
                        }
                        topv[j] = top;
                        botv[j] = bot;
                    }
                }
            }
        }
        if (kGradFeat && dsbb) {                  // dsbb is non-null for every lane or for none
            __syncthreads();                       // the previous batch's s_bot has been consumed
#pragma unroll
            for (int j = 0; j < CB; ++j) s_bot[j][tid] = gives_down ? botv[j] : 0.f;
            __syncthreads();
            if (active) {
#pragma unroll
                for (int j = 0; j < CB; ++j) {
                    if (c0 + j < C) {
                        const unsigned co = (unsigned)(c0 + j) * uhw;
                        const float top = topv[j] + (takes_up ? s_bot[j][tid - 32] : 0.f);
                        atomicAdd(dsbb + (co + (unsigned)o00), top);
                        if (bt.vy && !gives_down) atomicAdd(dsbb + (co + (unsigned)o10), botv[j]);
                        if (!absorbed && bt.vx) {
                            atomicAdd(dsbb + (co + (unsigned)o01), etop[j]);
                            if (bt.vy) atomicAdd(dsbb + (co + (unsigned)o11), ebot[j]);
                        }
                    }
                }
            }
        }
    }
    float aP[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) aP[k] = 0.f;
    if (active) {
        const float gu = gix * pr.mx, gv = giy * pr.my;
        const float rz = __frcp_rn(pr.z);
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


// ---------------------------------------------------------------------------------------------------------------
// Backward with trainable features, bucketed: grid_sample's d_src scatter as a GATHER.
//
// The scatter pattern (which source pixels a target pixel touches, with which bilinear weights) is the same for all C
// channels, and a global reduction costs ~1.3 issue cycles per LANE on B200 (REDG, spread addresses) -- 2..4 of them per
// pixel and channel bounded the atomic kernel above at ~150 us for the bench shape.  Here the per-pixel kernel
//   * writes g[c] = d loss / d warped value channel-LAST into G[b][pixel][C] (one 16-byte store per 4 channels), and
//   * registers its (up to) four taps in the bucket of the source pixel they touch: (target pixel, weight), one integer
//     atomic per tap instead of one float reduction per tap AND channel;
// then feat_gather_kernel walks the source pixels: 16 lanes x float4 = 64 channels read the G rows of the registered taps
// (256 contiguous bytes each), accumulate in registers, and each warp writes its NCHW d_src tile through a shared-memory
// transpose with full 128-byte lines -- no memset of d_src, no float atomics.  Buckets hold kFeatBucketCap taps; the rare
// excess goes to an overflow list that feat_overflow_kernel adds with atomics afterwards.
constexpr int kBucketRows = 4;     // CTA = 32 columns x 4 rows, one warp per row (8 rows: same time, 16 rows: 25 % slower)
// min-blocks = 1 on purpose: with the bare thread bound ptxas settles for 72 registers and a shallow load schedule (143 us);
// told that one resident CTA is acceptable it uses 125 registers, keeps a whole channel batch of gathers in flight and the
// kernel takes 134 us at 4 CTAs/SM (measured)
__global__ void __launch_bounds__(kBucketRows * 32, 1) feat_bwd_bucket_kernel(const FeatDev p) {
    __shared__ float s_cam[TDL_MAX_SRC * 12 + 9];
    __shared__ float s_dP[TDL_MAX_SRC * 12];
    const int tid = threadIdx.x, lane = tid & 31, wr = tid >> 5;
    const int b = blockIdx.z;
    const int h = p.h, w = p.w, C = p.C, S = p.S;
    const size_t hw = (size_t)h * w;
    if (tid < S * 12) {
        s_cam[tid] = __ldg(p.P + (size_t)b * S * 12 + tid);
        s_dP[tid] = 0.f;
    }
    if (tid >= 64 && tid < 64 + 9) s_cam[TDL_MAX_SRC * 12 + tid - 64] = __ldg(p.invK + (size_t)b * 9 + tid - 64);
    __syncthreads();
    const int xq = blockIdx.x * 32 + lane, yq = blockIdx.y * kBucketRows + wr;
    const bool active = xq < w && yq < h;
    const int y = active ? yq : 0, x = active ? xq : 0;
    const int pix = y * w + x;
    const int fsel = active ? p.argmin[(size_t)b * hw + pix] : -1;
    const DepthParams dp{p.min_disp, p.range};
    const UpTap ut = up_tap(y, x, p.sy, p.sx, p.dh, p.dw);
    const Geo g = backproject(up_value(p.disp + (size_t)b * p.dh * p.dw, p.dw, ut), dp, s_cam + TDL_MAX_SRC * 12, x, y);
    const float* Pf = s_cam + max(fsel, 0) * 12;
    const Proj pr = project<true>(g, Pf, h, w, p.align_corners);
    const Bilin bt = bilin_taps(pr.ix, pr.iy, h, w);
    const float up = active ? __ldg(p.dloss) * p.coef / ((float)p.Bnorm * (float)h * (float)w) / (float)C : 0.f;
    const float* sb = p.src[0];
#pragma unroll
    for (int f = 1; f < TDL_MAX_SRC; ++f)
        if (f == fsel) sb = p.src[f];
    const int x1 = min(bt.x0 + 1, w - 1), y1 = min(bt.y0 + 1, h - 1);
    const int o00 = bt.y0 * w + bt.x0, o01 = bt.y0 * w + x1, o10 = y1 * w + bt.x0, o11 = y1 * w + x1;

    if (active) {                                            // register the taps (channel independent)
        const int fb = fsel * p.B + b;
        int* cnt = p.bk_cnt + (size_t)fb * hw;
        int2* ent = p.bk_ent + (size_t)fb * hw * kFeatBucketCap;
        auto reg = [&](int o, float wgt) {
            if (wgt != 0.f) {                                // a zero weight contributes exactly 0
                const int slot = atomicAdd(cnt + o, 1);
                if (slot < kFeatBucketCap) {
                    ent[(size_t)o * kFeatBucketCap + slot] = make_int2(pix, __float_as_int(wgt));
                } else {
                    // per-image counter and list: one global counter serialised ~270 k same-address atomics at batch 64
                    // (warp-aggregating the counter instead measured 38 us SLOWER at batch 8)
                    const int k = atomicAdd(p.ov_cnt + b, 1);
                    p.ov_ent[(size_t)b * 4 * hw + k] = make_int4(fsel, o, pix, __float_as_int(wgt));
                }
            }
        };
        reg(o00, bt.nw);
        if (bt.vx) reg(o01, bt.ne);
        if (bt.vy) reg(o10, bt.sw);
        if (bt.vx && bt.vy) reg(o11, bt.se);
    }

    constexpr int CB = 4;                                    // == the float4 written to G; C % 4 == 0 on this path
    const unsigned uhw = (unsigned)hw, upix = (unsigned)pix;
    const float* tbase = p.tgt + (size_t)b * C * hw;
    const float* sbb = sb + (size_t)b * C * hw;
    float* dtb = p.d_tgt ? p.d_tgt + (size_t)b * C * hw : nullptr;
    float* Gp = p.G + ((size_t)b * hw + pix) * C;
    float gix = 0.f, giy = 0.f;
    for (int c0 = 0; c0 < C; c0 += CB) {
        float t[CB], v[CB][4];
#pragma unroll
        for (int j = 0; j < CB; ++j) {
            const unsigned co = (unsigned)(c0 + j) * uhw;
            t[j] = active ? __ldcs(tbase + (co + upix)) : 0.f;
            v[j][0] = active ? __ldg(sbb + (co + (unsigned)o00)) : 0.f;
            v[j][1] = active ? __ldg(sbb + (co + (unsigned)o01)) : 0.f;
            v[j][2] = active ? __ldg(sbb + (co + (unsigned)o10)) : 0.f;
            v[j][3] = active ? __ldg(sbb + (co + (unsigned)o11)) : 0.f;
        }
        float gq[CB];
#pragma unroll
        for (int j = 0; j < CB; ++j) {
            const unsigned co = (unsigned)(c0 + j) * uhw;
            const float v00 = v[j][0], v01 = bt.vx ? v[j][1] : 0.f, v10 = bt.vy ? v[j][2] : 0.f;
            const float v11 = (bt.vx && bt.vy) ? v[j][3] : 0.f;
            const float val = v00 * bt.nw + v01 * bt.ne + v10 * bt.sw + v11 * bt.se;
            const float dix = -v00 * bt.ey + v01 * bt.ey - v10 * bt.ay + v11 * bt.ay;
            const float diy = -v00 * bt.ex - v01 * bt.ax + v10 * bt.ex + v11 * bt.ax;
            const float df = val - t[j];
            const float gvv = up * df * rsqrt_approx(df * df + kL1Eps2);      // d loss / d warped value
            gix += gvv * dix;
            giy += gvv * diy;
            gq[j] = gvv;
            if (dtb && active) __stcs(dtb + (co + upix), -gvv);
        }
        if (active) *reinterpret_cast<float4*>(Gp + c0) = make_float4(gq[0], gq[1], gq[2], gq[3]);
    }
    float aP[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) aP[k] = 0.f;
    if (active) {
        const float gu = gix * pr.mx, gv = giy * pr.my;
        const float rz = __frcp_rn(pr.z);
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
            dd[pix] = gdisp;
        } else {
            const float hy = 1.f - ut.ly, hx = 1.f - ut.lx;
            atomicAdd(dd + (size_t)ut.y0 * p.dw + ut.x0, hy * hx * gdisp);
            atomicAdd(dd + (size_t)ut.y0 * p.dw + ut.x1, hy * ut.lx * gdisp);
            atomicAdd(dd + (size_t)ut.y1 * p.dw + ut.x0, ut.ly * hx * gdisp);
            atomicAdd(dd + (size_t)ut.y1 * p.dw + ut.x1, ut.ly * ut.lx * gdisp);
        }
    }
    for (int f = 0; f < S; ++f) {
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            const float v = warp_sum((active && f == fsel) ? aP[k] : 0.f);
            if (lane == 0 && v != 0.f) atomicAdd(&s_dP[f * 12 + k], v);
        }
    }
    __syncthreads();
    if (tid < S * 12 && s_dP[tid] != 0.f) atomicAdd(p.dP_acc + (size_t)b * S * 12 + tid, s_dP[tid]);
}

constexpr int kGatherWarps = 4;     // warps per CTA; every warp owns one tile and never synchronises with the others
constexpr int kGatherPix = 32;      // source pixels (consecutive in the plane) per tile = one 128-byte line per channel
constexpr int kGatherBatch = 8;     // pipeline depth in steps; a step = one G row per HALF-warp, i.e. 16 rows in flight per warp
constexpr int kGatherSub = 16 * kFeatBucketCap + kGatherBatch;   // list capacity of a half-warp (16 pixels) + read-ahead padding

// One WARP per tile of 32 consecutive source pixels, no CTA-level synchronisation.  ncu (profiles/) shows this kernel bound by
// L1/shared-memory wavefronts (82 % busy in its first version), so every shared-memory access below is conflict-free and
// the per-tap list entry is ONE 8-byte word.
//   1. lane L reads the bucket (size + 8 entries, 4 x 16 B) of pixel o0+L; a warp scan compacts the registered taps
//      into two shared lists (pixels 0-15 / 16-31 of the tile).  Entry = (G row offset in 16-byte units << 5 | last-tap
//      flag << 4 | pixel-in-half, weight); lists are padded with zero-weight taps to a common multiple of the depth;
//   2. each half-warp walks its list: 16 lanes x float4 read one G row (64 channels = 256 contiguous bytes) per step
//      through a software pipeline of depth kGatherBatch;
//   3. the taps of a pixel are consecutive: they accumulate in registers and go to the per-warp shared tile at the pixel's
//      last tap.  Tile cell (channel c, pixel q) lives at word c*32 + (q ^ (c >> 2)): the writes (16 lanes = 16 channel
//      quads, fixed q) and the reads (32 lanes = 32 pixels, fixed c) both touch 32 distinct banks;
//   4. the tile leaves as 64 full 128-byte lines of the NCHW gradient.
__global__ void __launch_bounds__(kGatherWarps * 32) feat_gather_kernel(const FeatDev p) {
    __shared__ float s_t[kGatherWarps][64 * kGatherPix];
    __shared__ int2 s_l[kGatherWarps][2 * kGatherSub];
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5, half = lane >> 4, l16 = lane & 15;
    const int fb = blockIdx.y;                               // frame * B + b
    const int f = fb / p.B, b = fb - f * p.B;
    const int hw = p.h * p.w, C = p.C;
    if (blockIdx.x == 0 && blockIdx.y == 0)                  // dP of the per-pixel kernel: scratch accumulator -> output
        for (int i = threadIdx.x; i < p.B * p.S * 12; i += kGatherWarps * 32) p.dP[i] = p.dP_acc[i];
    const int o0 = (blockIdx.x * kGatherWarps + wq) * kGatherPix;
    if (o0 >= hw) return;
    float* dst = p.d_src[0];
#pragma unroll
    for (int k = 1; k < TDL_MAX_SRC; ++k)
        if (k == f) dst = p.d_src[k];
    dst += (size_t)b * C * hw;
    const float4* Gb = reinterpret_cast<const float4*>(p.G + (size_t)b * hw * C);
    float* st = s_t[wq];
    int2* sl = s_l[wq] + half * kGatherSub;
    // 1. my pixel's bucket -> my half's list
    const int o = min(o0 + lane, hw - 1);
    const int4* e4 = reinterpret_cast<const int4*>(p.bk_ent + ((size_t)fb * hw + o) * kFeatBucketCap);
    const int n = (o0 + lane < hw) ? min(__ldg(p.bk_cnt + (size_t)fb * hw + o), kFeatBucketCap) : 0;
    int4 e[kFeatBucketCap / 2];
#pragma unroll
    for (int k = 0; k < kFeatBucketCap / 2; ++k) e[k] = __ldg(e4 + k);
    int off = n;                                             // inclusive scan inside each half
#pragma unroll
    for (int d = 1; d < 16; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, off, d);
        if (l16 >= d) off += v;
    }
    const int Th = __shfl_sync(0xffffffffu, off, 15, 16);    // taps of my half
    const int Tmax = max(__shfl_sync(0xffffffffu, Th, 0), __shfl_sync(0xffffffffu, Th, 16));
    const int Tpad = (Tmax + kGatherBatch - 1) / kGatherBatch * kGatherBatch;
    const unsigned empty = __ballot_sync(0xffffffffu, n == 0);
    off -= n;
    const int row16 = C >> 2;                                // G row length in 16-byte units
#pragma unroll
    for (int k = 0; k < kFeatBucketCap / 2; ++k) {
        if (2 * k < n) sl[off + 2 * k] = make_int2(((e[k].x * row16) << 5) | (2 * k == n - 1 ? 16 : 0) | l16, e[k].y);
        if (2 * k + 1 < n) sl[off + 2 * k + 1] = make_int2(((e[k].z * row16) << 5) | (2 * k + 1 == n - 1 ? 16 : 0) | l16, e[k].w);
    }
    for (int t = Th + l16; t < Tpad + kGatherBatch; t += 16) sl[t] = make_int2(0, 0);   // padding + read-ahead: row 0, weight 0
    const unsigned uhw = (unsigned)hw;
    for (int cc = 0; cc < C; cc += 64) {
        const int c = min(cc + 4 * l16, C - 4);              // lanes past C redo the last quad (not stored)
        const float4* Gc = Gb + (c >> 2);
        float* col = st + (4 * l16) * kGatherPix;            // my four channel rows of the tile; (4*l16 + j) >> 2 == l16
        __syncwarp();                                        // lists written / previous chunk's lines read
        for (unsigned m = empty; m; m &= m - 1) {            // pixels without taps
            const int q = __ffs(m) - 1;
            if ((q >> 4) == half) {
#pragma unroll
                for (int j = 0; j < 4; ++j) col[j * kGatherPix + (q ^ l16)] = 0.f;
            }
        }
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int2 ev[kGatherBatch];
        float4 g4[kGatherBatch];
#pragma unroll
        for (int k = 0; k < kGatherBatch; ++k) {
            ev[k] = sl[k];
            g4[k] = __ldg(Gc + ((unsigned)ev[k].x >> 5));
        }
        for (int t0 = 0; t0 < Tpad; t0 += kGatherBatch) {
#pragma unroll
            for (int k = 0; k < kGatherBatch; ++k) {
                const float wgt = __int_as_float(ev[k].y);
                const int meta = ev[k].x;
                acc.x = fmaf(wgt, g4[k].x, acc.x);
                acc.y = fmaf(wgt, g4[k].y, acc.y);
                acc.z = fmaf(wgt, g4[k].z, acc.z);
                acc.w = fmaf(wgt, g4[k].w, acc.w);
                ev[k] = sl[t0 + kGatherBatch + k];
                g4[k] = __ldg(Gc + ((unsigned)ev[k].x >> 5));
                if (meta & 16) {                             // last tap of pixel half*16 + (meta & 15): uniform in the half-warp
                    float* cq = col + ((half * 16 + (meta & 15)) ^ l16);
                    cq[0] = acc.x;
                    cq[kGatherPix] = acc.y;
                    cq[2 * kGatherPix] = acc.z;
                    cq[3 * kGatherPix] = acc.w;
                    acc = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
        __syncwarp();
        if (o0 + lane < hw) {
            const int nch = min(64, C - cc);
            float* d0 = dst + (size_t)cc * hw + o0 + lane;
            int ch = 0;
            for (; ch + 8 <= nch; ch += 8) {                 // ch % 4 == 0: the two quads of this group are ch>>2 and (ch>>2)+1
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = st[(ch + j) * kGatherPix + (lane ^ ((ch + j) >> 2))];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    *d0 = v[j];
                    d0 += uhw;
                }
            }
            for (; ch < nch; ++ch, d0 += uhw) *d0 = st[ch * kGatherPix + (lane ^ (ch >> 2))];
        }
    }
}

__global__ void __launch_bounds__(256) feat_overflow_kernel(const FeatDev p) {
    const int b = blockIdx.y;
    const int n = p.ov_cnt[b];
    const int ngrp = (int)(gridDim.x * blockDim.x) >> 4;
    const int l16 = threadIdx.x & 15;
    const int hw = p.h * p.w, C = p.C;
    const int4* list = p.ov_ent + (size_t)b * 4 * hw;
    for (int i = (int)(blockIdx.x * blockDim.x + threadIdx.x) >> 4; i < n; i += ngrp) {
        const int4 e = list[i];                              // (frame, source pixel, target pixel, weight)
        float* dst = p.d_src[0];
#pragma unroll
        for (int k = 1; k < TDL_MAX_SRC; ++k)
            if (k == e.x) dst = p.d_src[k];
        dst += (size_t)b * C * hw + e.y;
        const float wgt = __int_as_float(e.w);
        const float* Gr = p.G + ((size_t)b * hw + e.z) * C;
        for (int c = 4 * l16; c < C; c += 64) {
            const float4 gv = *reinterpret_cast<const float4*>(Gr + c);
            atomicAdd(dst + (size_t)(c + 0) * hw, wgt * gv.x);
            atomicAdd(dst + (size_t)(c + 1) * hw, wgt * gv.y);
            atomicAdd(dst + (size_t)(c + 2) * hw, wgt * gv.z);
            atomicAdd(dst + (size_t)(c + 3) * hw, wgt * gv.w);
        }
    }
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
    dim3 grid((p.w + 31) / 32, (p.h + 3) / 4, p.B);
    if (p.G)
        feat_bwd_bucket_kernel<<<dim3((p.w + 31) / 32, (p.h + kBucketRows - 1) / kBucketRows, p.B), kBucketRows * 32, 0, st>>>(p);
    else if (p.d_tgt || p.d_src[0])
        feat_bwd_kernel<true><<<grid, kFeatNT, 0, st>>>(p);
    else
        feat_bwd_kernel<false><<<grid, kFeatNT, 0, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_feat_bwd_gather(const FeatDev& p, cudaStream_t st) {
    const int tiles = (p.h * p.w + kGatherPix - 1) / kGatherPix;
    dim3 grid((tiles + kGatherWarps - 1) / kGatherWarps, p.S * p.B);
    feat_gather_kernel<<<grid, kGatherWarps * 32, 0, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_feat_bwd_overflow(const FeatDev& p, cudaStream_t st) {
    feat_overflow_kernel<<<dim3((148 * 8 + p.B - 1) / p.B, p.B), 256, 0, st>>>(p);
    return cudaGetLastError();
}

}  // namespace tdl
