// Device helpers shared by the fused view-synthesis-loss kernels (sm_100a).
//
// Arithmetic follows the reference's fp32 operation order;
// where PyTorch evaluates an expression as separate element-wise kernels (no FMA
// contraction possible) the explicit __f*_rn intrinsics are used so that nvcc does
// not contract them either.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "tdl.h"

#define TDL_DEV __device__ __forceinline__

namespace tdl {

// python evaluates these in double; torch then rounds the scalar to fp32 once
constexpr float kSsimC1 = static_cast<float>(0.01 * 0.01);   // mono/model/mono_fm/layers.py:94
constexpr float kSsimC2 = static_cast<float>(0.03 * 0.03);   // mono/model/mono_fm/layers.py:95
constexpr float kL1Eps2 = static_cast<float>(1e-3 * 1e-3);   // mono/model/mono_fm/net.py:56 (eps ** 2)
constexpr float kProjEps = 1e-7f;          // mono/model/mono_fm/layers.py:65

// nn.AvgPool2d(3,1) divides the 9-sum by 9 (correctly rounded).  q = s*(1/9) followed by one
// Newton step reproduces the IEEE quotient at a third of the cost of __fdiv_rn; a bare multiply
// by fl(1/9) biases sigma = E[x^2] - mu^2 by ~2e-9, i.e. 2.5e-6 of C2 (measured on round 1).
TDL_DEV float div9(float s) {
    const float r9 = 1.f / 9.f;
    const float q = s * r9;
    return fmaf(fmaf(-9.f, q, s), r9, q);
}

// Correctly-rounded quotients without the slow-path check of __fdiv_rn (FCHK + branch + call, ~12
// instructions): q0 = n * rcp(d) refined by one Newton step on the residual.  For the operand ranges of
// this path (|d| in [1e-7, 1e4], no denormals, no overflow) the result equals the IEEE quotient; ncu (round 1)
// attributed 22 % of the forward's instructions to the 140 divisions per pixel.
TDL_DEV float rcp_approx(float d) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));      // MUFU.RCP, 1 ulp
    return r;
}
TDL_DEV float div_rn(float n, float d) {
    const float r = rcp_approx(d);
    const float q = n * r;
    return fmaf(fmaf(-d, q, n), r, q);
}
// sqrt(x) for x in [1e-7, 1e2]: rsqrt seed + one Newton step on the residual (correctly rounded except for
// rare last-bit ties); __fsqrt_rn's IEEE sequence was 5 % of the forward's instructions (line profile, r1)
TDL_DEV float rsqrt_approx(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));   // MUFU.RSQ alone (rsqrtf() adds denormal handling)
    return r;
}
TDL_DEV float sqrt_fast(float x) {
    const float r = rsqrt_approx(x);
    const float s = x * r;
    return fmaf(fmaf(-s, s, x), 0.5f * r, s);
}
// 1/d to ~1 ulp without the IEEE slow-path call of __frcp_rn (d in [1e-7, 1e4]): MUFU seed + one Newton step
TDL_DEV float rcp_newton(float d) {
    const float r = rcp_approx(d);
    return fmaf(fmaf(-d, r, 1.f), r, r);
}
// division by a constant whose reciprocal rc = fl(1/c) is known
TDL_DEV float div_const(float n, float c, float rc) {
    const float q = n * rc;
    return fmaf(fmaf(-c, q, n), rc, q);
}
TDL_DEV float div3(float s) { return div_const(s, 3.f, 1.f / 3.f); }

// ---------------------------------------------------------------- index helpers
// nn.ReflectionPad2d(1) index map: -1 -> 1, n -> n-2 (then clamped for partial tiles)
TDL_DEV int reflect1(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return max(0, min(i, n - 1));
}

// F.interpolate(mode="bilinear", align_corners=False) source index
// (ATen area_pixel_compute_source_index): src = scale*(dst+0.5)-0.5, clamped at 0.
TDL_DEV void up_index(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
    float src = scale * (static_cast<float>(dst) + 0.5f) - 0.5f;
    src = src < 0.f ? 0.f : src;
    i0 = static_cast<int>(src);
    i0 = min(i0, in_size - 1);
    i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
    l1 = src - static_cast<float>(i0);
}

struct UpTap {            // the four taps + weights of one up-sampled pixel
    int y0, y1, x0, x1;
    float ly, lx;
};

TDL_DEV UpTap up_tap(int y, int x, float sy, float sx, int h, int w) {
    UpTap t;
    up_index(y, sy, h, t.y0, t.y1, t.ly);
    up_index(x, sx, w, t.x0, t.x1, t.lx);
    return t;
}

TDL_DEV float up_value(const float* __restrict__ d, int w, const UpTap& t) {
    const float hy = 1.f - t.ly, hx = 1.f - t.lx;
    const float* r0 = d + (size_t)t.y0 * w;
    const float* r1 = d + (size_t)t.y1 * w;
    return hy * (hx * __ldg(r0 + t.x0) + t.lx * __ldg(r0 + t.x1)) +
           t.ly * (hx * __ldg(r1 + t.x0) + t.lx * __ldg(r1 + t.x1));
}

// ---------------------------------------------------------------- geometry
struct DepthParams {      // disp_to_depth constants, rounded to fp32 like torch does with python scalars
    float min_disp;       // 1/max_depth
    float range;          // 1/min_depth - 1/max_depth
};

struct Cam {              // per (image, source frame)
    float P[12];          // (K@T)[:3,:] row-major
};

struct Geo {              // per pixel, shared by all source frames
    float D;              // depth
    float r0, r1, r2;     // inv_K[:3,:3] @ (x,y,1)
    float X0, X1, X2;     // camera-space point
};

TDL_DEV Geo backproject(float disp, const DepthParams& dp, const float* iK, int x, int y) {
    Geo g;
    const float scaled = __fadd_rn(dp.min_disp, __fmul_rn(dp.range, disp));   // net.py:138
    g.D = div_rn(1.0f, scaled);                                               // net.py:139
    const float fx = static_cast<float>(x), fy = static_cast<float>(y);
    g.r0 = iK[0] * fx + iK[1] * fy + iK[2];                                   // layers.py:58
    g.r1 = iK[3] * fx + iK[4] * fy + iK[5];
    g.r2 = iK[6] * fx + iK[7] * fy + iK[8];
    g.X0 = __fmul_rn(g.D, g.r0);                                              // layers.py:59
    g.X1 = __fmul_rn(g.D, g.r1);
    g.X2 = __fmul_rn(g.D, g.r2);
    return g;
}

struct Proj {
    float ix, iy;         // clipped source pixel coordinates
    float u, v, z;        // projected pixel coords and (depth + eps)
    float mx, my;         // d ix / d u and d iy / d v (0 where the coordinate was clipped)
};

// image-size constants of Project / grid_sample: computed ONCE per kernel (the IEEE divisions compile to a call with a
// slow path; inside a loop with many live registers every such call spills them)
struct ProjConst {
    float wm1, hm1, rwm1, rhm1;   // W-1, H-1 and their reciprocals
    float Wf, Hf;
    float sx, sy;                 // d ix / d u, d iy / d v where the coordinate is not clipped
    int align_corners;
};
TDL_DEV ProjConst make_proj_const(int H, int W, int align_corners) {
    ProjConst c;
    c.wm1 = static_cast<float>(W - 1);
    c.hm1 = static_cast<float>(H - 1);
    c.rwm1 = 1.f / c.wm1;
    c.rhm1 = 1.f / c.hm1;
    c.Wf = static_cast<float>(W);
    c.Hf = static_cast<float>(H);
    c.sx = align_corners ? 1.f : c.Wf / c.wm1;
    c.sy = align_corners ? 1.f : c.Hf / c.hm1;
    c.align_corners = align_corners;
    return c;
}

template <bool kGrad>
TDL_DEV Proj project(const Geo& g, const float* P, const ProjConst& c) {
    Proj o;
    const float p0 = P[0] * g.X0 + P[1] * g.X1 + P[2] * g.X2 + P[3];          // layers.py:75
    const float p1 = P[4] * g.X0 + P[5] * g.X1 + P[6] * g.X2 + P[7];
    const float p2 = P[8] * g.X0 + P[9] * g.X1 + P[10] * g.X2 + P[11];
    o.z = __fadd_rn(p2, kProjEps);                                            // layers.py:76
    {
        const float rz = rcp_approx(o.z);                                     // both quotients share rcp(z)
        const float qu = p0 * rz, qv = p1 * rz;
        o.u = fmaf(fmaf(-o.z, qu, p0), rz, qu);
        o.v = fmaf(fmaf(-o.z, qv, p1), rz, qv);
    }
    const float gx = __fmul_rn(__fsub_rn(div_const(o.u, c.wm1, c.rwm1), 0.5f), 2.0f);   // layers.py:79-81
    const float gy = __fmul_rn(__fsub_rn(div_const(o.v, c.hm1, c.rhm1), 0.5f), 2.0f);
    float ix, iy;
    if (c.align_corners) {                                                    // grid_sampler_unnormalize
        ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.f), 0.5f), c.wm1);           // x/2 == x*0.5 exactly
        iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.f), 0.5f), c.hm1);
    } else {
        ix = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), c.Wf), 1.f), 0.5f);
        iy = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), c.Hf), 1.f), 0.5f);
    }
    if (kGrad) {
        // clip_coordinates_set_grad: zero gradient when ix <= 0 or ix >= size-1
        o.mx = (ix <= 0.f || ix >= c.wm1) ? 0.f : c.sx;
        o.my = (iy <= 0.f || iy >= c.hm1) ? 0.f : c.sy;
        if (!(ix == ix)) o.mx = 0.f;
        if (!(iy == iy)) o.my = 0.f;
    }
    // padding_mode="border": clip to [0, size-1]; NaN -> 0 keeps the gather in range
    ix = fminf(c.wm1, fmaxf(ix, 0.f));
    iy = fminf(c.hm1, fmaxf(iy, 0.f));
    o.ix = ix;
    o.iy = iy;
    return o;
}

template <bool kGrad>
TDL_DEV Proj project(const Geo& g, const float* P, int H, int W, int align_corners) {
    return project<kGrad>(g, P, make_proj_const(H, W, align_corners));
}

struct Bilin {            // ATen grid_sampler_2d bilinear taps
    int x0, y0;           // north-west corner (always inside the image after clipping)
    bool vx, vy;          // east column / south row inside the image
    float nw, ne, sw, se;
    float ax, ay;         // ix - x0, iy - y0   (for the coordinate gradient)
    float ex, ey;         // (x0+1) - ix, (y0+1) - iy
};

TDL_DEV Bilin bilin_taps(float ix, float iy, int H, int W) {
    Bilin t;
    const float fx = floorf(ix), fy = floorf(iy);
    t.x0 = static_cast<int>(fx);
    t.y0 = static_cast<int>(fy);
    t.vx = t.x0 + 1 <= W - 1;
    t.vy = t.y0 + 1 <= H - 1;
    const float ex = (fx + 1.f) - ix, ey = (fy + 1.f) - iy;
    t.ex = ex;
    t.ey = ey;
    t.ax = ix - fx;
    t.ay = iy - fy;
    t.nw = ex * ey;
    t.ne = t.ax * ey;
    t.sw = ex * t.ay;
    t.se = t.ax * t.ay;
    return t;
}

// one channel plane; returns the interpolated value.  The east / south taps are loaded from a clamped
// index: a clamped tap always has weight exactly 0 (ix == W-1 or iy == H-1 after border clipping), so it
// adds 0 like ATen's skipped out-of-range tap, without predicated loads or branches.
// Hides a pointer's provenance from the optimiser.  Left symbolic, "kernel-parameter base + uniform 64-bit image offset +
// per-thread index" is re-assembled for every access (add, add-with-carry, shift-add pair); an opaque per-thread base
// register makes each access one widening multiply-add of the 32-bit element index.
template <typename P>
TDL_DEV P* opaque(P* q) {
    asm volatile("" : "+l"(q));
    return q;
}

TDL_DEV float bilin_sample(const float* __restrict__ plane, int W, const Bilin& t) {
    const float* p = plane + (size_t)t.y0 * W + t.x0;
    const int dx = t.vx ? 1 : 0, dy = t.vy ? W : 0;
    float acc = __ldg(p) * t.nw;
    acc += __ldg(p + dx) * t.ne;
    acc += __ldg(p + dy) * t.sw;
    acc += __ldg(p + dy + dx) * t.se;
    return acc;
}

// value + d/d ix, d/d iy (un-scaled, un-clipped) of one channel plane
TDL_DEV float bilin_sample_grad(const float* __restrict__ plane, int W, const Bilin& t, float& dix, float& diy) {
    const float* p = plane + (size_t)t.y0 * W + t.x0;
    const float v00 = __ldg(p);
    const float v01 = t.vx ? __ldg(p + 1) : 0.f;
    const float v10 = t.vy ? __ldg(p + W) : 0.f;
    const float v11 = (t.vx && t.vy) ? __ldg(p + W + 1) : 0.f;
    const float ey = t.ey, ex = t.ex;
    // ATen grid_sampler_2d_backward: gix -= nw_val*(iy_se-iy) ... with out-of-range taps skipped
    dix = -v00 * ey + v01 * ey - v10 * t.ay + v11 * t.ay;
    diy = -v00 * ex - v01 * t.ax + v10 * ex + v11 * t.ax;
    return v00 * t.nw + v01 * t.ne + v10 * t.sw + v11 * t.se;
}

// ---------------------------------------------------------------- reductions
TDL_DEV float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sums 12 per-lane values over the warp with 16 shuffles instead of 60: at every butterfly step a lane keeps one half of
// the (remaining) values and hands the other half to its partner.  On return lane l (l < 12 after the value-splitting
// steps, see below) holds the total of value `slot`; returns that slot index, or -1 when the lane holds nothing.
TDL_DEV int warp_sum12(const float in[12], float& total) {
    const int lane = threadIdx.x & 31;
    float a[8];
    // step 1 (xor 16): 12 -> 6 values (+2 zero pads -> 8 slots of which 6 used)
    {
        const bool hi = lane & 16;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const float keep = hi ? in[6 + k] : in[k], give = hi ? in[k] : in[6 + k];
            a[k] = keep + __shfl_xor_sync(0xffffffffu, give, 16);
        }
    }
    // step 2 (xor 8): 6 -> 3
    float b3[3];
    {
        const bool hi = lane & 8;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float keep = hi ? a[3 + k] : a[k], give = hi ? a[k] : a[3 + k];
            b3[k] = keep + __shfl_xor_sync(0xffffffffu, give, 8);
        }
    }
    // steps 3-5 (xor 4, 2, 1): plain butterfly on the 3 remaining values
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; ++k) b3[k] += __shfl_xor_sync(0xffffffffu, b3[k], o);
    }
    // lane layout: bit4 selects {0..5 | 6..11}, bit3 selects {first | second} triple; lanes 0, 8, 16, 24 publish
    const int base = ((lane & 16) ? 6 : 0) + ((lane & 8) ? 3 : 0);
    const int sub = lane & 7;                   // lanes sub = 0,1,2 of each octet publish one value each
    total = sub == 0 ? b3[0] : (sub == 1 ? b3[1] : b3[2]);
    return sub < 3 ? base + sub : -1;
}

// Sums `v` over the block; result valid in thread 0.  `scratch` holds >= 32 floats.
TDL_DEV float block_sum(float v, float* scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? scratch[threadIdx.x] : 0.f;
    if (wid == 0) v = warp_sum(v);
    return v;
}

// ---------------------------------------------------------------- Philox4x32-10 -> N(0,1)
// Counter-based generator for the automask tie-break noise when the caller does not
// supply the reference's torch.randn draws (mono/model/mono_fm/net.py:94).
TDL_DEV uint4 philox4x32(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}

TDL_DEV float2 box_muller(uint32_t a, uint32_t b) {
    const float u1 = (static_cast<float>(a) + 1.0f) * 2.3283064365386963e-10f;   // (0,1]
    const float u2 = static_cast<float>(b) * 2.3283064365386963e-10f;
    const float r = sqrtf(-2.0f * __logf(u1));
    float s, c;
    __sincosf(6.283185307179586f * u2, &s, &c);
    return make_float2(r * c, r * s);
}

}  // namespace tdl
