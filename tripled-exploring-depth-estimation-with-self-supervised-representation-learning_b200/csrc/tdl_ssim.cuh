// Register micro-tile SSIM + robust-L1 for the photometric kernels (sm_100a).
//
// One thread owns a patch of kPC = 2 adjacent columns x kPR = 4 rows of pixels and streams the kPR + 2 rows of the
// 3x3 windows that cover it.  Per row it loads four values per image (left neighbour, the two patch columns as one
// 8-byte word, right neighbour), forms the horizontal 3-sums of x, x*x and x*y for both columns with the shared middle
// pair (3 adds / 1 mul + 3 fma per row and statistic instead of 2 x 2 adds + 3 muls), and combines three consecutive
// rows with the shared vertical pair (1.5 adds per window).  That is ~13 instructions per pixel, image and channel for
// the three window sums, against 9 loads + 6 multiplies + 24 adds (39) for the one-column strip of round 1.
//
// SSIM is evaluated on the raw 9-sums: with Sx = 9 mu_x, Sxx = 9 E[x^2], ... numerator and denominator of
// mono/model/mono_fm/layers.py:104-105 are both scaled by 81^2,
//     n = (2 Sx Sy + 81 C1) (18 Sxy - 2 Sx Sy + 81 C2),   d = (Sx^2 + Sy^2 + 81 C1) (9 Sxx - Sx^2 + 9 Syy - Sy^2 + 81 C2),
// so the three divisions by 9 disappear; the target-only parts (Sy and Qy = 9 Syy - Sy^2 + 81 C2) are computed once per
// tile and shared by every source frame and scale.  (1 - n/d)/2 clamped to [0,1] is one reciprocal, one subtraction
// and a saturating multiply.  The arithmetic is fp32 throughout but no longer follows ATen's summation order: the
// per-pixel difference to the reference is the reference's own rounding noise (~1e-5 absolute on smooth content,
// tests/test_gpu_parity.py TIE_ATOL), zero-mean, and the loss scalars stay within 1e-5 relative.
#pragma once

#include "tdl_common.cuh"

namespace tdl {

constexpr int kPC = 2;                     // patch columns per thread
constexpr int kPR = 4;                     // patch rows per thread
constexpr int kPP = kPC * kPR;             // pixels per thread
constexpr float kK1 = 81.f * kSsimC1;      // 81 C1
constexpr float kK2 = 81.f * kSsimC2;      // 81 C2

TDL_DEV float mul_sat(float a, float b) {
    float r;
    asm("mul.sat.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

struct Row4 {                              // left neighbour, two patch columns, right neighbour of one image row
    float v0, v1, v2, v3;
};

// `row` points at the patch's first column in a staged plane (8-byte aligned); row[-1] is its left neighbour -- for the
// first patch of a tile row that address is the last (spare) cell of the previous staged row, where the kernels park
// the tile's left-halo column (see tdl_photo2.cu), so every lane runs the same three loads.
TDL_DEV Row4 load_row(const float* __restrict__ row) {
    Row4 r;
    const float2 m = *reinterpret_cast<const float2*>(row);
    r.v0 = row[-1];
    r.v1 = m.x;
    r.v2 = m.y;
    r.v3 = row[2];
    return r;
}

struct H2 {                                // horizontal 3-sums of the two patch columns
    float a, b;
};
TDL_DEV H2 hsum(const Row4& r) {
    const float t = r.v1 + r.v2;
    return H2{r.v0 + t, t + r.v3};
}
TDL_DEV H2 hsum_prod(const Row4& p, const Row4& q) {      // sums of p*q
    const float t = fmaf(p.v1, q.v1, p.v2 * q.v2);
    return H2{fmaf(p.v0, q.v0, t), fmaf(p.v3, q.v3, t)};
}
TDL_DEV H2 operator+(const H2& x, const H2& y) { return H2{x.a + y.a, x.b + y.b}; }

// ---- target-only statistics of the thread's patch: (Sy, Qy) per pixel, Qy = 9 Syy - Sy^2 + 81 C2
// ys: plane pointer at (patch row 0 - 1, patch column 0); PITCH: floats per staged row
template <int PITCH>
TDL_DEV void patch_target_stats(const float* __restrict__ ys, float2 out[kPP]) {
    H2 s[kPR + 2], q[kPR + 2];
#pragma unroll
    for (int j = 0; j < kPR + 2; ++j) {
        const Row4 y = load_row(ys + j * PITCH);
        s[j] = hsum(y);
        q[j] = hsum_prod(y, y);
    }
#pragma unroll
    for (int i = 0; i < kPR; i += 2) {     // windows i, i+1 share rows i+1, i+2
        const H2 ps = s[i + 1] + s[i + 2], pq = q[i + 1] + q[i + 2];
        const H2 s0 = s[i] + ps, s1 = ps + s[i + 3], q0 = q[i] + pq, q1 = pq + q[i + 3];
        out[i * kPC + 0] = make_float2(s0.a, fmaf(9.f, q0.a, fmaf(-s0.a, s0.a, kK2)));
        out[i * kPC + 1] = make_float2(s0.b, fmaf(9.f, q0.b, fmaf(-s0.b, s0.b, kK2)));
        out[(i + 1) * kPC + 0] = make_float2(s1.a, fmaf(9.f, q1.a, fmaf(-s1.a, s1.a, kK2)));
        out[(i + 1) * kPC + 1] = make_float2(s1.b, fmaf(9.f, q1.b, fmaf(-s1.b, s1.b, kK2)));
    }
}

// (1 - SSIM)/2 clamped to [0,1] from the window sums of the prediction (Sx, Sxx, Sxy) and the target's (Sy, Qy, Ay);
// Ay = fl(Sy^2 + 81 C1), Qy = fl(9 Syy + fl(81 C2 - Sy^2)).
// Two properties of the reference's arithmetic are kept by construction:
//  * SSIM(x, x) == 0 EXACTLY (auto-masking relies on it for static scenes): numerator and denominator are evaluated
//    by the same nest of fused multiply-adds with (Sx, Sy, Sxy) in place of (Sx, Sx, Sxx) / (Sy, Sy, Syy), so that for
//    identical windows n1 == d1 and n2 == d2 bit for bit;
//  * no systematic rounding offset: the constants 81 C1 / 81 C2 are only ever added to EXACT products inside an fma.
//    Adding them to an already-rounded value of magnitude ~100 (e.g. 2 fl(Sx Sy) + K1) is off by the same fraction of
//    an ulp for every pixel, which showed up as +1.4e-6 on the mean SSIM term (5e-5 of the loss) when first measured.
TDL_DEV float ssim_half(float Sx, float Sxx, float Sxy, float Sy, float Qy, float Ay) {
    const float n1 = fmaf(Sx, Sy, fmaf(Sx, Sy, kK1));                       // 2 Sx Sy + K1        (d1 with Sy -> Sx, Sx -> Sy)
    const float d1 = fmaf(Sx, Sx, Ay);                                      // Sx^2 + Sy^2 + K1
    const float qn = fmaf(9.f, Sxy, fmaf(-Sx, Sy, kK2));                    // 9 Sxy - Sx Sy + K2  (Qy with y -> x in one factor)
    const float n2 = fmaf(-Sx, Sy, fmaf(9.f, Sxy, qn));                     // 18 Sxy - 2 Sx Sy + K2
    const float d2 = fmaf(-Sx, Sx, fmaf(9.f, Sxx, Qy));                     // 9 Sxx - Sx^2 + 9 Syy - Sy^2 + K2
    // (__fmul_rn / __fsub_rn: nvcc would otherwise contract d1*d2 - n into fma(d1, d2, -n), whose result for identical
    //  windows is the rounding error of n instead of 0 -- measured as +3e-5 on the loss of a static scene)
    const float n = __fmul_rn(n1, n2), d = __fmul_rn(d1, d2);
    return mul_sat(__fsub_rn(d, n), 0.5f * rcp_approx(d));
}

TDL_DEV float robust_l1_fast(float x, float y) {          // sqrt((y-x)^2 + eps^2) = t * rsqrt(t)
    const float df = y - x;
    const float t = fmaf(df, df, kL1Eps2);
    return t * rsqrt_approx(t);
}

// One channel of NF prediction images against the target, for the thread's patch.  xs[f] / ys point at
// (patch row 0 - 1, patch column 0) of the staged planes.
// st: the patch's (Sy, Qy) pairs.  Accumulates sum_c SSIM-half into sa[f][p] and sum_c robust-L1 into la[f][p].
template <int NF, int PITCH>
TDL_DEV void patch_ssim_l1(const float* const (&xs)[NF], const float* __restrict__ ys, const float2 (&st)[kPP],
                           float (&sa)[NF][kPP], float (&la)[NF][kPP]) {
    // rows are streamed; three rows of horizontal sums are alive per image (see file header)
    H2 hx[NF][3], hxx[NF][3], hxy[NF][3];
    H2 px[NF], pxx[NF], pxy[NF];                         // vertical pair sums (rows i+1, i+2 of an even window i)
#pragma unroll
    for (int j = 0; j < kPR + 2; ++j) {
        const Row4 y = load_row(ys + j * PITCH);
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            const Row4 x = load_row(xs[f] + j * PITCH);
            hx[f][j % 3] = hsum(x);
            hxx[f][j % 3] = hsum_prod(x, x);
            hxy[f][j % 3] = hsum_prod(x, y);
            if (j >= 1 && j <= kPR) {                      // centre pixels of window row j - 1
                la[f][(j - 1) * kPC + 0] += robust_l1_fast(x.v1, y.v1);
                la[f][(j - 1) * kPC + 1] += robust_l1_fast(x.v2, y.v2);
            }
            if (j >= 2) {
                const int i = j - 2;                       // window row completed by this image row
                H2 Sx, Sxx, Sxy;
                if ((i & 1) == 0) {                        // even window: rows i, (i+1, i+2)
                    px[f] = hx[f][(j - 1) % 3] + hx[f][j % 3];
                    pxx[f] = hxx[f][(j - 1) % 3] + hxx[f][j % 3];
                    pxy[f] = hxy[f][(j - 1) % 3] + hxy[f][j % 3];
                    Sx = hx[f][(j - 2) % 3] + px[f];
                    Sxx = hxx[f][(j - 2) % 3] + pxx[f];
                    Sxy = hxy[f][(j - 2) % 3] + pxy[f];
                } else {                                   // odd window: (rows i, i+1) = the previous pair, + row i+2
                    Sx = px[f] + hx[f][j % 3];
                    Sxx = pxx[f] + hxx[f][j % 3];
                    Sxy = pxy[f] + hxy[f][j % 3];
                }
                const float2 s0 = st[i * kPC + 0], s1 = st[i * kPC + 1];
                sa[f][i * kPC + 0] += ssim_half(Sx.a, Sxx.a, Sxy.a, s0.x, s0.y, fmaf(s0.x, s0.x, kK1));
                sa[f][i * kPC + 1] += ssim_half(Sx.b, Sxx.b, Sxy.b, s1.x, s1.y, fmaf(s1.x, s1.x, kK1));
            }
        }
    }
}

// rho = 0.85 * mean_c SSIM-half + 0.15 * mean_c L1 (mono/model/mono_fm/net.py:63-67)
TDL_DEV float rho_from_sums(float sa, float la) { return fmaf(sa, 0.85f / 3.f, la * (0.15f / 3.f)); }

}  // namespace tdl
