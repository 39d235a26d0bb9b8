// Pose prologue of the view-synthesis loss: axis-angle + translation -> cam_T_cam.
//
//   transformation_from_parameters   mono/model/mono_fm/net.py:201-212
//   get_translation_matrix           mono/model/mono_fm/net.py:214-222
//   rot_from_axisangle               mono/model/mono_fm/net.py:224-253
//
// The reference spends ~40 element-wise launches (plus two `.cuda()` zero tensors) per source frame on 12 numbers
// per image; here it is one launch forward and one backward.  The forward follows the reference's fp32 operation
// order with explicit round-to-nearest intrinsics (no FMA contraction); the products with the constant 0 / 1 entries of
// the homogeneous matrices are exact, so T @ R and R^T @ T' are written out directly.  The backward differentiates
// the same expression with forward-mode dual numbers, one thread per (image, parameter): six tangents of a 4x4
// matrix are cheaper than a hand-derived adjoint is error-prone.
#include "tdl_common.cuh"
#include "tdl_internal.h"

namespace tdl {

namespace {

struct Dual {                       // value + derivative along one input parameter
    float v, d;
};
TDL_DEV Dual operator+(Dual a, Dual b) { return {a.v + b.v, a.d + b.d}; }
TDL_DEV Dual operator-(Dual a, Dual b) { return {a.v - b.v, a.d - b.d}; }
TDL_DEV Dual operator*(Dual a, Dual b) { return {a.v * b.v, a.d * b.v + a.v * b.d}; }
TDL_DEV Dual operator/(Dual a, Dual b) {
    const float q = a.v / b.v;
    return {q, (a.d - q * b.d) / b.v};
}
TDL_DEV Dual dneg(Dual a) { return {-a.v, -a.d}; }
TDL_DEV Dual dsqrt(Dual a) {
    const float s = sqrtf(a.v);
    return {s, s > 0.f ? a.d / (2.f * s) : 0.f};          // torch.norm's sub-gradient at 0 is 0
}
TDL_DEV Dual dcos(Dual a) { return {cosf(a.v), -sinf(a.v) * a.d}; }
TDL_DEV Dual dsin(Dual a) { return {sinf(a.v), cosf(a.v) * a.d}; }

// value type: plain float with the reference's un-fused operation order
struct Exact {
    float v;
};
TDL_DEV Exact operator+(Exact a, Exact b) { return {__fadd_rn(a.v, b.v)}; }
TDL_DEV Exact operator-(Exact a, Exact b) { return {__fsub_rn(a.v, b.v)}; }
TDL_DEV Exact operator*(Exact a, Exact b) { return {__fmul_rn(a.v, b.v)}; }
TDL_DEV Exact operator/(Exact a, Exact b) { return {__fdiv_rn(a.v, b.v)}; }
TDL_DEV Exact dneg(Exact a) { return {-a.v}; }
TDL_DEV Exact dsqrt(Exact a) { return {__fsqrt_rn(a.v)}; }
TDL_DEV Exact dcos(Exact a) { return {cosf(a.v)}; }
TDL_DEV Exact dsin(Exact a) { return {sinf(a.v)}; }

template <class T>
TDL_DEV T lit(float c);
template <>
TDL_DEV Dual lit<Dual>(float c) { return {c, 0.f}; }
template <>
TDL_DEV Exact lit<Exact>(float c) { return {c}; }

// M = transformation_from_parameters(v, t, invert), row-major 4x4
template <class T>
TDL_DEV void pose_matrix(const T v[3], const T t[3], bool invert, T M[16]) {
    const T angle = dsqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);          // torch.norm(vec, 2, 2, True)
    const T den = angle + lit<T>(1e-7f);
    const T x = v[0] / den, y = v[1] / den, z = v[2] / den;                  // axis
    const T ca = dcos(angle), sa = dsin(angle);
    const T C = lit<T>(1.f) - ca;
    const T xs = x * sa, ys = y * sa, zs = z * sa;
    const T xC = x * C, yC = y * C, zC = z * C;
    const T xyC = x * yC, yzC = y * zC, zxC = z * xC;
    T R[9];
    R[0] = x * xC + ca;  R[1] = xyC - zs;     R[2] = zxC + ys;
    R[3] = xyC + zs;     R[4] = y * yC + ca;  R[5] = yzC - xs;
    R[6] = zxC - ys;     R[7] = yzC + xs;     R[8] = z * zC + ca;
    if (!invert) {                                                            // T @ R = [R | t]
#pragma unroll
        for (int i = 0; i < 3; ++i) {
#pragma unroll
            for (int j = 0; j < 3; ++j) M[i * 4 + j] = R[i * 3 + j];
            M[i * 4 + 3] = t[i];
        }
    } else {                                                                  // R^T @ [I | -t] = [R^T | -R^T t]
        const T n0 = dneg(t[0]), n1 = dneg(t[1]), n2 = dneg(t[2]);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
#pragma unroll
            for (int j = 0; j < 3; ++j) M[i * 4 + j] = R[j * 3 + i];
            M[i * 4 + 3] = (R[0 * 3 + i] * n0 + R[1 * 3 + i] * n1) + R[2 * 3 + i] * n2;
        }
    }
    M[12] = lit<T>(0.f);
    M[13] = lit<T>(0.f);
    M[14] = lit<T>(0.f);
    M[15] = lit<T>(1.f);
}

__global__ void pose_fwd_kernel(const float* __restrict__ aa, const float* __restrict__ tr, int B, int invert,
                                float* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    Exact v[3], t[3], M[16];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        v[k] = {aa[b * 3 + k]};
        t[k] = {tr[b * 3 + k]};
    }
    pose_matrix<Exact>(v, t, invert != 0, M);
#pragma unroll
    for (int k = 0; k < 16; ++k) out[b * 16 + k] = M[k].v;
}

// one thread per (image, parameter): parameter 0..2 = axis-angle, 3..5 = translation
__global__ void pose_bwd_kernel(const float* __restrict__ aa, const float* __restrict__ tr, const float* __restrict__ dT,
                                int B, int invert, float* __restrict__ d_aa, float* __restrict__ d_tr) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * 6) return;
    const int b = idx / 6, k = idx - b * 6;
    Dual v[3], t[3], M[16];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        v[j] = {aa[b * 3 + j], k == j ? 1.f : 0.f};
        t[j] = {tr[b * 3 + j], k == 3 + j ? 1.f : 0.f};
    }
    pose_matrix<Dual>(v, t, invert != 0, M);
    float g = 0.f;
#pragma unroll
    for (int j = 0; j < 12; ++j) g += dT[b * 16 + j] * M[j].d;                // the last row is constant
    if (k < 3)
        d_aa[b * 3 + k] = g;
    else
        d_tr[b * 3 + k - 3] = g;
}

}  // namespace

cudaError_t launch_pose_fwd(const float* aa, const float* tr, int B, int invert, float* T, cudaStream_t st) {
    pose_fwd_kernel<<<(B + 63) / 64, 64, 0, st>>>(aa, tr, B, invert, T);
    return cudaGetLastError();
}

cudaError_t launch_pose_bwd(const float* aa, const float* tr, const float* dT, int B, int invert, float* d_aa, float* d_tr,
                            cudaStream_t st) {
    pose_bwd_kernel<<<(B * 6 + 63) / 64, 64, 0, st>>>(aa, tr, dT, B, invert, d_aa, d_tr);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// Projection prologue (include/tdl.h tdl_proj_args): one thread per output element.
struct ProjDev {
    int B, S;
    const float* K;
    const float* inv_K;
    const float* T[TDL_MAX_SRC];
    float* P_full;
    float* P_half;
    float* invK3;
    float* invKh3;
    const float* dP_full;
    const float* dP_half;
    float* dT[TDL_MAX_SRC];
};

__global__ void proj_fwd_kernel(const ProjDev p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int nP = p.B * p.S * 12;
    if (i < nP) {
        const int c = i & 3, r = (i >> 2) % 3, f = (i / 12) % p.S, b = i / (12 * p.S);
        const float* K = p.K + (size_t)b * 16 + r * 4;
        const float* T = p.T[0];
#pragma unroll
        for (int q = 1; q < TDL_MAX_SRC; ++q)
            if (q == f) T = p.T[q];
        T += (size_t)b * 16 + c;
        float acc = __fmul_rn(K[0], T[0]);                    // k = 0..3 in order, one rounding per term like a fp32 matmul
        acc = fmaf(K[1], T[4], acc);
        acc = fmaf(K[2], T[8], acc);
        acc = fmaf(K[3], T[12], acc);
        p.P_full[i] = acc;
        p.P_half[i] = r < 2 ? 0.5f * acc : acc;               // (K/2) @ T == (K @ T)/2 exactly: a power-of-two scale
    } else if (i < nP + p.B * 9) {
        const int j = i - nP, c = j % 3, r = (j / 3) % 3, b = j / 9;
        const float v = p.inv_K[(size_t)b * 16 + r * 4 + c];
        p.invK3[j] = v;
        p.invKh3[j] = c < 2 ? 2.f * v : v;                    // pinv(diag(1/2,1/2,1,1) K) = pinv(K) diag(2,2,1,1)
    }
}

__global__ void proj_bwd_kernel(const ProjDev p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.B * p.S * 16) return;
    const int c = i & 3, k = (i >> 2) & 3, f = (i / 16) % p.S, b = i / (16 * p.S);
    float* dT = p.dT[0];
#pragma unroll
    for (int q = 1; q < TDL_MAX_SRC; ++q)
        if (q == f) dT = p.dT[q];
    if (!dT) return;
    const float* K = p.K + (size_t)b * 16;
    const size_t o = ((size_t)b * p.S + f) * 12 + c;
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float g = p.dP_full ? p.dP_full[o + r * 4] : 0.f;
        if (p.dP_half) g += (r < 2 ? 0.5f : 1.f) * p.dP_half[o + r * 4];
        acc = fmaf(K[r * 4 + k], g, acc);
    }
    dT[(size_t)b * 16 + k * 4 + c] = acc;
}

cudaError_t launch_proj_fwd(const tdl_proj_args& a, cudaStream_t st) {
    ProjDev p{a.B, a.S, a.K, a.inv_K, {a.T[0], a.T[1], a.T[2], a.T[3]}, a.P_full, a.P_half, a.invK3, a.invKh3, nullptr, nullptr,
              {nullptr, nullptr, nullptr, nullptr}};
    const int n = a.B * a.S * 12 + a.B * 9;
    proj_fwd_kernel<<<(n + 127) / 128, 128, 0, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_proj_bwd(const tdl_proj_args& a, cudaStream_t st) {
    ProjDev p{a.B, a.S, a.K, a.inv_K, {nullptr, nullptr, nullptr, nullptr}, nullptr, nullptr, nullptr, nullptr, a.dP_full, a.dP_half,
              {a.dT[0], a.dT[1], a.dT[2], a.dT[3]}};
    const int n = a.B * a.S * 16;
    proj_bwd_kernel<<<(n + 127) / 128, 128, 0, st>>>(p);
    return cudaGetLastError();
}

}  // namespace tdl
