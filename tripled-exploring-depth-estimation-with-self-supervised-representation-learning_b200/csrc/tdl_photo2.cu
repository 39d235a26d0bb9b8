// Photometric scoring kernel, round 2: register micro-tiles (tdl_ssim.cuh) over TMA-staged tiles (sm_100a).
//
//   photo_score2_kernel   one CTA = one 32x32 full-resolution tile of one image, ALL scales; 128 threads, each owning
//                         a 2 x 4 pixel patch.  Target + source tiles (identity / auto-mask terms) are staged once with
//                         3-D TMA box copies, then per scale the S warped tiles written by photo_warp_kernel; the
//                         target-only window statistics are computed once per tile and shared by all 2 + 4*S image
//                         comparisons.  Per-image sums leave through warp shuffles + one fp64 atomic per CTA.
//
// Staged box: global columns tx0 .. tx0+35 (the tile, its right halo column and 3 spare columns), rows ty0-1 .. ty0+32.
// The box origin must sit on a 16-byte boundary, so the LEFT halo column tx0-1 cannot be part of it without widening
// the box to 40 columns: its 34 values per plane are fetched with plain loads and parked in the spare cell that
// precedes each staged row (row r, column -1 == row r-1, column 35 in the dense box), after the TMA copy has landed --
// every patch then reads its left neighbour at row[-1].  44 KB of tiles + 24 KB of statistics + 4 KB of stashed noise
// per CTA: three CTAs (12 warps) per SM.
//
// Reference behaviour: mono/model/mono_fm/net.py:63-67,90-106 and mono/model/mono_fm/layers.py:97-107 (include/tdl.h).
#include "tdl_common.cuh"
#include "tdl_internal.h"
#include "tdl_ssim.cuh"
#include "tdl_tma.cuh"

#include <cuda_fp16.h>

namespace tdl {

namespace s2 {
constexpr int TW = 32, TH = 32;                  // tile
constexpr int NT = (TW / kPC) * (TH / kPR);      // 128 threads
constexpr int BW = TW + 4, BH = TH + 2;          // staged box (columns tx0 .. tx0+35, rows ty0-1 .. ty0+32)
constexpr int BPLANE = BW * BH;
constexpr int GROUP = (3 * BPLANE + 31) / 32 * 32;       // 3 planes, 128-byte multiple (TMA destination)
constexpr int PAD = 32;                          // floats in front of group 0: cell "row 0, column -1" of its first plane
constexpr uint32_t kGroupBytes = 3 * BPLANE * sizeof(float);
static_assert(GROUP > 3 * BPLANE, "the cell in front of a group's first plane lives in the previous group's padding");
}  // namespace s2

// N(0,1) draws for the identity channels (see automask_noise in tdl_photo.cu); one Philox4x32-10 call yields four
// normals: for S <= 2 they serve TWO scales (counter word z = scale pair, the second pair is stashed in shared memory),
// for S > 2 one scale.
template <int S>
TDL_DEV void philox_normals(const PhotoDev& p, int zword, int b, size_t pix, size_t HW, float out[4]) {
    const unsigned long long idx = (unsigned long long)b * HW + pix;
    const uint4 r = philox4x32(make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)zword, S <= 2 ? 1u : 0u),
                               make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32)));
    const float2 a = box_muller(r.x, r.y), c = box_muller(r.z, r.w);
    out[0] = a.x;
    out[1] = a.y;
    out[2] = c.x;
    out[3] = c.y;
}

template <int S>
__global__ void __launch_bounds__(s2::NT, S <= 2 ? 3 : 2) photo_score2_kernel(const PhotoDev p, const __grid_constant__ PhotoMaps maps,
                                                                 const __grid_constant__ PhotoMaps wmaps) {
    using namespace s2;
    extern __shared__ __align__(128) float smem2[];
    float* s_tgt = smem2 + PAD;                             // group 0: target
    float* s_img = s_tgt + GROUP;                           // groups 1..S: sources (identity terms), then warped per scale
    float4* s_stat = reinterpret_cast<float4*>(s_img + S * GROUP);
                                                            // [3 ch][kPR][NT] float4 = (Sy, Qy) of the two patch columns
    __half* s_nz = reinterpret_cast<__half*>(s_stat + 3 * kPR * NT);      // [2][kPP][NT] stashed normals (S <= 2)
    __shared__ uint64_t s_bar;
    __shared__ float s_part[2 * TDL_MAX_SCALES][NT / 32];
    __shared__ float s_pyr[NT / 32][2][3];                  // pyramid: per warp (8 tile rows) and column half, per channel

    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int k16 = tid & 15, rg = tid >> 4;                // patch column pair / row group
    const int b = blockIdx.z, tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
    const int H = p.H, W = p.W;
    const size_t HW = (size_t)H * W;

    if (tid == 0) {
        mbar_init(&s_bar, 1);
        mbar_arrive_expect_tx(&s_bar, (1 + S) * kGroupBytes);
        tma_load_3d(s_tgt, &maps.tgt, &s_bar, tx0, ty0 - 1, b * 3);
#pragma unroll
        for (int f = 0; f < S; ++f) tma_load_3d(s_img + f * GROUP, &maps.img[f], &s_bar, tx0, ty0 - 1, b * 3);
    }
    // left-halo column (global column reflect(tx0 - 1), rows reflect(ty0 - 1 + r)) of `ngroups` staged groups: the values
    // are fetched while the TMA copies fly and parked at (row r, column -1) once the copies have landed
    constexpr int kLeftPerGroup = 3 * BH;                   // 102 cells per group
    constexpr int kLeftIt = ((1 + S) * kLeftPerGroup + NT - 1) / NT;
    const int lx = reflect1(tx0 - 1, W);
    float lreg[kLeftIt];
    auto left_fetch = [&](int ngroups, const float* const* bases) {
#pragma unroll
        for (int it = 0; it < kLeftIt; ++it) {
            const int e = tid + it * NT;
            lreg[it] = 0.f;
            if (e < ngroups * kLeftPerGroup) {
                const int g = e / kLeftPerGroup, rem = e - g * kLeftPerGroup, ch = rem / BH, r = rem - ch * BH;
                const float* base = bases[0];
#pragma unroll
                for (int q = 1; q < 1 + S; ++q)
                    if (q == g) base = bases[q];
                lreg[it] = __ldg(base + ((size_t)b * 3 + ch) * HW + (size_t)reflect1(ty0 - 1 + r, H) * W + lx);
            }
        }
    };
    auto left_park = [&](float* groups, int ngroups) {
#pragma unroll
        for (int it = 0; it < kLeftIt; ++it) {
            const int e = tid + it * NT;
            if (e < ngroups * kLeftPerGroup) {
                const int g = e / kLeftPerGroup, rem = e - g * kLeftPerGroup, ch = rem / BH, r = rem - ch * BH;
                groups[g * GROUP + ch * BPLANE + r * BW - 1] = lreg[it];
            }
        }
    };
    {
        const float* bases[1 + S];
        bases[0] = p.target;
#pragma unroll
        for (int f = 0; f < S; ++f) bases[1 + f] = p.src[f];
        left_fetch(1 + S, bases);
    }
    __syncthreads();                                        // barrier initialisation visible to the waiting threads
    uint32_t parity = 0;
    mbar_wait(&s_bar, parity);
    parity ^= 1;

    // nn.ReflectionPad2d(1): the zero-filled cells one pixel outside the image take their mirror value (columns 0..TW
    // only: the spare columns behind them hold the parked left halo)
    auto reflect_fix2 = [&](float* groups, int ngroups) {
        const bool br = tx0 + TW >= W, bt_ = ty0 == 0, bb = ty0 + TH >= H;
        if (br) {
            for (int e = tid; e < ngroups * 3 * BH; e += NT) {
                const int g = e / (3 * BH), rem = e - g * 3 * BH, ch = rem / BH, r = rem - ch * BH;
                float* row = groups + g * GROUP + ch * BPLANE + r * BW;
                row[W - tx0] = row[W - tx0 - 2];
            }
        }
        if (bt_ || bb) {
            if (br) __syncthreads();
            for (int e = tid; e < ngroups * 3 * (TW + 1); e += NT) {
                const int g = e / (3 * (TW + 1)), rem = e - g * 3 * (TW + 1), ch = rem / (TW + 1), c = rem - ch * (TW + 1);
                float* col = groups + g * GROUP + ch * BPLANE + c;
                if (bt_) col[0] = col[2 * BW];
                if (bb) col[(H - ty0 + 1) * BW] = col[(H - ty0 - 1) * BW];
            }
        }
    };
    left_park(s_tgt, 1 + S);
    reflect_fix2(s_tgt, 1 + S);
    __syncthreads();

    // ---- the thread's patch: columns 2*k16, 2*k16+1, rows 4*rg .. 4*rg+3 of the tile
    const int pc0 = kPC * k16, pr0 = kPR * rg;
    const int poff = pr0 * BW + pc0;                        // (patch row 0 - 1, patch column 0) inside a staged plane

    // ---- area-downsampled target pyramid (F.interpolate(mode='area'), net.py:259) and disparity sums, as a butterfly over
    //      the patches: a thread adds its own 2x2 blocks, lanes exchange partial sums with shuffles (k16 is lane & 15, the
    //      row group the lane's bit 4) and only the 16- and 32-pixel cells cross warps through 24 floats of shared
    //      memory -- ~80 instructions per thread where the per-cell loops (runtime divisors, serial fac*fac sums, a handful
    //      of busy threads on the coarse scales) took ~1900: 12 % of this kernel's instructions and of its stall samples
    float dsum[TDL_MAX_SCALES], lsum[TDL_MAX_SCALES];
#pragma unroll
    for (int s = 0; s < TDL_MAX_SCALES; ++s) dsum[s] = lsum[s] = 0.f;
    {
        float lv2[3][2], lv4[3], lv8[3];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const float* t = s_tgt + ch * BPLANE + poff + BW;          // patch row 0
            float2 v[kPR];
#pragma unroll
            for (int r = 0; r < kPR; ++r) v[r] = *reinterpret_cast<const float2*>(t + r * BW);
            lv2[ch][0] = (v[0].x + v[0].y) + (v[1].x + v[1].y);
            lv2[ch][1] = (v[2].x + v[2].y) + (v[3].x + v[3].y);
            float a = lv2[ch][0] + lv2[ch][1];
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            lv4[ch] = a;                                                // 4 x 4 pixels
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            a += __shfl_xor_sync(0xffffffffu, a, 16);
            lv8[ch] = a;                                                // 8 x 8
            a += __shfl_xor_sync(0xffffffffu, a, 4);                    // 16 columns x the warp's 8 rows
            if ((lane & 23) == 0) s_pyr[wrp][lane >> 3][ch] = a;        // lanes 0 and 8: column halves 0 and 1
        }
#pragma unroll
        for (int s = 0; s < TDL_MAX_SCALES; ++s) {
            if (s < p.nscales) {
                const int fac = p.fac[s], h = p.dh[s], w = p.dw[s];
                if (fac > 8) continue;                                  // the 16- / 32-pixel cells follow the next barrier
                const bool owner = fac <= 2 || (fac == 4 ? (k16 & 1) == 0 : ((k16 & 3) == 0 && (rg & 1) == 0));
                const int ncell = fac == 1 ? kPP : (fac == 2 ? 2 : 1);
                const float inv = 1.f / (float)(fac * fac);
#pragma unroll
                for (int c = 0; c < kPP; ++c) {
                    if (c < ncell && owner) {
                        // cell c of the patch at this scale: fac 1: pixel c; fac 2: row pair c; fac 4 / 8: the one cell
                        const int cj = (ty0 + pr0 + (fac == 1 ? c / kPC : (fac == 2 ? 2 * c : 0))) / fac;
                        const int ci = (tx0 + pc0 + (fac == 1 ? c % kPC : 0)) / fac;
                        if (cj < h && ci < w) {
#pragma unroll
                            for (int ch = 0; ch < 3; ++ch) {
                                const float v = fac == 1 ? s_tgt[ch * BPLANE + poff + BW + (c / kPC) * BW + c % kPC]
                                                         : (fac == 2 ? lv2[ch][c & 1] : (fac == 4 ? lv4[ch] : lv8[ch]));
                                p.J[s][(((size_t)b * 3 + ch) * h + cj) * w + ci] = v * inv;
                            }
                            dsum[s] += __ldg(p.disp[s] + ((size_t)b * h + cj) * w + ci);
                        }
                    }
                }
            }
        }
    }

    float4* my_stat = s_stat + tid;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        float2 st[kPP];
        patch_target_stats<BW>(s_tgt + ch * BPLANE + poff, st);
#pragma unroll
        for (int i = 0; i < kPR; ++i)
            my_stat[(ch * kPR + i) * NT] = make_float4(st[i * kPC].x, st[i * kPC].y, st[i * kPC + 1].x, st[i * kPC + 1].y);
    }

    // rho of the S prediction groups starting at `groups` (frame f at groups + f*GROUP)
    auto score_frames = [&](const float* groups, float (&rho)[S][kPP]) {
        constexpr int NFA = S >= 2 ? 2 : 1;                // frames processed together (they share the target rows)
#pragma unroll
        for (int f0 = 0; f0 < S; f0 += NFA) {
            const bool pair = (S - f0) >= NFA;
            float sa[NFA][kPP], la[NFA][kPP];
#pragma unroll
            for (int f = 0; f < NFA; ++f)
#pragma unroll
                for (int i = 0; i < kPP; ++i) sa[f][i] = la[f][i] = 0.f;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                float2 st[kPP];
#pragma unroll
                for (int i = 0; i < kPR; ++i) {
                    const float4 q = my_stat[(ch * kPR + i) * NT];
                    st[i * kPC] = make_float2(q.x, q.y);
                    st[i * kPC + 1] = make_float2(q.z, q.w);
                }
                const float* ys = s_tgt + ch * BPLANE + poff;
                if (pair) {
                    const float* xs[NFA];
#pragma unroll
                    for (int f = 0; f < NFA; ++f) xs[f] = groups + (f0 + f) * GROUP + ch * BPLANE + poff;
                    patch_ssim_l1<NFA, BW>(xs, ys, st, sa, la);
                } else {
                    const float* xs[1] = {groups + f0 * GROUP + ch * BPLANE + poff};
                    float sa1[1][kPP], la1[1][kPP];
#pragma unroll
                    for (int i = 0; i < kPP; ++i) {
                        sa1[0][i] = sa[0][i];
                        la1[0][i] = la[0][i];
                    }
                    patch_ssim_l1<1, BW>(xs, ys, st, sa1, la1);
#pragma unroll
                    for (int i = 0; i < kPP; ++i) {
                        sa[0][i] = sa1[0][i];
                        la[0][i] = la1[0][i];
                    }
                }
            }
#pragma unroll
            for (int f = 0; f < NFA; ++f)
                if (f == 0 || pair)
#pragma unroll
                    for (int i = 0; i < kPP; ++i) rho[f0 + f][i] = rho_from_sums(sa[f][i], la[f][i]);
        }
    };

    float rho_id[S][kPP];
    if (p.automask) score_frames(s_img, rho_id);
    __syncthreads();                                        // every thread is done with the source tiles

    // ---- the pyramid's 16- and 32-pixel cells from the per-warp partial sums (written before the barrier above)
#pragma unroll
    for (int s = 0; s < TDL_MAX_SCALES; ++s) {
        if (s < p.nscales && p.fac[s] > 8) {
            const int fac = p.fac[s], h = p.dh[s], w = p.dw[s];
            const int cells = TW / fac;                              // 2 or 1 per tile side
            if (tid < cells * cells * 3) {
                const int ch = tid % 3, cell = tid / 3, cy = cell / cells, cx = cell - cy * cells;
                float v;
                if (fac == 16) v = s_pyr[2 * cy][cx][ch] + s_pyr[2 * cy + 1][cx][ch];
                else v = ((s_pyr[0][0][ch] + s_pyr[0][1][ch]) + (s_pyr[1][0][ch] + s_pyr[1][1][ch])) +
                         ((s_pyr[2][0][ch] + s_pyr[2][1][ch]) + (s_pyr[3][0][ch] + s_pyr[3][1][ch]));
                const int cj = ty0 / fac + cy, ci = tx0 / fac + cx;
                if (cj < h && ci < w) {
                    p.J[s][(((size_t)b * 3 + ch) * h + cj) * w + ci] = v * (1.f / (float)(fac * fac));
                    if (ch == 0) dsum[s] += __ldg(p.disp[s] + ((size_t)b * h + cj) * w + ci);
                }
            }
        }
    }

    const int gx0 = tx0 + pc0, gy0 = ty0 + pr0;
    for (int s = 0; s < p.nscales; ++s) {
        // ---- stage the S warped tiles of this scale (written by photo_warp_kernel)
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive_expect_tx(&s_bar, S * kGroupBytes);
#pragma unroll
            for (int f = 0; f < S; ++f) tma_load_3d(s_img + f * GROUP, &wmaps.img[s * S + f], &s_bar, tx0, ty0 - 1, b * 3);
        }
        {
            const float* bases[1 + S];
#pragma unroll
            for (int f = 0; f < S; ++f) bases[f] = p.warped[s][f];
            bases[S] = nullptr;
            left_fetch(S, bases);
        }
        // ---- identity channels + tie-break noise while the copies fly
        float best[kPP];
        int arg[kPP];
#pragma unroll
        for (int i = 0; i < kPP; ++i) {
            best[i] = 0.f;
            arg[i] = -1;
        }
        if (p.automask) {
#pragma unroll
            for (int i = 0; i < kPP; ++i) {
                const int gy = gy0 + i / kPC, gx = gx0 + i % kPC;
                float nz[S];
#pragma unroll
                for (int f = 0; f < S; ++f) nz[f] = 0.f;
                if (gx < W && gy < H) {
                    const size_t pix = (size_t)gy * W + gx;
                    if (p.noise[s][0]) {
#pragma unroll
                        for (int f = 0; f < S; ++f) nz[f] = __ldg(p.noise[s][f] + (size_t)b * HW + pix);
                    } else if (S <= 2) {
                        if ((s & 1) == 0) {
                            float n4[4];
                            philox_normals<S>(p, s >> 1, b, pix, HW, n4);
#pragma unroll
                            for (int f = 0; f < S; ++f) {
                                nz[f] = n4[f];
                                s_nz[(f * kPP + i) * NT + tid] = __float2half_rn(n4[2 + f]);
                            }
                        } else {
#pragma unroll
                            for (int f = 0; f < S; ++f) nz[f] = __half2float(s_nz[(f * kPP + i) * NT + tid]);
                        }
                    } else {
                        float n4[4];
                        philox_normals<S>(p, s, b, pix, HW, n4);
#pragma unroll
                        for (int f = 0; f < S; ++f) nz[f] = n4[f];
                    }
                }
#pragma unroll
                for (int f = 0; f < S; ++f) {
                    const float v = fmaf(nz[f], 1e-5f, rho_id[f][i]);                   // net.py:94
                    if (f == 0 || v < best[i]) {
                        best[i] = v;
                        arg[i] = f;
                    }
                }
            }
        }
        mbar_wait(&s_bar, parity);
        parity ^= 1;
        left_park(s_img, S);
        reflect_fix2(s_img, S);
        __syncthreads();

        // ---- reprojection errors of the warped frames, minimum over all channels
        const int chan0 = p.automask ? S : 0;
        {
            float rho[S][kPP];
            score_frames(s_img, rho);
#pragma unroll
            for (int f = 0; f < S; ++f)
#pragma unroll
                for (int i = 0; i < kPP; ++i)
                    if (arg[i] < 0 || rho[f][i] < best[i]) {
                        best[i] = rho[f][i];
                        arg[i] = chan0 + f;
                    }
        }
        float ls = 0.f;
#pragma unroll
        for (int r = 0; r < kPR; ++r) {
            const int gy = gy0 + r;
            if (gy < H && gx0 < W) {                       // W % 4 == 0 on this path: both patch columns are inside together
                const size_t pix = (size_t)gy * W + gx0;
                ls += best[r * kPC] + best[r * kPC + 1];
                *reinterpret_cast<uchar2*>(p.argmin + ((size_t)s * p.B + b) * HW + pix) =
                    make_uchar2((unsigned char)arg[r * kPC], (unsigned char)arg[r * kPC + 1]);
                if (p.min_index[s])
                    *reinterpret_cast<longlong2*>(p.min_index[s] + (size_t)b * HW + pix) =
                        make_longlong2(arg[r * kPC], arg[r * kPC + 1]);
            }
        }
#pragma unroll
        for (int q = 0; q < TDL_MAX_SCALES; ++q)
            if (q == s) lsum[q] = ls;
        // ---- work list of the backward (photo_bwd_list_kernel): every window whose arg-min is a warped frame is appended
        //      to the list of its (scale, image) as pixel | frame << 28; one atomic per warp reserves the slots.  The count
        //      keeps running past the capacity (the backward then knows the list is incomplete and runs the tile kernel).
        // (a list that already overflowed is closed: its count only has to stay above the capacity, so the warps of a
        //  densely selected image skip the scan, the atomic and the stores -- the count is read past L1, one lane decides)
        bool list_open = p.list_max >= 0;
        if (list_open) list_open = __shfl_sync(0xffffffffu, __ldcg(p.lcnt + s * p.B + b) <= kListCap ? 1 : 0, 0) != 0;
        if (list_open) {
            int cnt = 0;
#pragma unroll
            for (int i = 0; i < kPP; ++i) cnt += (arg[i] >= chan0 && gy0 + i / kPC < H && gx0 < W) ? 1 : 0;
            int incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            int base = 0;
            if (lane == 31 && incl > 0) base = atomicAdd(p.lcnt + s * p.B + b, incl);
            base = __shfl_sync(0xffffffffu, base, 31);
            if (cnt > 0 && base < kListCap) {
                int off = base + incl - cnt;
                uint32_t* wl = p.wlist + ((size_t)s * p.B + b) * kListCap;
#pragma unroll
                for (int i = 0; i < kPP; ++i) {
                    const int gy = gy0 + i / kPC, gx = gx0 + i % kPC;
                    if (arg[i] >= chan0 && gy < H && gx0 < W) {
                        if (off < kListCap) wl[off] = (uint32_t)(gy * W + gx) | ((uint32_t)(arg[i] - chan0) << 28);
                        ++off;
                    }
                }
            }
        }
        __syncthreads();           // every thread is done with this scale's tiles before the next TMA overwrites them
    }
    if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) p.lcnt[p.nscales * p.B] = kListMagic;
    // ---- one CTA reduction for the 2 * nscales partial sums
#pragma unroll
    for (int q = 0; q < TDL_MAX_SCALES; ++q) {
        const float a = warp_sum(lsum[q]), d = warp_sum(dsum[q]);
        if (lane == 0) {
            s_part[q][wrp] = a;
            s_part[TDL_MAX_SCALES + q][wrp] = d;
        }
    }
    __syncthreads();
    if (tid < 2 * TDL_MAX_SCALES) {
        const int q = tid % TDL_MAX_SCALES;
        if (q < p.nscales) {
            float t = 0.f;
#pragma unroll
            for (int j = 0; j < NT / 32; ++j) t += s_part[tid][j];
            atomicAdd(p.acc + ((size_t)q * p.B + b) * 4 + (tid < TDL_MAX_SCALES ? 0 : 1), (double)t);
        }
    }
}

template <int S>
static size_t score2_smem() {
    using namespace s2;
    const size_t fl = (size_t)PAD + (size_t)(1 + S) * GROUP;
    return fl * sizeof(float) + (size_t)3 * kPR * NT * sizeof(float4) + (S <= 2 ? (size_t)2 * kPP * NT * sizeof(__half) : 0);
}

template <int S>
static bool encode_score2_maps(const PhotoDev& p, PhotoMaps* maps, PhotoMaps* wmaps) {
    using namespace s2;
    bool ok = p.use_tma && p.split_fwd && (p.W % 4) == 0 && encode_image_map(&maps->tgt, p.target, p.B * 3, p.H, p.W, BW, BH, 3);
    for (int f = 0; ok && f < S; ++f) ok = encode_image_map(&maps->img[f], p.src[f], p.B * 3, p.H, p.W, BW, BH, 3);
    for (int s = 0; ok && s < p.nscales; ++s)
        for (int f = 0; ok && f < S; ++f)
            ok = p.warped[s][f] && encode_image_map(&wmaps->img[s * S + f], p.warped[s][f], p.B * 3, p.H, p.W, BW, BH, 3);
    return ok;
}

template <int S>
static cudaError_t launch_score2_t(const PhotoDev& p, cudaStream_t st) {
    using namespace s2;
    PhotoMaps maps, wmaps;
    if (!encode_score2_maps<S>(p, &maps, &wmaps)) return cudaErrorInvalidValue;
    const size_t smem = score2_smem<S>();
    static SmemOptIn opt_in;
    if (cudaError_t e = opt_in(photo_score2_kernel<S>, smem)) return e;
    dim3 grid((p.W + TW - 1) / TW, (p.H + TH - 1) / TH, p.B);
    photo_score2_kernel<S><<<grid, NT, smem, st>>>(p, maps, wmaps);
    return cudaGetLastError();
}

cudaError_t launch_photo_score2(const PhotoDev& p, cudaStream_t st) {
    switch (p.S) {
        case 1: return launch_score2_t<1>(p, st);
        case 2: return launch_score2_t<2>(p, st);
        case 3: return launch_score2_t<3>(p, st);
        case 4: return launch_score2_t<4>(p, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace tdl
