// Micro-benchmark: gathering 256-byte feature rows (the access pattern of feat_fwd_nhwc_kernel: per pixel one target row and,
// per source frame, two 512-byte row pairs at a data-dependent position) into shared memory with
//   (a) cp.async.ca 16 bytes per lane (LDGSTS), the kernel's present 3-stage ring, and
//   (b) cp.async.bulk (TMA bulk copy, UBLKCP) issued by one lane per copy, completion on an mbarrier per ring stage.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rowgather rowgather.cu ; run: ./rowgather
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int C = 64, H = 96, W = 320, B = 8, S = 2;
constexpr int NT = 128, PIX = 32, ROWS = 1 + 4 * S;        // rows per pixel and step (two pixels per step)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(n));
}
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra D;\n bra W;\n D:\n}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// tap position of pixel (y, x) for frame f: a smooth displacement field, like a coherent flow
#ifndef COHERENT
#define COHERENT 0
#endif
__device__ __forceinline__ int tap_of(int y, int x, int f) {
    int ty = y + ((x * 7 + y * 3 + f * 11) % 5) - 2, tx = x + ((x * 5 + y * 13 + f * 7) % 9) - 4;
    if (COHERENT) {                                        // a smooth flow: neighbouring pixels sample neighbouring positions
        ty = y + 1 + f + ((y >> 3) & 1);
        tx = x + 2 - 3 * f + ((x >> 4) & 1);
    }
    ty = min(max(ty, 0), H - 2);
    tx = min(max(tx, 0), W - 2);
    return ty * W + tx;
}

template <int MODE, int STAGES, bool MERGE = false, bool PAR = false>
__global__ void __launch_bounds__(NT) gather(const float* __restrict__ tgt, const float* __restrict__ s0, const float* __restrict__ s1,
                                             float* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char raw[];
    __shared__ uint64_t bars[NT / 32][STAGES];
    const int tid = threadIdx.x, lane = tid & 31, wq = tid >> 5, half = lane >> 4, l16 = lane & 15;
    const int b = blockIdx.y, pix0 = (blockIdx.x * (NT / 32) + wq) * PIX;
    const size_t img = (size_t)b * H * W * C;
    const float* src[2] = {s0 + img, s1 + img};
    // ring[warp][stage][pixel of the step: 2][row][64 floats]
    float* ring = reinterpret_cast<float*>(raw) + (size_t)wq * STAGES * 2 * ROWS * C;
    if (MODE == 1) {
        if (lane == 0)
            for (int k = 0; k < STAGES; ++k) mbar_init(&bars[wq][k], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncwarp();
    }
    auto issue = [&](int q, int stage) {
        float* st = ring + (size_t)stage * 2 * ROWS * C;
        if (MODE == 0) {
            const int pix = pix0 + q + half, y = pix / W, x = pix - y * W;
            float* d = st + (size_t)half * ROWS * C + 4 * l16;
            auto cp = [&](float* dst, const float* s) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(s) : "memory");
            };
            cp(d, tgt + img + (size_t)pix * C + 4 * l16);
#pragma unroll
            for (int f = 0; f < S; ++f) {
                const int o = tap_of(y, x, f);
                const float* sb = src[f] + (size_t)o * C + 4 * l16;
                cp(d + (1 + 4 * f) * C, sb);
                cp(d + (2 + 4 * f) * C, sb + C);
                cp(d + (3 + 4 * f) * C, sb + (size_t)W * C);
                cp(d + (4 + 4 * f) * C, sb + (size_t)W * C + C);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        } else {
            __syncwarp();                                   // every lane has read the stage that is refilled
            if (PAR) {
                // lanes 0 .. 2 * (1 + 2S) - 1 request one copy each: lane = pixel of the step x {target, (frame, north | south)}
                if (lane == 0) mbar_expect(&bars[wq][stage], 2 * ROWS * C * 4);
                __syncwarp();
                constexpr int PER = 1 + 2 * S;
                if (lane < 2 * PER) {
                    const int hh = lane / PER, k = lane - hh * PER;
                    const int pix = pix0 + q + hh, y = pix / W, x = pix - y * W;
                    float* d = st + (size_t)hh * ROWS * C;
                    if (k == 0) bulk(d, tgt + img + (size_t)pix * C, C * 4, &bars[wq][stage]);
                    else {
                        const int f = (k - 1) >> 1, south = (k - 1) & 1;
                        const int o = tap_of(y, x, f);
                        bulk(d + (1 + 4 * f + 2 * south) * C, src[f] + (size_t)o * C + (south ? (size_t)W * C : 0), 2 * C * 4, &bars[wq][stage]);
                    }
                }
            } else if (lane == 0) {
                uint32_t bytes = 2 * ROWS * C * 4;
                if (MERGE) {
                    const int pixa = pix0 + q, ya = pixa / W, xa = pixa - ya * W;
#pragma unroll
                    for (int f = 0; f < S; ++f)
                        if (tap_of(ya, xa + 1, f) == tap_of(ya, xa, f) + 1) bytes -= 2 * C * 4;   // 2 x 3 rows instead of 4 x 2
                }
                mbar_expect(&bars[wq][stage], bytes);
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int pix = pix0 + q + hh, y = pix / W, x = pix - y * W;
                    float* d = st + (size_t)hh * ROWS * C;
                    bulk(d, tgt + img + (size_t)pix * C, C * 4, &bars[wq][stage]);
#pragma unroll
                    for (int f = 0; f < S; ++f) {
                        const int o = tap_of(y, x, f);
                        const float* sb = src[f] + (size_t)o * C;
                        if (MERGE) {
                            // the two pixels of a step are neighbours: when their taps are too, one 3-row copy serves both
                            const int oa = tap_of(y, x - hh, f), ob = tap_of(y, x - hh + 1, f);
                            if (ob == oa + 1) {
                                if (hh == 0) {
                                    bulk(d + (1 + 4 * f) * C, sb, 3 * C * 4, &bars[wq][stage]);
                                    bulk(d + (1 + 4 * f) * C + (size_t)ROWS * C, sb + (size_t)W * C, 3 * C * 4, &bars[wq][stage]);
                                }
                                continue;
                            }
                        }
                        bulk(d + (1 + 4 * f) * C, sb, 2 * C * 4, &bars[wq][stage]);                       // nw, ne
                        bulk(d + (3 + 4 * f) * C, sb + (size_t)W * C, 2 * C * 4, &bars[wq][stage]);       // sw, se
                    }
                }
            }
        }
    };
    float acc = 0.f;
    int iq = 0;
    for (int k = 0; k < STAGES - 1; ++k, iq += 2) issue(iq, k);
    uint32_t par[STAGES] = {};
    int stage = 0;
    for (int q = 0; q < PIX; q += 2) {
        if (iq < PIX) issue(iq, (stage + STAGES - 1) % STAGES);
        else if (MODE == 0) asm volatile("cp.async.commit_group;" ::: "memory");
        iq += 2;
        if (MODE == 0) asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 1) : "memory");
        else {
#pragma unroll
            for (int k = 0; k < STAGES; ++k)
                if (k == stage) {
                    mbar_wait(&bars[wq][k], par[k]);
                    par[k] ^= 1;
                }
        }
        const float* st = ring + (size_t)stage * 2 * ROWS * C + (size_t)half * ROWS * C + 4 * l16;
        const float4 t = *reinterpret_cast<const float4*>(st);
#pragma unroll
        for (int r = 1; r < ROWS; ++r) {
            const float4 v = *reinterpret_cast<const float4*>(st + r * C);
            acc += (v.x - t.x) + (v.y - t.y) + (v.z - t.z) + (v.w - t.w);
        }
        stage = stage + 1 == STAGES ? 0 : stage + 1;
    }
    out[((size_t)b * gridDim.x + blockIdx.x) * NT + tid] = acc;
}

template <int MODE, int STAGES, bool MERGE = false, bool PAR = false>
float run(const float* t, const float* a, const float* b, float* o) {
    const size_t smem = (size_t)(NT / 32) * STAGES * 2 * ROWS * C * 4;
    cudaFuncSetAttribute(gather<MODE, STAGES, MERGE, PAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(H * W / ((NT / 32) * PIX), B);
    gather<MODE, STAGES, MERGE, PAR><<<grid, NT, smem>>>(t, a, b, o);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < 20; ++i) gather<MODE, STAGES, MERGE, PAR><<<grid, NT, smem>>>(t, a, b, o);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("mode %d stages %d: %s, smem %zu B/CTA, %s\n", MODE, STAGES, MODE ? "cp.async.bulk + mbarrier" : "cp.async.ca 16 B/lane", smem,
           cudaGetErrorString(cudaGetLastError()));
    return ms / 20 * 1e3f;
}

// MODE 2: bulk copies, ONE pixel per step (the whole warp reads it, 8 bytes per lane): half the stage size, twice the CTAs
template <int STAGES>
__global__ void __launch_bounds__(NT) gather1(const float* __restrict__ tgt, const float* __restrict__ s0, const float* __restrict__ s1,
                                              float* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char raw[];
    __shared__ uint64_t bars[NT / 32][STAGES];
    const int tid = threadIdx.x, lane = tid & 31, wq = tid >> 5;
    const int b = blockIdx.y, pix0 = (blockIdx.x * (NT / 32) + wq) * PIX;
    const size_t img = (size_t)b * H * W * C;
    const float* src[2] = {s0 + img, s1 + img};
    float* ring = reinterpret_cast<float*>(raw) + (size_t)wq * STAGES * ROWS * C;
    if (lane == 0)
        for (int k = 0; k < STAGES; ++k) mbar_init(&bars[wq][k], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    auto issue = [&](int q, int stage) {
        float* d = ring + (size_t)stage * ROWS * C;
        __syncwarp();
        if (lane == 0) {
            mbar_expect(&bars[wq][stage], ROWS * C * 4);
            const int pix = pix0 + q, y = pix / W, x = pix - y * W;
            bulk(d, tgt + img + (size_t)pix * C, C * 4, &bars[wq][stage]);
#pragma unroll
            for (int f = 0; f < S; ++f) {
                const int o = tap_of(y, x, f);
                const float* sb = src[f] + (size_t)o * C;
                bulk(d + (1 + 4 * f) * C, sb, 2 * C * 4, &bars[wq][stage]);
                bulk(d + (3 + 4 * f) * C, sb + (size_t)W * C, 2 * C * 4, &bars[wq][stage]);
            }
        }
    };
    float acc = 0.f;
    int iq = 0;
    for (int k = 0; k < STAGES - 1; ++k, ++iq) issue(iq, k);
    uint32_t par[STAGES] = {};
    int stage = 0;
    for (int q = 0; q < PIX; ++q) {
        if (iq < PIX) issue(iq, (stage + STAGES - 1) % STAGES);
        ++iq;
#pragma unroll
        for (int k = 0; k < STAGES; ++k)
            if (k == stage) {
                mbar_wait(&bars[wq][k], par[k]);
                par[k] ^= 1;
            }
        const float* st = ring + (size_t)stage * ROWS * C + 2 * lane;
        const float2 t = *reinterpret_cast<const float2*>(st);
#pragma unroll
        for (int r = 1; r < ROWS; ++r) {
            const float2 v = *reinterpret_cast<const float2*>(st + r * C);
            acc += (v.x - t.x) + (v.y - t.y);
        }
        stage = stage + 1 == STAGES ? 0 : stage + 1;
    }
    out[((size_t)b * gridDim.x + blockIdx.x) * NT + tid] = acc;
}

template <int STAGES>
float run1(const float* t, const float* a, const float* b, float* o) {
    const size_t smem = (size_t)(NT / 32) * STAGES * ROWS * C * 4;
    cudaFuncSetAttribute(gather1<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(H * W / ((NT / 32) * PIX), B);
    gather1<STAGES><<<grid, NT, smem>>>(t, a, b, o);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < 20; ++i) gather1<STAGES><<<grid, NT, smem>>>(t, a, b, o);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / 20 * 1e3f;
}

int main() {
    const size_t n = (size_t)B * H * W * C;
    float *t, *a, *b, *o;
    cudaMalloc(&t, n * 4);
    cudaMalloc(&a, n * 4);
    cudaMalloc(&b, n * 4);
    cudaMalloc(&o, n * 4);
    cudaMemset(t, 0, n * 4);
    cudaMemset(a, 0, n * 4);
    cudaMemset(b, 0, n * 4);
    const double bytes = (double)n * 4 * 3;                 // compulsory: each array read once
    for (int rep = 0; rep < 2; ++rep) {
        const float us0 = run<0, 3>(t, a, b, o), us1 = run<1, 3>(t, a, b, o);
        printf("LDGSTS ring: %.1f us (%.0f GB/s of compulsory reads)   bulk ring: %.1f us (%.0f GB/s)\n", us0, bytes / us0 * 1e-3, us1,
               bytes / us1 * 1e-3);
        const float u2 = run<1, 2>(t, a, b, o), u4 = run<1, 4>(t, a, b, o), u6 = run<1, 6>(t, a, b, o), l4 = run<0, 4>(t, a, b, o);
        printf("bulk ring 2 / 4 / 6 stages: %.1f / %.1f / %.1f us   LDGSTS 4 stages: %.1f us\n", u2, u4, u6, l4);
        printf("COHERENT=%d: bulk 2 stages unmerged %.1f us, merged triples %.1f us\n", COHERENT, run<1, 2>(t, a, b, o), run<1, 2, true>(t, a, b, o));
        printf("bulk 2 / 3 stages, one copy per lane (10 lanes): %.1f / %.1f us\n", run<1, 2, false, true>(t, a, b, o), run<1, 3, false, true>(t, a, b, o));
        printf("bulk, one pixel per step, 2 / 3 / 4 stages: %.1f / %.1f / %.1f us   LDGSTS 2 stages: %.1f us\n", run1<2>(t, a, b, o),
               run1<3>(t, a, b, o), run1<4>(t, a, b, o), run<0, 2>(t, a, b, o));
    }
    return 0;
}
