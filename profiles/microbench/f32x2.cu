// Micro-benchmark: packed fp32 (FADD2 / FMUL2 / FFMA2, PTX add/mul/fma.rn.f32x2) issue rate vs scalar FADD / FMUL on sm_100a,
// and whether ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (it must not for the parity arithmetic).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 pk(float lo, float hi) { return ((u64)__float_as_uint(hi) << 32) | __float_as_uint(lo); }

constexpr int ITERS = 4096, ACC = 8;
template <int MODE> __global__ void bench(float* out, float seed) {
    float r = 0.f;
    if (MODE == 0) {          // scalar FADD, 2*ACC independent chains  (same flop count as MODE 1)
        float a[2 * ACC];
        for (int k = 0; k < 2 * ACC; ++k) a[k] = seed + k + threadIdx.x;
        for (int i = 0; i < ITERS; ++i)
#pragma unroll
            for (int k = 0; k < 2 * ACC; ++k) a[k] = __fadd_rn(a[k], seed);
        for (int k = 0; k < 2 * ACC; ++k) r += a[k];
    } else if (MODE == 1) {   // FADD2, ACC independent chains
        u64 a[ACC]; const u64 s2 = pk(seed, seed);
        for (int k = 0; k < ACC; ++k) a[k] = pk(seed + k + threadIdx.x, seed - k);
        for (int i = 0; i < ITERS; ++i)
#pragma unroll
            for (int k = 0; k < ACC; ++k) a[k] = add2(a[k], s2);
        for (int k = 0; k < ACC; ++k) r += __uint_as_float((unsigned)a[k]) + __uint_as_float((unsigned)(a[k] >> 32));
    } else if (MODE == 2) {   // scalar FMUL+FADD un-contracted
        float a[2 * ACC];
        for (int k = 0; k < 2 * ACC; ++k) a[k] = seed + k + threadIdx.x;
        for (int i = 0; i < ITERS / 2; ++i)
#pragma unroll
            for (int k = 0; k < 2 * ACC; ++k) a[k] = __fadd_rn(__fmul_rn(a[k], seed), seed);
        for (int k = 0; k < 2 * ACC; ++k) r += a[k];
    } else if (MODE == 3) {   // FMUL2 + FADD2 (volatile asm: not contracted)
        u64 a[ACC]; const u64 s2 = pk(seed, seed);
        for (int k = 0; k < ACC; ++k) a[k] = pk(seed + k + threadIdx.x, seed - k);
        for (int i = 0; i < ITERS / 2; ++i)
#pragma unroll
            for (int k = 0; k < ACC; ++k) a[k] = add2(mul2(a[k], s2), s2);
        for (int k = 0; k < ACC; ++k) r += __uint_as_float((unsigned)a[k]) + __uint_as_float((unsigned)(a[k] >> 32));
    } else if (MODE == 4) {   // scalar FFMA
        float a[2 * ACC];
        for (int k = 0; k < 2 * ACC; ++k) a[k] = seed + k + threadIdx.x;
        for (int i = 0; i < ITERS; ++i)
#pragma unroll
            for (int k = 0; k < 2 * ACC; ++k) a[k] = fmaf(a[k], seed, seed);
        for (int k = 0; k < 2 * ACC; ++k) r += a[k];
    } else {                  // FFMA2
        u64 a[ACC]; const u64 s2 = pk(seed, seed);
        for (int k = 0; k < ACC; ++k) a[k] = pk(seed + k + threadIdx.x, seed - k);
        for (int i = 0; i < ITERS; ++i)
#pragma unroll
            for (int k = 0; k < ACC; ++k) a[k] = fma2(a[k], s2, s2);
        for (int k = 0; k < ACC; ++k) r += __uint_as_float((unsigned)a[k]) + __uint_as_float((unsigned)(a[k] >> 32));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// contraction check (non-volatile asm so that ptxas is free to do what it would do in a real kernel)
__device__ __forceinline__ u64 add2n(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2n(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__global__ void contract(const float* x, float* o) {
    const float a = x[0], b = x[1], c = x[2];
    const u64 r = add2n(mul2n(pk(a, a), pk(b, b)), pk(c, c));
    o[0] = __uint_as_float((unsigned)r);
    o[1] = __fadd_rn(__fmul_rn(a, b), c);
    o[2] = fmaf(a, b, c);
}
template <int MODE> void run(const char* name, float* d) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 8, nt = 256;
    bench<MODE><<<grid, nt>>>(d, 1.0001f); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) bench<MODE><<<grid, nt>>>(d, 1.0001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
    const double lane_ops = (double)grid * nt * ITERS * 2 * ACC;     // fp32 lane-operations (add, mul+add pair counts 1 per op)
    printf("%-28s %8.3f ms  %8.1f G lane-ops/s  (%.2f per SM-clk at 1.965 GHz)\n", name, ms, lane_ops / ms * 1e-6,
           lane_ops / ms * 1e-6 / 148 / 1.965);
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("FADD scalar", d); run<1>("FADD2 packed", d);
    run<2>("FMUL+FADD scalar", d); run<3>("FMUL2+FADD2 packed", d);
    run<4>("FFMA scalar", d); run<5>("FFMA2 packed", d);
    float hx[3] = {1.0000001f, 1.0000001f, -1.0000002f}, ho[3];
    float *dx, *dout; cudaMalloc(&dx, 12); cudaMalloc(&dout, 12);
    cudaMemcpy(dx, hx, 12, cudaMemcpyHostToDevice);
    contract<<<1, 1>>>(dx, dout); cudaMemcpy(ho, dout, 12, cudaMemcpyDeviceToHost);
    printf("contraction check: packed mul.rn+add.rn = %.9g, scalar un-fused = %.9g, fma = %.9g -> %s\n", ho[0], ho[1], ho[2],
           ho[0] == ho[1] ? "NOT contracted (ok)" : "CONTRACTED");
    return 0;
}
