// TMA (cp.async.bulk.tensor) + mbarrier wrappers and the host-side tensor-map encoder.
//
// Image tiles with their halo are staged into shared memory by ONE elected thread issuing 3-D box
// copies (x, y, plane) -- the hardware generates the addresses, zero-fills outside the image and signals an
// mbarrier with the byte count; the reflect cells of nn.ReflectionPad2d are patched by the CTA afterwards
// (border tiles only).  SASS: UTMALDG + SYNCS.  No libcuda link dependency: cuTensorMapEncodeTiled is
// resolved through cudaGetDriverEntryPoint at first use.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tdl {

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 3-D tiled box load: coordinates (x, y, z) of the box origin in elements; out-of-range elements are zero-filled
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
        : "memory");
}
#endif  // __CUDACC__

// fp32 tensor viewed as (planes, H, W) -> tensor map with a (box_w, box_h, box_p) box.  Returns false when the
// tensor cannot be described (alignment / stride rules of TMA) -- callers then use the plain-load path.
inline bool encode_image_map(CUtensorMap* map, const float* base, int planes, int H, int W, int box_w, int box_h,
                             int box_p) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static const EncodeFn fn = [] {          // resolved once (thread-safe static initialisation)
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            return reinterpret_cast<EncodeFn>(ptr);
        return static_cast<EncodeFn>(nullptr);
    }();
    if (!fn) return false;
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (W & 3) != 0 || ((box_w * 4) & 15) != 0) return false;
    if (box_w > 256 || box_h > 256 || box_p > 256) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
    const cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_p};
    const cuuint32_t estr[3] = {1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct PhotoMaps {                     // kernel parameter (__grid_constant__)
    CUtensorMap tgt;                   // target (B*3, H, W)
    CUtensorMap img[TDL_MAX_SCALES * TDL_MAX_SRC];   // fwd: src[f] in [f]; bwd: warped[s][f] in [s*S+f]
};

}  // namespace tdl
