// Host-side TMA descriptor helpers: cuTensorMapEncodeTiled is fetched through the runtime (no libcuda link, so
// the library also loads on a CPU-only box).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <mutex>

namespace mg {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode() {
    static EncodeTiledFn fn_cached = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn_cached = reinterpret_cast<EncodeTiledFn>(fn);
    });
    return fn_cached;
}

inline CUtensorMapSwizzle swz_enum(int bytes) {
    return bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// 4-D map over an NHWC fp16 tensor [N,H,W,C]: box {box_c, box_w, box_h, 1}, element strides {1, sx, sy, 1}.
inline CUresult encode_nhwc(CUtensorMap* tm, const void* ptr, int N, int H, int W, int C, int box_c, int box_w, int box_h,
                            int sx, int sy, int swizzle_bytes) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)sx, (cuuint32_t)sy, 1};
    return get_encode()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz_enum(swizzle_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

}  // namespace mg
