// q4_tma.h -- host-side helper: 2-D TMA tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point, so the
// library does not link libcuda).  Shared by the prefill GEMM (q4_gemm.cu) and the decode GEMV's weight ring (q4_gemv.cu).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

namespace q4 {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn tma_encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// [outer, inner] row-major tensor of `dt` elements with `row_bytes` between rows; box = box_outer rows x box_inner elements
inline bool make_map_2d(CUtensorMap* m, CUtensorMapDataType dt, const void* base, uint64_t inner, uint64_t outer,
                        uint64_t row_bytes, uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle sw)
{
    EncodeTiledFn fn = tma_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {row_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    return fn(m, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
#ifdef Q4_TMA_L2_PROMOTION
              Q4_TMA_L2_PROMOTION,
#else
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
#endif
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace q4
