// tma_host.cuh -- host-side tensor-map encoding shared by the kernels that use TMA (gemm_tc.cu, attn_decode.cu).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D row-major matrix [rows, cols] of `elem_bytes`-byte elements, box = [box_rows, box_cols].
inline int make_map_2d(CUtensorMap *map, const void *base, CUtensorMapDataType dt, int elem_bytes, uint64_t rows,
                uint64_t cols, uint32_t box_rows, uint32_t box_cols, CUtensorMapSwizzle swz, uint64_t pitch_elems = 0,
                CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B) {
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) {
        wq_set_error("cuTensorMapEncodeTiled is not available from the driver");
        return WQ_ERR_CUDA;
    }
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {(pitch_elems ? pitch_elems : cols) * (uint64_t)elem_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, dt, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                    promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        wq_set_error("cuTensorMapEncodeTiled failed (%d): rows %llu cols %llu elem %d box %ux%u", (int)r,
                     (unsigned long long)rows, (unsigned long long)cols, elem_bytes, box_rows, box_cols);
        return WQ_ERR_CUDA;
    }
    return WQ_OK;
}


}  // namespace
