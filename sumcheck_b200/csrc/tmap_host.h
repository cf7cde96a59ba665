// Host side of the TMA path: CUtensorMap descriptors for the evaluation tables.  A table of L field elements is seen
// as a 2-D tensor of L/4 rows x 128 bytes (one row = old[4b..4b+3], what one output pair of a fold round consumes);
// a box is 128 rows = 16 KiB, written to shared memory with the 128-byte swizzle that both the threads
// (conflict-free LDS.128) and the tensor core (SWIZZLE_128B K-major operand) read.
// The driver entry point is resolved at run time (no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace tmaph {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// rows x 128 bytes at `base` (device pointer, 16-byte aligned), boxes of `box_rows` rows.  Returns false when the
// descriptor cannot be built (driver too old, misaligned borrowed table, too many rows) — callers then keep the
// non-TMA kernels.
inline bool make_table_map(CUtensorMap* out, const void* base, uint64_t rows, uint32_t box_rows) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc || rows == 0 || rows >= ((uint64_t)1 << 31) || ((uintptr_t)base & 15)) return false;
    const cuuint64_t dims[2] = {32, rows};       // 32 x u32 = 128 bytes per row
    const cuuint64_t strides[1] = {128};         // bytes between rows
    const cuuint32_t box[2] = {32, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// rows x 64 bytes (one PAIR of field elements per row) with SWIZZLE_64B: what lands in shared memory is the MN-major operand of
// the tensor-core contraction (gemm_sum.cuh), the pair index being its K dimension.
inline bool make_pair_map(CUtensorMap* out, const void* base, uint64_t rows, uint32_t box_rows) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc || rows == 0 || rows >= ((uint64_t)1 << 31) || ((uintptr_t)base & 15)) return false;
    const cuuint64_t dims[2] = {16, rows};
    const cuuint64_t strides[1] = {64};
    const cuuint32_t box[2] = {16, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tmaph
