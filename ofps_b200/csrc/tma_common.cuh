// TMA plumbing shared by the block-matching kernels: tensor-map encoding on the host (driver entry point
// fetched through the runtime, no libcuda link), box loads and mbarrier waits on the device.
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>

#include "block_match_common.cuh"

namespace ofpsb {
namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int x, int y, int z, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar)
        : "memory");
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
        : "memory");
}

__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global [%0, {%1, %2, %3, %4}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

// L2 prefetch of a box (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int x, int y, int z)
{
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"(map), "r"(x), "r"(y), "r"(z) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t b32, uint32_t parity)
{
    uint32_t done = 0;
    while (!done) {
        asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2; selp.u32 %0, 1, 0, q; }"
                     : "=r"(done)
                     : "r"(b32), "r"(parity)
                     : "memory");
    }
}

inline PFN_cuTensorMapEncodeTiled get_encode()
{
    // C++11 magic static: initialised exactly once even when contexts live on different host threads
    static const PFN_cuTensorMapEncodeTiled fn = [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            return reinterpret_cast<PFN_cuTensorMapEncodeTiled>(ptr);
        return static_cast<PFN_cuTensorMapEncodeTiled>(nullptr);
    }();
    return fn;
}

// u8 tensor (x: w valid bytes of each `stride`-byte row, y: rows, z: pairs)
inline bool make_map(CUtensorMap* map, const uint8_t* base, int w, int rows, long long stride, long long pair_stride, int pairs,
              int box_w, int box_h)
{
    PFN_cuTensorMapEncodeTiled enc = get_encode();
    if (!enc) return false;
    if (pairs <= 1 || pair_stride <= 0) pair_stride = ((stride * (long long)rows) + 15) & ~15ll;
    cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)rows, (cuuint64_t)(pairs > 0 ? pairs : 1)};
    cuuint64_t strides[2] = {(cuuint64_t)stride, (cuuint64_t)pair_stride};
    cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}


// "Transposing" u8 tensor: dims {16 bytes, rows, 16-byte column blocks, pairs} with byte strides {row stride, 16, pair
// stride}; a box {16, box_rows, box_cols16, 1} lands in shared memory as [column block][row][16 bytes] — the 16 rows of a
// 16-pixel-wide block are then 256 contiguous bytes (measured on B200: the driver accepts the non-ascending strides and
// out-of-tensor rows / column blocks arrive as zeros, tools/tma_transpose_test.cu).
inline bool make_map_colblocks(CUtensorMap* map, const uint8_t* base, int w, int rows, long long stride, long long pair_stride,
                               int pairs, int box_rows, int box_cols16)
{
    PFN_cuTensorMapEncodeTiled enc = get_encode();
    if (!enc) return false;
    if (pairs <= 1 || pair_stride <= 0) pair_stride = ((stride * (long long)rows) + 15) & ~15ll;
    cuuint64_t dims[4] = {16, (cuuint64_t)rows, (cuuint64_t)((w + 15) / 16), (cuuint64_t)(pairs > 0 ? pairs : 1)};
    cuuint64_t strides[3] = {(cuuint64_t)stride, 16, (cuuint64_t)pair_stride};
    cuuint32_t box[4] = {16, (cuuint32_t)box_rows, (cuuint32_t)box_cols16, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<uint8_t*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Same layout described with `elem_bytes`-wide unsigned elements (2 or 4): lets one box span up to
// 256 elements instead of 256 bytes.  w_elems / box_w are in elements, strides in bytes.
inline bool make_map_elems(CUtensorMap* map, int elem_bytes, const void* base, long long w_elems, long long rows,
                           long long stride, long long plane_stride, int planes, int box_w, int box_h)
{
    PFN_cuTensorMapEncodeTiled enc = get_encode();
    if (!enc) return false;
    if (planes <= 1 || plane_stride <= 0) plane_stride = ((stride * rows) + 15) & ~15ll;
    cuuint64_t dims[3] = {(cuuint64_t)w_elems, (cuuint64_t)rows, (cuuint64_t)(planes > 0 ? planes : 1)};
    cuuint64_t strides[2] = {(cuuint64_t)stride, (cuuint64_t)plane_stride};
    cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapDataType dt = elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32
                                                   : (elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8);
    return enc(map, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tma
}  // namespace ofpsb
