// Device helpers shared by the block-matching kernels (block_match.cu, block_match_tma.cu).
#pragma once
#include "common.cuh"

namespace ofpsb {
namespace bm {

constexpr unsigned long long KEY_MAX = ~0ull;

__device__ __forceinline__ uint32_t sad4_acc(uint32_t a, uint32_t b, uint32_t acc)
{
#ifdef OFPSB_EMU
    return __dp4a(__vabsdiffu4(a, b), 0x01010101u, acc);
#else
    uint32_t r;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(acc));
    return r;
#endif
}

__device__ __forceinline__ uint32_t ssd4_acc(uint32_t a, uint32_t b, uint32_t acc)
{
    uint32_t d = __vabsdiffu4(a, b);
    return __dp4a(d, d, acc);
}

template <int METRIC>
__device__ __forceinline__ uint32_t cost4(uint32_t a, uint32_t b, uint32_t acc)
{
    return METRIC == OFPSB_METRIC_SAD ? sad4_acc(a, b, acc) : ssd4_acc(a, b, acc);
}

__device__ __forceinline__ unsigned long long pack_key(uint32_t cost, int dx, int dy, int range)
{
    return ((unsigned long long)cost << 27) | ((unsigned long long)(dx * dx + dy * dy) << 14) |
           ((unsigned long long)(dy + range) << 7) | (unsigned long long)(dx + range);
}

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, v, o);
        v = other < v ? other : v;
    }
    return v;
}

// One u32 word of row `row_ptr` starting at pixel x (may be partly or wholly outside [0,w)).
__device__ __forceinline__ uint32_t load_word(const uint8_t* row_ptr, int x, int w, bool aligned)
{
    if (x >= 0 && x + 4 <= w) {
        if (aligned) return __ldg(reinterpret_cast<const uint32_t*>(row_ptr + x));
        return (uint32_t)__ldg(row_ptr + x) | ((uint32_t)__ldg(row_ptr + x + 1) << 8) |
               ((uint32_t)__ldg(row_ptr + x + 2) << 16) | ((uint32_t)__ldg(row_ptr + x + 3) << 24);
    }
    uint32_t v = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int xx = x + i;
        if (xx >= 0 && xx < w) v |= (uint32_t)__ldg(row_ptr + xx) << (8 * i);
    }
    return v;
}

// K2: block winner -> outputs (av-decoder/src/lib.rs:404-419 convention).
__device__ __forceinline__ void write_block_outputs(const BlockMatchParams& p, size_t out_idx, unsigned long long key,
                                                    int bx, int by)
{
    const int range = p.range;
    const int dx = (int)(key & 127) - range;
    const int dy = (int)((key >> 7) & 127) - range;
    if (p.mv_xy) {
        p.mv_xy[2 * out_idx] = (int16_t)dx;
        p.mv_xy[2 * out_idx + 1] = (int16_t)dy;
    }
    if (p.cost) p.cost[out_idx] = (uint32_t)(key >> 27);
    if (p.entries) {
        const float nx = __fdiv_rn(1.0f, (float)p.w);
        const float ny = __fdiv_rn(1.0f, (float)p.full_h);
        const int src_x = bx * p.block + p.block / 2 + dx;
        const int src_y = p.y_offset + by * p.block + p.block / 2 + dy;
        ofps_mv e;
        e.px = __fmul_rn((float)src_x, nx);
        e.py = __fmul_rn((float)src_y, ny);
        e.mx = __fmul_rn((float)dx, -nx);
        e.my = __fmul_rn((float)dy, -ny);
        p.entries[out_idx] = e;
    }
}


}  // namespace bm
}  // namespace ofpsb
