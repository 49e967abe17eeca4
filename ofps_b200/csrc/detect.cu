// K4: BlockMotionDetection::detect_motion after densification
// (block-motion-detector/src/lib.rs:63-118).
//
// The reference thresholds the cell means, then flood-fills 8-connected islands in row-major
// seed order with a stack, keeps the island with strictly greater cell count (ties -> the
// earliest seed), and returns a field holding the island's means EXCEPT at the seed cell, which
// is never `set_motion`'d (:80) — a quirk kept here.  The flood fill order does not influence the
// result: an island is a connected component, its seed is the component's smallest row-major
// index, its area the component size.  So the kernel labels components with a lock-free
// union-find (links always point from the larger root to the smaller, hence the final root IS
// the seed), counts areas, and picks max(area, then smallest seed) with one 64-bit atomicMax.
// One CTA; the grid is 14x14 by default and 160x160 at the UI bounds (dynamic beyond that).
#include "common.cuh"

namespace ofpsb {

namespace {

constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr int DET_NT = 1024;

__device__ __forceinline__ uint32_t uf_find(volatile uint32_t* parent, uint32_t x)
{
    uint32_t p = parent[x];
    while (p != x) {
        x = p;
        p = parent[x];
    }
    return x;
}

__device__ __forceinline__ void uf_union(volatile uint32_t* parent, uint32_t a, uint32_t b)
{
    for (;;) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) { const uint32_t t = a; a = b; b = t; }
        const uint32_t old = atomicMin(const_cast<uint32_t*>(parent) + a, b);   // a > b
        if (old == a) return;   // a was still a root: linked
        a = old;                // someone re-rooted a meanwhile: retry from its new parent
    }
}

__global__ void __launch_bounds__(DET_NT) detect_kernel(const float* __restrict__ mean, uint32_t dim, float target_motion,
                                                        float min_size, float* __restrict__ out_field,
                                                        DetectResult* __restrict__ result, uint32_t* parent_g,
                                                        uint32_t* area_g, int use_smem)
{
    extern __shared__ uint32_t dyn[];
    __shared__ unsigned long long best;
    const uint32_t cells = dim * dim;
    uint32_t* parent_p = use_smem ? dyn : parent_g;
    uint32_t* area = use_smem ? dyn + cells : area_g;
    volatile uint32_t* parent = parent_p;
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) best = 0ull;

    // map[y][x] = |mean| >= target_motion, magnitude = sqrt(x*x + y*y) un-fused (:63-68)
    for (uint32_t c = tid; c < cells; c += nt) {
        const float mx = mean[2 * c], my = mean[2 * c + 1];
        const float mag = __fsqrt_rn(__fadd_rn(__fmul_rn(mx, mx), __fmul_rn(my, my)));
        parent[c] = (mag >= target_motion) ? c : NONE;
        area[c] = 0;
        out_field[2 * c] = 0.0f;
        out_field[2 * c + 1] = 0.0f;
    }
    __syncthreads();
    // 8-connectivity: uniting each cell with its 4 raster-earlier neighbours covers every edge
    for (uint32_t c = tid; c < cells; c += nt) {
        if (parent[c] == NONE) continue;
        const uint32_t x = c % dim, y = c / dim;
        if (x > 0 && parent[c - 1] != NONE) uf_union(parent, c, c - 1);
        if (y > 0) {
            const uint32_t up = c - dim;
            if (parent[up] != NONE) uf_union(parent, c, up);
            if (x > 0 && parent[up - 1] != NONE) uf_union(parent, c, up - 1);
            if (x + 1 < dim && parent[up + 1] != NONE) uf_union(parent, c, up + 1);
        }
    }
    __syncthreads();
    // flatten: every cell points at its root (= the island's row-major seed), so the passes below are O(1) per cell
    uint32_t root = NONE;
    {
        const uint32_t c = tid;
        if (cells <= nt && c < cells && parent[c] != NONE) root = uf_find(parent, c);
    }
    if (cells <= nt) {
        __syncthreads();
        if (tid < cells && root != NONE) parent[tid] = root;
        __syncthreads();
    }
    for (uint32_t c = tid; c < cells; c += nt) {
        if (parent[c] == NONE) continue;
        const uint32_t r = uf_find(parent, c);
        atomicAdd(&area[r], 1u);
    }
    __syncthreads();
    // strictly greater area wins, ties -> earliest seed (:106-109)
    for (uint32_t c = tid; c < cells; c += nt) {
        if (parent[c] == c) atomicMax(&best, ((unsigned long long)area[c] << 32) | (unsigned long long)(NONE - c));
    }
    __syncthreads();
    const unsigned long long bk = best;
    const uint32_t best_area = (uint32_t)(bk >> 32);
    const uint32_t seed = NONE - (uint32_t)(bk & 0xFFFFFFFFull);
    // biggest_area as f32 / (dim*dim) as f32 >= min_size (:114)
    const bool some = best_area > 0 &&
                      __fdiv_rn((float)(unsigned long long)best_area, (float)(unsigned long long)cells) >= min_size;
    if (some) {
        for (uint32_t c = tid; c < cells; c += nt) {
            if (parent[c] == NONE || c == seed) continue;   // the seed cell keeps zero motion (:80)
            if (uf_find(parent, c) == seed) {
                out_field[2 * c] = mean[2 * c];
                out_field[2 * c + 1] = mean[2 * c + 1];
            }
        }
    }
    if (tid == 0) {
        result->best_key = bk;
        result->area = some ? best_area : 0u;
        result->seed_cell = some ? seed : NONE;
        result->has_motion = some ? 1 : 0;
        result->pad = 0;
    }
}

// ---- grids up to 32 x 32 (the default 14 x 14; `subdivide` <= 7 at the default `min_size`): one warp, no atomics.
// Lane r holds row r of the threshold map as a 32-bit mask.  Islands are grown from the row-major first unvisited cell
// by dilating the island mask with its 8-neighbourhood (the rows above / below arrive by shuffle) until it stops
// changing; islands are found in seed order, so "strictly greater area wins" is the reference's tie rule (:106-109).
// ncu r2: the union-find kernel above took 35 us for 196 cells (dependent shared-memory atomics); this one is a few
// hundred warp instructions.
constexpr int DETS_NT = 128;

__global__ void __launch_bounds__(DETS_NT) detect_small_kernel(const float* __restrict__ mean, uint32_t dim, float target_motion,
                                                               float min_size, float* __restrict__ out_field,
                                                               DetectResult* __restrict__ result)
{
    __shared__ uint32_t s_bits[32];      // threshold bits, 32 consecutive cells per word
    __shared__ uint32_t s_island[32];    // winning island, one mask per row
    __shared__ uint32_t s_res[3];        // area, seed, some
    const uint32_t cells = dim * dim, tid = threadIdx.x, lane = tid & 31;
    // map[y][x] = |mean| >= target_motion, magnitude = sqrt(x*x + y*y) un-fused (:63-68)
    for (uint32_t base = (tid >> 5) * 32; base < cells; base += DETS_NT) {
        const uint32_t c = base + lane;
        bool bit = false;
        if (c < cells) {
            const float2 m = *reinterpret_cast<const float2*>(mean + 2 * c);
            bit = __fsqrt_rn(__fadd_rn(__fmul_rn(m.x, m.x), __fmul_rn(m.y, m.y))) >= target_motion;
            *reinterpret_cast<float2*>(out_field + 2 * c) = make_float2(0.0f, 0.0f);
        }
        const unsigned b = __ballot_sync(0xffffffffu, bit);
        if (lane == 0) s_bits[base >> 5] = b;
    }
    __syncthreads();
    if (tid < 32) {
        const uint32_t rowmask = dim == 32 ? 0xFFFFFFFFu : (1u << dim) - 1u;
        uint32_t map = 0;
        if (lane < dim) {   // row `lane` = cells [lane*dim, lane*dim + dim): at most two words
            const uint32_t c0 = lane * dim, w = c0 >> 5, sh = c0 & 31;
            const uint32_t lo = s_bits[w], hi = (w + 1 < 32 && ((cells + 31) >> 5) > w + 1) ? s_bits[w + 1] : 0u;
            map = __funnelshift_r(lo, hi, sh) & rowmask;
        }
        uint32_t left = map, best_island = 0, best_area = 0, best_seed = NONE;
        for (;;) {
            // row-major first unvisited cell
            const unsigned rows = __ballot_sync(0xffffffffu, left != 0u);
            if (!rows) break;
            const int sr = __ffs(rows) - 1;
            const uint32_t srow = __shfl_sync(0xffffffffu, left, sr);
            const uint32_t sx = (uint32_t)(__ffs(srow) - 1);
            uint32_t isl = lane == (uint32_t)sr ? (1u << sx) : 0u;
            for (;;) {
                uint32_t up = __shfl_up_sync(0xffffffffu, isl, 1), dn = __shfl_down_sync(0xffffffffu, isl, 1);
                if (lane == 0) up = 0;
                if (lane == 31) dn = 0;
                uint32_t g = isl | up | dn;
                g |= (g << 1) | (g >> 1);
                g &= left;   // one step in all eight directions, inside the not-yet-visited set
                const uint32_t grown = g | isl;
                const bool changed = grown != isl;
                isl = grown;
                if (!__any_sync(0xffffffffu, changed)) break;
            }
            left &= ~isl;
            const uint32_t area = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(isl));
            if (area > best_area) {   // strictly greater: ties keep the earliest seed
                best_area = area;
                best_island = isl;
                best_seed = (uint32_t)sr * dim + sx;
            }
        }
        // biggest_area as f32 / (dim*dim) as f32 >= min_size (:114)
        const bool some = best_area > 0 &&
                          __fdiv_rn((float)(unsigned long long)best_area, (float)(unsigned long long)cells) >= min_size;
        s_island[lane] = some ? best_island : 0u;
        if (lane == 0) {
            s_res[0] = some ? best_area : 0u;
            s_res[1] = some ? best_seed : NONE;
            result->best_key = ((unsigned long long)best_area << 32) | (unsigned long long)(NONE - best_seed);
            result->area = some ? best_area : 0u;
            result->seed_cell = some ? best_seed : NONE;
            result->has_motion = some ? 1 : 0;
            result->pad = 0;
        }
    }
    __syncthreads();
    const uint32_t seed = s_res[1];
    for (uint32_t c = tid; c < cells; c += DETS_NT) {
        const uint32_t y = c / dim, x = c - y * dim;
        if (((s_island[y] >> x) & 1u) && c != seed) {   // the seed cell keeps zero motion (:80)
            *reinterpret_cast<float2*>(out_field + 2 * c) = *reinterpret_cast<const float2*>(mean + 2 * c);
        }
    }
}

}  // namespace

int launch_detect(const float* d_mean_field, size_t dim, float target_motion, float min_size, float* d_out_field,
                  DetectResult* d_result, DevBuf& scratch, cudaStream_t stream, uint64_t* launches, int force_union_find)
{
    if (dim == 0 || dim > 65535) {
        set_error("detect: block_dim %zu out of range (1..65535)", dim);
        return OFPSB_E_INVALID;
    }
    const size_t cells = dim * dim;
    if (dim <= 32 && !force_union_find) {
        detect_small_kernel<<<1, DETS_NT, 0, stream>>>(d_mean_field, (uint32_t)dim, target_motion, min_size, d_out_field,
                                                       d_result);
        OFPSB_CUDA_TRY(cudaGetLastError());
        if (launches) ++*launches;
        return OFPSB_OK;
    }
    const size_t bytes = cells * 8;
    const int use_smem = bytes <= 200 * 1024;
    uint32_t *parent = nullptr, *area = nullptr;
    if (!use_smem) {
        if (int rc = scratch.reserve(bytes)) return rc;
        parent = scratch.as<uint32_t>();
        area = parent + cells;
    }
    if (use_smem && bytes > 48 * 1024)
        OFPSB_CUDA_TRY(cudaFuncSetAttribute(detect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    // 14 x 14 cells by default: a CTA sized to the grid (every barrier and the launch itself scale with the CTA)
    const unsigned nt = cells >= (size_t)DET_NT ? (unsigned)DET_NT : (unsigned)((cells + 31) & ~(size_t)31);
    detect_kernel<<<1, nt < 64 ? 64 : nt, use_smem ? bytes : 0, stream>>>(d_mean_field, (uint32_t)dim, target_motion, min_size,
                                                               d_out_field, d_result, parent, area, use_smem);
    OFPSB_CUDA_TRY(cudaGetLastError());
    if (launches) ++*launches;
    return OFPSB_OK;
}

}  // namespace ofpsb
