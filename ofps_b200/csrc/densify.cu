// K3: MotionFieldDensifier::add_vector over a slice + `MotionField::from(densifier)`.
//
// Reference: ofps/src/motion_field.rs:133-190 (new / add_vector_weighted / add_vector_idx) and
// :297-308 (sum ./ counts).  The reference adds the vectors of one cell in INPUT ORDER with
// un-fused f32 arithmetic, so the result depends on the order of additions; atomics are out.
// Both paths below keep the reference order and are bit-exact:
//
//   * scan path   (cells * n small — the detector's 14x14 grid over a block-match field): one
//     warp per cell streams the whole entry list, ballots the lanes whose entry lands in its cell
//     and folds them in lane (= input) order.  One launch, entries read coalesced as float4.
//   * sort path   (anything bigger): cell ids -> stable LSD radix sort of (cell, index) with
//     8-bit digits (hand-written: per-tile match.any ranking, tile histograms, exclusive scan,
//     scatter) -> segment boundaries -> one thread / one warp per cell walks its segment in order.
//
// All float ops use the explicit round-to-nearest intrinsics: no FMA contraction, IEEE divide.
#include "common.cuh"

namespace ofpsb {

namespace {

// `f32 as usize` (saturating, NaN -> 0) after f32::round (half away from zero = roundf).
__device__ __forceinline__ unsigned long long f32_as_usize(float v)
{
    if (!(v > 0.0f)) return 0ull;
    if (v >= 18446744073709551616.0f) return ~0ull;
    return (unsigned long long)v;
}

// nalgebra::clamp(pos, (0,0), (1,1)) with Point2's all-components partial order
// (motion_field.rs:170; see oracle/ofps_oracle.c clamp_point) then cell lookup (:171-176).
__device__ __forceinline__ unsigned long long cell_of(float px, float py, float wm1, float hm1, unsigned long long gw)
{
    if (px > 0.0f && py > 0.0f) {
        if (!(px < 1.0f && py < 1.0f)) { px = 1.0f; py = 1.0f; }
    } else {
        px = 0.0f; py = 0.0f;
    }
    const unsigned long long x = f32_as_usize(roundf(__fmul_rn(px, wm1)));
    const unsigned long long y = f32_as_usize(roundf(__fmul_rn(py, hm1)));
    return y * gw + x;
}

__device__ __forceinline__ void finish_cell(float sx, float sy, float cx, float cy, size_t cell, float* field,
                                            float* counts, int raw)
{
    // MotionField::from(densifier): component_div (motion_field.rs:297-308); raw != 0 keeps the densifier's
    // un-divided sums (the state interpolate_empty_cells works on, motion_field.rs:193-294)
    field[2 * cell] = raw ? sx : __fdiv_rn(sx, cx);
    field[2 * cell + 1] = raw ? sy : __fdiv_rn(sy, cy);
    if (counts) {
        counts[2 * cell] = cx;
        counts[2 * cell + 1] = cy;
    }
}

constexpr float F32_EPSILON = 1.1920928955078125e-07f;

// ------------------------------------------------------------------ scan path
__global__ void __launch_bounds__(256) densify_scan_kernel(const ofps_mv* __restrict__ entries, size_t n, size_t gw,
                                                           size_t gh, float* __restrict__ field,
                                                           float* __restrict__ counts, int raw)
{
    const unsigned lane = threadIdx.x & 31;
    const size_t cell = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const size_t cells = gw * gh;
    if (cell >= cells) return;
    const float wm1 = (float)(unsigned long long)(gw - 1), hm1 = (float)(unsigned long long)(gh - 1);
    float sx = 0.0f, sy = 0.0f, cx = F32_EPSILON, cy = F32_EPSILON;
    const float4* e4 = reinterpret_cast<const float4*>(entries);
    constexpr int DEPTH = 8;   // independent 16-byte loads in flight per lane: the scan is latency-bound otherwise
    for (size_t base = 0; base < n; base += 32 * DEPTH) {
        float4 e[DEPTH];
#pragma unroll
        for (int k = 0; k < DEPTH; k++) {
            const size_t i = base + (size_t)k * 32 + lane;
            e[k] = i < n ? __ldg(e4 + i) : make_float4(-1.f, -1.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < DEPTH; k++) {
            const size_t i = base + (size_t)k * 32 + lane;
            const bool hit = i < n && cell_of(e[k].x, e[k].y, wm1, hm1, gw) == cell;
            unsigned m = __ballot_sync(0xffffffffu, hit);
            while (m) {   // fold the hits in lane (= input) order
                const int l = __ffs(m) - 1;
                m &= m - 1;
                const float mx = __shfl_sync(0xffffffffu, e[k].z, l);
                const float my = __shfl_sync(0xffffffffu, e[k].w, l);
                // add_vector_idx (motion_field.rs:141-147), weight = 1
                cx = __fadd_rn(cx, 1.0f);
                cy = __fadd_rn(cy, 1.0f);
                sx = __fadd_rn(__fmul_rn(mx, 1.0f), sx);
                sy = __fadd_rn(__fmul_rn(my, 1.0f), sy);
            }
        }
    }
    if (lane == 0) finish_cell(sx, sy, cx, cy, cell, field, counts, raw);
}

// Scan path for small inputs (the detector's case: a few thousand vectors into 14 x 14 cells).  The kernel above
// recomputes the cell of every entry in every warp (ncu r2: 15,800 warp instructions per cell for 8,040 entries, 57 us,
// latency-bound).  Here every CTA first stages the cell ids of ALL entries in shared memory (u16, one pass of the whole
// CTA), then each warp scans the ids for its cell — one shared load, one compare and one ballot per 32 entries — and
// touches the entries themselves only at its hits, still folded in input order.
constexpr size_t SCAN_IDS_MAX = 48 * 1024;   // entries whose u16 ids fit 96 KB of dynamic shared memory

__global__ void __launch_bounds__(256) densify_scan_ids_kernel(const ofps_mv* __restrict__ entries, uint32_t n, size_t gw,
                                                               size_t gh, float* __restrict__ field,
                                                               float* __restrict__ counts, int raw)
{
    extern __shared__ uint16_t s_ids[];
    const unsigned lane = threadIdx.x & 31;
    const size_t cells = gw * gh;
    const float wm1 = (float)(unsigned long long)(gw - 1), hm1 = (float)(unsigned long long)(gh - 1);
    const float4* e4 = reinterpret_cast<const float4*>(entries);
#pragma unroll 8
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {   // unrolled: eight independent 16-byte loads in flight
        const float4 e = __ldg(e4 + i);
        s_ids[i] = (uint16_t)cell_of(e.x, e.y, wm1, hm1, gw);
    }
    __syncthreads();
    const size_t cell = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (cell >= cells) return;
    float sx = 0.0f, sy = 0.0f, cx = F32_EPSILON, cy = F32_EPSILON;
#pragma unroll 4
    for (uint32_t base = 0; base < n; base += 32) {
        const uint32_t i = base + lane;
        const bool hit = i < n && (size_t)s_ids[i] == cell;
        unsigned m = __ballot_sync(0xffffffffu, hit);
        while (m) {   // fold the hits in lane (= input) order
            const int l = __ffs(m) - 1;
            m &= m - 1;
            const float4 e = __ldg(e4 + base + l);
            // add_vector_idx (motion_field.rs:141-147), weight = 1
            cx = __fadd_rn(cx, 1.0f);
            cy = __fadd_rn(cy, 1.0f);
            sx = __fadd_rn(__fmul_rn(e.z, 1.0f), sx);
            sy = __fadd_rn(__fmul_rn(e.w, 1.0f), sy);
        }
    }
    if (lane == 0) finish_cell(sx, sy, cx, cy, cell, field, counts, raw);
}

// ------------------------------------------------------------------ sort path
__global__ void __launch_bounds__(256) cell_id_kernel(const ofps_mv* __restrict__ entries, size_t n, size_t gw, size_t gh,
                                                      uint32_t* __restrict__ keys, uint32_t* __restrict__ vals)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float wm1 = (float)(unsigned long long)(gw - 1), hm1 = (float)(unsigned long long)(gh - 1);
    const float4 e = __ldg(reinterpret_cast<const float4*>(entries) + i);
    keys[i] = (uint32_t)cell_of(e.x, e.y, wm1, hm1, gw);
    vals[i] = (uint32_t)i;
}

constexpr int RS_TILE = 1024;   // one element per thread; tile order = warp-major, lane-minor = input order

__global__ void __launch_bounds__(RS_TILE) radix_hist_kernel(const uint32_t* __restrict__ keys, size_t n, int shift,
                                                             uint32_t* __restrict__ hist, unsigned nblk)
{
    __shared__ uint32_t h[256];
    if (threadIdx.x < 256) h[threadIdx.x] = 0;
    __syncthreads();
    const size_t i = (size_t)blockIdx.x * RS_TILE + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
    __syncthreads();
    if (threadIdx.x < 256) hist[(size_t)threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];
}

// Exclusive scan of `len` counters in place, one CTA (len = 256 * tiles is at most a few 1e5).
__global__ void __launch_bounds__(1024) exclusive_scan_kernel(uint32_t* __restrict__ data, size_t len)
{
    __shared__ uint32_t part[1024];
    const size_t per = (len + 1023) / 1024;
    const size_t b = (size_t)threadIdx.x * per;
    const size_t e = b + per < len ? b + per : len;
    uint32_t s = 0;
    for (size_t i = b; i < e; i++) s += data[i];
    part[threadIdx.x] = s;
    __syncthreads();
    // Hillis-Steele inclusive scan over the 1024 partials
    for (int off = 1; off < 1024; off <<= 1) {
        const uint32_t v = threadIdx.x >= (unsigned)off ? part[threadIdx.x - off] : 0u;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = threadIdx.x ? part[threadIdx.x - 1] : 0u;
    for (size_t i = b; i < e; i++) {
        const uint32_t v = data[i];
        data[i] = run;
        run += v;
    }
}

__global__ void __launch_bounds__(RS_TILE) radix_scatter_kernel(const uint32_t* __restrict__ keys_in,
                                                                const uint32_t* __restrict__ vals_in, size_t n, int shift,
                                                                const uint32_t* __restrict__ hist, unsigned nblk,
                                                                uint32_t* __restrict__ keys_out,
                                                                uint32_t* __restrict__ vals_out)
{
    __shared__ uint32_t cnt[32][256];   // [warp][digit]: peers before this warp in the tile
    __shared__ uint32_t gbase[256];
    for (int k = threadIdx.x; k < 32 * 256; k += RS_TILE) (&cnt[0][0])[k] = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t i = (size_t)blockIdx.x * RS_TILE + threadIdx.x;
    const bool valid = i < n;
    uint32_t key = 0, val = 0, digit = 0, rank = 0;
    const unsigned active = __ballot_sync(0xffffffffu, valid);
    if (valid) {
        key = keys_in[i];
        val = vals_in[i];
        digit = (key >> shift) & 255u;
        const unsigned peers = __match_any_sync(active, digit);
        rank = __popc(peers & ((1u << lane) - 1u));
        if (rank == 0) cnt[warp][digit] = __popc(peers);
    }
    __syncthreads();
    if (threadIdx.x < 256) {
        uint32_t run = 0;
        for (int w = 0; w < 32; w++) {
            const uint32_t t = cnt[w][threadIdx.x];
            cnt[w][threadIdx.x] = run;
            run += t;
        }
        gbase[threadIdx.x] = hist[(size_t)threadIdx.x * nblk + blockIdx.x];
    }
    __syncthreads();
    if (valid) {
        const size_t pos = (size_t)gbase[digit] + cnt[warp][digit] + rank;
        keys_out[pos] = key;
        vals_out[pos] = val;
    }
}

__global__ void __launch_bounds__(256) segment_bounds_kernel(const uint32_t* __restrict__ keys, size_t n,
                                                             uint32_t* __restrict__ seg_start,
                                                             uint32_t* __restrict__ seg_end)
{
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t k = keys[p];
    if (p == 0 || keys[p - 1] != k) seg_start[k] = (uint32_t)p;
    if (p + 1 == n || keys[p + 1] != k) seg_end[k] = (uint32_t)(p + 1);
}

// LANES = 1: one thread per cell; LANES = 32: one warp per cell (coalesced gathers, fold in lane order).
template <int LANES>
__global__ void __launch_bounds__(256) segment_sum_kernel(const ofps_mv* __restrict__ entries,
                                                          const uint32_t* __restrict__ vals,
                                                          const uint32_t* __restrict__ seg_start,
                                                          const uint32_t* __restrict__ seg_end, size_t cells,
                                                          float* __restrict__ field, float* __restrict__ counts, int raw)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t cell = t / LANES;
    if (cell >= cells) return;
    const unsigned lane = LANES == 32 ? (threadIdx.x & 31) : 0;
    const uint32_t b = seg_start[cell], e = seg_end[cell];
    float sx = 0.0f, sy = 0.0f, cx = F32_EPSILON, cy = F32_EPSILON;
    const float4* e4 = reinterpret_cast<const float4*>(entries);
    if (LANES == 1) {
        for (uint32_t p = b; p < e; p++) {
            const float4 v = __ldg(e4 + vals[p]);
            cx = __fadd_rn(cx, 1.0f);
            cy = __fadd_rn(cy, 1.0f);
            sx = __fadd_rn(__fmul_rn(v.z, 1.0f), sx);
            sy = __fadd_rn(__fmul_rn(v.w, 1.0f), sy);
        }
        finish_cell(sx, sy, cx, cy, cell, field, counts, raw);
    } else {
        for (uint32_t base = b; base < e; base += 32) {
            const uint32_t p = base + lane;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p < e) v = __ldg(e4 + vals[p]);
            const int cntl = (int)min(32u, e - base);
            for (int l = 0; l < cntl; l++) {
                const float mx = __shfl_sync(0xffffffffu, v.z, l);
                const float my = __shfl_sync(0xffffffffu, v.w, l);
                cx = __fadd_rn(cx, 1.0f);
                cy = __fadd_rn(cy, 1.0f);
                sx = __fadd_rn(__fmul_rn(mx, 1.0f), sx);
                sy = __fadd_rn(__fmul_rn(my, 1.0f), sy);
            }
        }
        if (lane == 0) finish_cell(sx, sy, cx, cy, cell, field, counts, raw);
    }
}

__global__ void __launch_bounds__(256) empty_field_kernel(size_t cells, float* __restrict__ field,
                                                          float* __restrict__ counts, int raw)
{
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cells) return;
    finish_cell(0.0f, 0.0f, F32_EPSILON, F32_EPSILON, c, field, counts, raw);
}

}  // namespace

int launch_densify(const ofps_mv* d_entries, size_t n, size_t gw, size_t gh, float* d_field, float* d_counts,
                   DensifyScratch& s, cudaStream_t stream, uint64_t* launches, int force_path, int raw)
{
    if (gw == 0 || gh == 0 || gw > (1u << 24) || gh > (1u << 24) || gw * gh > 0xFFFFFFF0ull || n > 0xFFFFFFF0ull) {
        set_error("densify: invalid grid %zux%zu or n=%zu (grid sides 1..2^24, cells and n < 2^32)", gw, gh, n);
        return OFPSB_E_INVALID;
    }
    const size_t cells = gw * gh;
    if (n == 0) {
        empty_field_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, stream>>>(cells, d_field, d_counts, raw);
        OFPSB_CUDA_TRY(cudaGetLastError());
        if (launches) ++*launches;
        return OFPSB_OK;
    }
    const bool scan = force_path == 1 || (force_path == 0 && (double)cells * (double)n <= 48.0e6);
    if (scan && n <= SCAN_IDS_MAX && cells <= 65535) {
        const size_t smem = (n * 2 + 15) & ~(size_t)15;
        static bool attr_set[64] = {};
        int dev = 0;
        OFPSB_CUDA_TRY(cudaGetDevice(&dev));
        if (smem > 48 * 1024 && (dev < 0 || dev >= 64 || !attr_set[dev])) {
            OFPSB_CUDA_TRY(cudaFuncSetAttribute(densify_scan_ids_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)(SCAN_IDS_MAX * 2)));
            if (dev >= 0 && dev < 64) attr_set[dev] = true;
        }
        densify_scan_ids_kernel<<<(unsigned)((cells + 7) / 8), 256, smem, stream>>>(d_entries, (uint32_t)n, gw, gh, d_field,
                                                                                 d_counts, raw);
        OFPSB_CUDA_TRY(cudaGetLastError());
        if (launches) ++*launches;
        return OFPSB_OK;
    }
    if (scan) {
        densify_scan_kernel<<<(unsigned)((cells + 7) / 8), 256, 0, stream>>>(d_entries, n, gw, gh, d_field, d_counts, raw);
        OFPSB_CUDA_TRY(cudaGetLastError());
        if (launches) ++*launches;
        return OFPSB_OK;
    }
    // ---- sort path
    const unsigned nblk = (unsigned)((n + RS_TILE - 1) / RS_TILE);
    if (int rc = s.keys_a.reserve(n * 4)) return rc;
    if (int rc = s.keys_b.reserve(n * 4)) return rc;
    if (int rc = s.vals_a.reserve(n * 4)) return rc;
    if (int rc = s.vals_b.reserve(n * 4)) return rc;
    if (int rc = s.hist.reserve((size_t)nblk * 256 * 4)) return rc;
    if (int rc = s.cell_start.reserve(cells * 8)) return rc;
    uint32_t *ka = s.keys_a.as<uint32_t>(), *kb = s.keys_b.as<uint32_t>();
    uint32_t *va = s.vals_a.as<uint32_t>(), *vb = s.vals_b.as<uint32_t>();
    uint32_t* hist = s.hist.as<uint32_t>();
    uint32_t* seg_start = s.cell_start.as<uint32_t>();
    uint32_t* seg_end = seg_start + cells;
    uint64_t nl = 0;
    cell_id_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_entries, n, gw, gh, ka, va);
    nl++;
    int bits = 0;
    while (bits < 32 && ((size_t)1 << bits) < cells) bits++;
    for (int shift = 0; shift < bits; shift += 8) {
        radix_hist_kernel<<<nblk, RS_TILE, 0, stream>>>(ka, n, shift, hist, nblk);
        exclusive_scan_kernel<<<1, 1024, 0, stream>>>(hist, (size_t)nblk * 256);
        radix_scatter_kernel<<<nblk, RS_TILE, 0, stream>>>(ka, va, n, shift, hist, nblk, kb, vb);
        nl += 3;
        uint32_t* t = ka; ka = kb; kb = t;
        t = va; va = vb; vb = t;
    }
    OFPSB_CUDA_TRY(cudaMemsetAsync(seg_start, 0, cells * 8, stream));
    segment_bounds_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(ka, n, seg_start, seg_end);
    nl++;
    if (n / cells >= 8) {
        const size_t threads = cells * 32;
        segment_sum_kernel<32><<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(d_entries, va, seg_start, seg_end,
                                                                                    cells, d_field, d_counts, raw);
    } else {
        segment_sum_kernel<1><<<(unsigned)((cells + 255) / 256), 256, 0, stream>>>(d_entries, va, seg_start, seg_end,
                                                                                  cells, d_field, d_counts, raw);
    }
    nl++;
    OFPSB_CUDA_TRY(cudaGetLastError());
    if (launches) *launches += nl;
    return OFPSB_OK;
}

}  // namespace ofpsb
