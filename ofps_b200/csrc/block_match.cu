// K1 + K2: exhaustive block matching (SAD / SSD) with the MotionEntry epilogue fused in.
//
// Specification: SURVEY.md §8c / oracle/ofps_oracle.h (orc_block_match).  The reference has
// no SAD search (its motion vectors come from the H.264 encoder through av-decoder); the
// output convention follows av-decoder/src/lib.rs:404-419.
//
// Tuned kernel layout (one CTA = one row of TBX blocks of one frame pair):
//   * the prev search window ((TBX*B + 2R) x (B + 2R) bytes) is staged once into shared
//     memory as u32 words, then expanded into 4 byte-shifted copies so that every candidate
//     column offset reads ALIGNED words (no per-candidate funnel shifts in the hot loop);
//   * the cur tile is stored transposed [block][word column][row] so a thread pulls the 16
//     rows of one word column with LDS.128;
//   * a work item = (block, dx, group of G consecutive dy).  It walks the B+G-1 prev rows of
//     its column once; each loaded word feeds up to min(B,G) VABSDIFF4.U8.ACC accumulators
//     (4 abs-diff + add per lane per instruction), all indices static after unrolling;
//   * winners are packed as cost<<27 | d2<<14 | (dy+R)<<7 | (dx+R) and reduced with a single
//     u64 min: warp shuffles, then one shared-memory atomicMin per (warp, block).
//   * epilogue (K2): the elected thread unpacks the key and writes mv, cost and the
//     normalised MotionEntry with the reference's operation order (1/W first, then multiply).
#include "block_match_common.cuh"

namespace ofpsb {

namespace {

using namespace bm;

template <int B, int R, int G, int TBX, int NT>
struct TileCfg {
    static constexpr int ND = 2 * R + 1;                 // candidates per axis
    static constexpr int NG = (ND + G - 1) / G;          // dy groups
    static constexpr int RA = (R + 3) & ~3;              // window x origin rounded to a word
    static constexpr int XPAD = RA - R;
    static constexpr int WCOLS = B / 4;                  // words per block row
    static constexpr int WIN_W = TBX * B + 2 * RA;       // bytes per window row
    static constexpr int WIN_H_VALID = B + 2 * R;
    static constexpr int WIN_H = B + NG * G - 1;         // rows incl. padding for masked dy
    static constexpr int ROW_WORDS = WIN_W / 4 + 1;      // +1: source word for the shifted copies
    static constexpr int COPY_RAW = WIN_H * ROW_WORDS;
    static constexpr int COPY_WORDS = COPY_RAW + ((8 - (COPY_RAW % 32)) + 32) % 32;  // == 8 (mod 32 banks)
    static constexpr int CUR_WORDS = TBX * WCOLS * B;
    static constexpr int ITEMS = TBX * ND * NG;
    static constexpr int ROUNDS = (ITEMS + NT - 1) / NT;
    static constexpr size_t SMEM_BYTES = (size_t)(4 * COPY_WORDS + CUR_WORDS) * 4 + (size_t)TBX * 8;
    static_assert(B % 4 == 0 && B >= 4 && B <= 32, "block must be a multiple of 4");
    static_assert(G >= 1 && R >= 0 && R <= 63, "bad search geometry");
};

template <int B, int R, int G, int TBX, int NT, int METRIC>
__global__ void __launch_bounds__(NT) block_match_tile_kernel(const BlockMatchParams p, const bool aligned)
{
    using C = TileCfg<B, R, G, TBX, NT>;
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t* win = smem;                                   // [4][WIN_H][ROW_WORDS]
    uint32_t* curt = smem + 4 * C::COPY_WORDS;              // [TBX][WCOLS][B]
    unsigned long long* best = reinterpret_cast<unsigned long long*>(curt + C::CUR_WORDS);  // [TBX]

    const int tid = threadIdx.x;
    const int tile_bx0 = blockIdx.x * TBX;
    const int by = blockIdx.y;
    const int pair = blockIdx.z;
    const uint8_t* prev = p.prev + (long long)pair * p.pair_stride;
    const uint8_t* cur = p.cur + (long long)pair * p.pair_stride;
    const int x0 = tile_bx0 * B;
    const int y0 = by * B;
    const int win_x0 = x0 - C::RA;
    const int win_y0 = y0 - R;

    // ---- stage the prev window (copy 0) and the transposed cur tile
    for (int idx = tid; idx < C::WIN_H * C::ROW_WORDS; idx += NT) {
        const int row = idx / C::ROW_WORDS, k = idx - row * C::ROW_WORDS;
        const int y = win_y0 + row;
        uint32_t v = 0;
        if (row < C::WIN_H_VALID && y >= -p.halo_top && y < p.strip_h + p.halo_bottom)
            v = load_word(prev + (long long)y * p.stride, win_x0 + 4 * k, p.w, aligned);
        win[idx] = v;
    }
    for (int idx = tid; idx < C::CUR_WORDS; idx += NT) {
        const int r = idx / (TBX * C::WCOLS), k = idx - r * (TBX * C::WCOLS);
        const int b = k / C::WCOLS, c = k - b * C::WCOLS;
        curt[(b * C::WCOLS + c) * B + r] = load_word(cur + (long long)(y0 + r) * p.stride, x0 + 4 * k, p.w, aligned);
    }
    if (tid < TBX) best[tid] = KEY_MAX;
    __syncthreads();
    // ---- byte-shifted copies 1..3: copy_s[row][k] = bytes [4k+s, 4k+s+4) of the row
    for (int idx = tid; idx < 3 * C::WIN_H * C::ROW_WORDS; idx += NT) {
        const int s = idx / (C::WIN_H * C::ROW_WORDS) + 1;
        const int rem = idx - (s - 1) * (C::WIN_H * C::ROW_WORDS);
        const int k = rem % C::ROW_WORDS;
        const uint32_t lo = win[rem];
        const uint32_t hi = (k + 1 < C::ROW_WORDS) ? win[rem + 1] : 0u;
        win[s * C::COPY_WORDS + rem] = __funnelshift_r(lo, hi, 8 * s);
    }
    __syncthreads();

    // ---- work items
#pragma unroll 1
    for (int round = 0; round < C::ROUNDS; round++) {
        int item = round * NT + tid;
        const bool item_ok = item < C::ITEMS;
        if (!item_ok) item = C::ITEMS - 1;
        const int g = item / (TBX * C::ND);
        const int j = item - g * (TBX * C::ND);
        const int b = j / C::ND;
        const int dxi = j - b * C::ND;
        const int xoff = b * B + dxi + C::XPAD;
        const int dyi0 = g * G;

        uint32_t acc[G];
#pragma unroll
        for (int i = 0; i < G; i++) acc[i] = 0;

        const uint32_t* wbase = win + (xoff & 3) * C::COPY_WORDS + (xoff >> 2) + dyi0 * C::ROW_WORDS;
#pragma unroll 1
        for (int c = 0; c < C::WCOLS; c++) {
            uint32_t cw[B];
            const uint4* cp = reinterpret_cast<const uint4*>(curt + (b * C::WCOLS + c) * B);
#pragma unroll
            for (int q = 0; q < B / 4; q++) {
                const uint4 v = cp[q];
                cw[4 * q] = v.x; cw[4 * q + 1] = v.y; cw[4 * q + 2] = v.z; cw[4 * q + 3] = v.w;
            }
            const uint32_t* wp = wbase + c;
#pragma unroll
            for (int rr = 0; rr < B + G - 1; rr++) {
                const uint32_t pw = wp[rr * C::ROW_WORDS];
#pragma unroll
                for (int gi = 0; gi < G; gi++) {
                    const int r = rr - gi;   // cur row matched against this prev row for dy index dyi0+gi
                    if (r >= 0 && r < B) acc[gi] = cost4<METRIC>(cw[r], pw, acc[gi]);
                }
            }
        }

        // ---- local winner over this item's G candidates
        const int bx = tile_bx0 + b;
        const int dx = dxi - R;
        const int px = bx * B + dx;
        const bool x_ok = item_ok && bx < p.nbx && px >= 0 && px + B <= p.w;
        unsigned long long key = KEY_MAX;
#pragma unroll
        for (int gi = 0; gi < G; gi++) {
            const int dyi = dyi0 + gi;
            const int dy = dyi - R;
            const int py = y0 + dy;
            const bool ok = x_ok && dyi < C::ND && py >= -p.halo_top && py + B <= p.strip_h + p.halo_bottom;
            const unsigned long long k = pack_key(acc[gi], dx, dy, R);
            if (ok && k < key) key = k;
        }

        // ---- warp-segmented min, one shared atomic per (warp, block)
        const unsigned lane = tid & 31;
        const bool has = key != KEY_MAX;
        unsigned pending = __ballot_sync(0xffffffffu, has);
        while (pending) {
            const int leader = __ffs(pending) - 1;
            const int lb = __shfl_sync(0xffffffffu, b, leader);
            const bool mine = has && b == lb;
            unsigned long long k = warp_min_u64(mine ? key : KEY_MAX);
            if ((int)lane == leader) atomicMin(&best[lb], k);
            pending &= ~__ballot_sync(0xffffffffu, mine);
        }
    }
    __syncthreads();

    if (tid < TBX) {
        const int bx = tile_bx0 + tid;
        if (bx < p.nbx) {
            const size_t out_idx = (size_t)pair * p.nbx * p.nby + (size_t)by * p.nbx + bx;
            write_block_outputs(p, out_idx, best[tid], bx, by);
        }
    }
}

// Generic kernel: any (block, range); one warp per block, lanes stride over candidates,
// pixels read straight from global/L2.  Correctness reference on the device and the path for
// geometries without a tuned instance.
__global__ void __launch_bounds__(256) block_match_generic_kernel(const BlockMatchParams p)
{
    const int warps_per_cta = blockDim.x / 32;
    const int lane = threadIdx.x & 31;
    const long long blocks_per_pair = (long long)p.nbx * p.nby;
    const long long gw = (long long)blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
    if (gw >= blocks_per_pair * p.n_pairs) return;
    const int pair = (int)(gw / blocks_per_pair);
    const long long bi = gw - (long long)pair * blocks_per_pair;
    const int by = (int)(bi / p.nbx), bx = (int)(bi - (long long)by * p.nbx);
    const uint8_t* prev = p.prev + (long long)pair * p.pair_stride;
    const uint8_t* cur = p.cur + (long long)pair * p.pair_stride;
    const int B = p.block, R = p.range, ND = 2 * R + 1;
    const int x0 = bx * B, y0 = by * B;
    unsigned long long key = KEY_MAX;
    for (int c = lane; c < ND * ND; c += 32) {
        const int dy = c / ND - R, dx = c % ND - R;
        const int px = x0 + dx, py = y0 + dy;
        if (px < 0 || px + B > p.w || py < -p.halo_top || py + B > p.strip_h + p.halo_bottom) continue;
        uint32_t acc = 0;
        for (int y = 0; y < B; y++) {
            const uint8_t* cr = cur + (long long)(y0 + y) * p.stride + x0;
            const uint8_t* pr = prev + (long long)(py + y) * p.stride + px;
            for (int x = 0; x < B; x++) {
                const int d = (int)__ldg(cr + x) - (int)__ldg(pr + x);
                acc += p.metric == OFPSB_METRIC_SAD ? (uint32_t)abs(d) : (uint32_t)(d * d);
            }
        }
        const unsigned long long k = pack_key(acc, dx, dy, R);
        if (k < key) key = k;
    }
    key = warp_min_u64(key);
    if (lane == 0) write_block_outputs(p, (size_t)gw, key, bx, by);
}

template <int B, int R, int G, int TBX, int NT>
int launch_tile(const BlockMatchParams& p, cudaStream_t stream)
{
    using C = TileCfg<B, R, G, TBX, NT>;
    const bool aligned = ((reinterpret_cast<uintptr_t>(p.prev) | reinterpret_cast<uintptr_t>(p.cur) |
                           (uintptr_t)p.stride | (uintptr_t)p.pair_stride) & 3) == 0;
    dim3 grid((p.nbx + TBX - 1) / TBX, p.nby, p.n_pairs);
    if (p.metric == OFPSB_METRIC_SAD) {
        auto k = block_match_tile_kernel<B, R, G, TBX, NT, OFPSB_METRIC_SAD>;
        OFPSB_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
        k<<<grid, NT, C::SMEM_BYTES, stream>>>(p, aligned);
    } else {
        auto k = block_match_tile_kernel<B, R, G, TBX, NT, OFPSB_METRIC_SSD>;
        OFPSB_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
        k<<<grid, NT, C::SMEM_BYTES, stream>>>(p, aligned);
    }
    OFPSB_CUDA_TRY(cudaGetLastError());
    return OFPSB_OK;
}

int check_params(const BlockMatchParams& p)
{
    if (!p.prev || !p.cur || p.w <= 0 || p.strip_h <= 0 || p.stride < p.w || p.n_pairs <= 0 || p.block < 4 ||
        p.block > 64 || (p.block & 3) || p.range < 0 || p.range > 63 ||
        (p.metric != OFPSB_METRIC_SAD && p.metric != OFPSB_METRIC_SSD) || p.halo_top < 0 || p.halo_bottom < 0 ||
        p.full_h <= 0) {
        set_error("block_match: invalid arguments (w=%d h=%d stride=%d block=%d range=%d metric=%d pairs=%d)", p.w,
                  p.strip_h, p.stride, p.block, p.range, p.metric, p.n_pairs);
        return OFPSB_E_INVALID;
    }
    if (p.nby > 65535 || p.n_pairs > 65535) {
        set_error("block_match: grid too large (nby=%d pairs=%d)", p.nby, p.n_pairs);
        return OFPSB_E_INVALID;
    }
    return OFPSB_OK;
}

}  // namespace

int launch_block_match_generic(const BlockMatchParams& p, cudaStream_t stream, uint64_t* launches)
{
    if (int rc = check_params(p)) return rc;
    if (p.nbx == 0 || p.nby == 0) return OFPSB_OK;
    const long long warps = (long long)p.nbx * p.nby * p.n_pairs;
    const int wpc = 8;
    block_match_generic_kernel<<<(unsigned)((warps + wpc - 1) / wpc), wpc * 32, 0, stream>>>(p);
    OFPSB_CUDA_TRY(cudaGetLastError());
    if (launches) ++*launches;
    return OFPSB_OK;
}

int launch_block_match_ldg(const BlockMatchParams& p, cudaStream_t stream, uint64_t* launches)
{
    if (int rc = check_params(p)) return rc;
    if (p.nbx == 0 || p.nby == 0) return OFPSB_OK;
    int rc = 1;
    //                                   B   R   G  TBX  NT
    if (p.block == 16 && p.range == 16) rc = launch_tile<16, 16, 17, 8, 544>(p, stream);
    else if (p.block == 16 && p.range == 8) rc = launch_tile<16, 8, 17, 15, 256>(p, stream);
    else if (p.block == 16 && p.range == 32) rc = launch_tile<16, 32, 13, 4, 672>(p, stream);
    else if (p.block == 8 && p.range == 32) rc = launch_tile<8, 32, 22, 8, 544>(p, stream);
    else if (p.block == 8 && p.range == 16) rc = launch_tile<8, 16, 17, 16, 544>(p, stream);
    else if (p.block == 8 && p.range == 8) rc = launch_tile<8, 8, 17, 15, 256>(p, stream);
    if (rc == 1) return launch_block_match_generic(p, stream, launches);
    if (rc == OFPSB_OK && launches) ++*launches;
    return rc;
}

// Dispatch: TMA-staged instance -> LDG-staged instance -> generic kernel.
int launch_block_match(const BlockMatchParams& p, cudaStream_t stream, uint64_t* launches, int variant)
{
    if (int rc = check_params(p)) return rc;
    if (p.nbx == 0 || p.nby == 0) return OFPSB_OK;
    const int rc = launch_block_match_tma(p, stream, variant);
    if (rc == 1) return launch_block_match_ldg(p, stream, launches);
    if (rc == OFPSB_OK && launches) ++*launches;
    return rc;
}

}  // namespace ofpsb
