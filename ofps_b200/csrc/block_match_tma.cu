// K1+K2, TMA-staged variant: exhaustive block matching where the frame tiles are moved from
// HBM/L2 into shared memory by the Tensor Memory Accelerator instead of by SM instructions.
//
// Why: on B200 `VABSDIFF4.U8.ACC` issues at 64 lanes/clk/SM on the ALU pipe (measured,
// profiles/r1_microbench_int_pipes.txt) and an exhaustive SAD search is bound by exactly that pipe.
// In the first tile kernel (block_match.cu) 57 % of the executed instructions were NOT the SAD
// instruction: window staging with bounds checks, the funnel-shift pass that builds the four
// byte-shifted window copies, 64-bit key packing (profiles/r1_block_match_v1_ncu.md).  Here
//   * the search window and the current tile are two TMA box loads (u8 tensor maps; out-of-frame
//     bytes are zero-filled by the TMA unit, so there is no bounds logic at all).  The three
//     byte-shifted window copies are then derived in shared memory with one funnel shift per word
//     (measured on B200: a TMA box must start on a 16-byte boundary of the innermost dimension —
//     a u8 box at x+1 raises "illegal instruction", tools/tma_test.cu — so the TMA unit cannot
//     produce the shifted copies itself);
//   * work items are ordered (dy group, shift class, block, dx/4) so that the lanes of a warp read
//     consecutive words of ONE copy: conflict-free without padding the TMA destination;
//   * each thread folds its G candidates into a 32-bit key `cost<<7 | rank(dy)` (one IMAD on the
//     FMA pipe + one VIMNMX per candidate), and the per-block argmin is finished by one warp per
//     block with two REDUX.MIN passes over shared memory.
// The result is bit-identical to the generic kernel and the oracle (same tie-break:
// lexicographic (cost, dx^2+dy^2, dy, dx)).
#include "tma_common.cuh"

namespace ofpsb {

namespace {

using namespace bm;
using namespace tma;

template <int B, int R, int G, int TBX, int NT>
struct TmaCfg {
    static constexpr int ND = 2 * R + 1;
    static constexpr int NG = (ND + G - 1) / G;
    static constexpr int RA = (R + 15) & ~15;                     // window origin on a 16-byte boundary (TMA)
    static constexpr int XPAD = RA - R;
    static constexpr int WCOLS = B / 4;
    static constexpr int WIN_W = TBX * B + 2 * RA;                // bytes per window row (TMA box inner dim)
    static constexpr int ROW_WORDS = WIN_W / 4;
    static constexpr int WIN_H_VALID = B + 2 * R;                 // rows the TMA box carries
    static constexpr int WIN_H = B + NG * G - 1;                  // rows addressed (incl. masked dy padding)
    static constexpr int COPY_BYTES = (WIN_H * WIN_W + 127) & ~127;
    static constexpr int COPY_WORDS = COPY_BYTES / 4;
    static constexpr int CUR_W = TBX * B;                         // bytes per current-tile row
    static constexpr int CUR_BYTES = (B * CUR_W + 127) & ~127;
    static constexpr int CUR_ROW_WORDS = CUR_W / 4;
    static constexpr int FIRST0 = (0 - XPAD) & 3, FIRST1 = (1 - XPAD) & 3, FIRST2 = (2 - XPAD) & 3, FIRST3 = (3 - XPAD) & 3;
    static constexpr int nq(int first) { return first < ND ? (ND - first + 3) / 4 : 0; }
    static constexpr int NQ0 = nq(FIRST0), NQ1 = nq(FIRST1), NQ2 = nq(FIRST2), NQ3 = nq(FIRST3);
    static constexpr int ITEMS_G = TBX * ND;
    static constexpr int ITEMS = NG * ITEMS_G;
    static constexpr int ROUNDS = (ITEMS + NT - 1) / NT;
    static constexpr int SLOTS = TBX * NG * ND;                   // per-item results, block-major
    static constexpr uint32_t TX_BYTES = (uint32_t)WIN_W * WIN_H_VALID + (uint32_t)B * CUR_W;
    static constexpr size_t SMEM_BYTES = 4 * (size_t)COPY_BYTES + CUR_BYTES + (size_t)SLOTS * 8 + 16;
    static_assert(B % 4 == 0 && B >= 4 && B <= 16, "TMA instances cover blocks up to 16x16 (32-bit keys)");
    static_assert(CUR_W % 16 == 0 && WIN_W % 16 == 0 && WIN_W <= 256 && WIN_H_VALID <= 256 && CUR_W <= 256, "TMA box limits");
    static_assert(NQ0 + NQ1 + NQ2 + NQ3 == ND, "shift classes must cover every dx");
    static_assert(ND <= 127 && 2 * (NG * G - R) < 128, "key fields: 7-bit dy rank, 7-bit dx/dy indices");
};

// cost * 128 + rank on the FMA pipe (IMAD): the ALU pipe is the one the SAD instruction saturates
__device__ __forceinline__ uint32_t fold_key(uint32_t cost, uint32_t rank)
{
    uint32_t r;
    asm("mad.lo.u32 %0, %1, 128, %2;" : "=r"(r) : "r"(cost), "r"(rank));
    return r;
}

// Per-block candidate limits of the tile, filled by the kernel prologue: legal dx / dy ranges (a block
// outside the frame or beyond the work list gets an empty dx range).
struct TileLimits {
    int dx_lo, dx_hi, dy_lo, dy_hi;
    int xadj;   // bytes between the 16-byte aligned TMA box origin and the block's window origin (0 or 8)
};

// The SAD work of one tile: every thread takes items (dy group, shift class, block, dx/4), walks its
// window column once and leaves (cost, position code) of its best candidate in shared memory.
// BPITCH = byte distance between the windows of consecutive blocks inside a window row.
template <typename C, int B, int R, int G, int TBX, int NT, int METRIC, int BPITCH>
__device__ __forceinline__ void tile_items(const uint32_t* __restrict__ win, const uint32_t* __restrict__ curs,
                                           const TileLimits* __restrict__ lim, uint32_t* __restrict__ r_cost,
                                           uint32_t* __restrict__ r_pos, int tid)
{
#pragma unroll 1
    for (int round = 0; round < C::ROUNDS; round++) {
        const int item = round * NT + tid;
        if (item >= C::ITEMS) break;
        const int g = item / C::ITEMS_G;
        int t = item - g * C::ITEMS_G;
        int s, b, q;
        if (t < TBX * C::NQ0) { s = 0; b = t / C::NQ0; q = t - b * C::NQ0; }
        else if ((t -= TBX * C::NQ0) < TBX * C::NQ1) { s = 1; b = t / C::NQ1; q = t - b * C::NQ1; }
        else if ((t -= TBX * C::NQ1) < TBX * C::NQ2) { s = 2; b = t / C::NQ2; q = t - b * C::NQ2; }
        else { t -= TBX * C::NQ2; s = 3; b = t / (C::NQ3 > 0 ? C::NQ3 : 1); q = t - b * C::NQ3; }
        const int dxi = 4 * q + ((s - C::XPAD) & 3);
        const TileLimits L = lim[b];
        const int xoff = b * BPITCH + dxi + C::XPAD + L.xadj;     // xoff & 3 == s (xadj is a multiple of 4)
        const int dyi0 = g * G;

        uint32_t acc[G];
#pragma unroll
        for (int i = 0; i < G; i++) acc[i] = 0;
        const uint32_t* wbase = win + s * C::COPY_WORDS + (xoff >> 2) + dyi0 * C::ROW_WORDS;
        const uint32_t* cbase = curs + b * C::WCOLS;
#pragma unroll 1
        for (int c = 0; c < C::WCOLS; c++) {
            uint32_t cw[B];
#pragma unroll
            for (int r = 0; r < B; r++) cw[r] = cbase[r * C::CUR_ROW_WORDS + c];
            const uint32_t* wp = wbase + c;
#pragma unroll
            for (int rr = 0; rr < B + G - 1; rr++) {
                const uint32_t pw = wp[rr * C::ROW_WORDS];
#pragma unroll
                for (int gi = 0; gi < G; gi++) {
                    const int r = rr - gi;
                    if (r >= 0 && r < B) acc[gi] = cost4<METRIC>(cw[r], pw, acc[gi]);
                }
            }
        }

        // fold the G candidates of this (block, dx): key = cost<<7 | rank(dy), rank orders (|dy|, dy).
        // The dy group index is made a compile-time constant so that every rank is an immediate, and
        // blocks whose every dy is legal skip the per-candidate range test.
        const int dx = dxi - R;
        const bool x_ok = dx >= L.dx_lo && dx <= L.dx_hi;
        const bool all_dy = L.dy_lo == -R && L.dy_hi == R;
        uint32_t best = 0xFFFFFFFFu;
#pragma unroll
        for (int gg = 0; gg < C::NG; gg++) {
            if (g != gg) continue;
            if (all_dy) {
#pragma unroll
                for (int gi = 0; gi < G; gi++) {
                    const int dy = gg * G + gi - R;
                    if (dy <= R) best = min(best, fold_key(acc[gi], (uint32_t)(2 * (dy < 0 ? -dy : dy) - (dy < 0 ? 1 : 0))));
                }
            } else {
#pragma unroll
                for (int gi = 0; gi < G; gi++) {
                    const int dy = gg * G + gi - R;
                    const uint32_t key = fold_key(acc[gi], (uint32_t)(2 * (dy < 0 ? -dy : dy) - (dy < 0 ? 1 : 0)));
                    if (dy >= L.dy_lo && dy <= L.dy_hi) best = min(best, key);
                }
            }
        }
        uint32_t cost = 0xFFFFFFFFu, pos = 0;
        if (x_ok && best != 0xFFFFFFFFu) {
            const int code = (int)(best & 127u);
            const int ady = (code + 1) >> 1;
            const int dy = (code & 1) ? -ady : ady;
            cost = best >> 7;
            pos = ((uint32_t)(dx * dx + dy * dy) << 14) | ((uint32_t)(dy + R) << 7) | (uint32_t)(dx + R);
        }
        const int slot = (b * C::NG + g) * C::ND + dxi;
        r_cost[slot] = cost;
        r_pos[slot] = pos;
    }
}

// min cost, then min position code among the candidates with that cost: the spec's lexicographic key
template <typename C>
__device__ __forceinline__ unsigned long long block_argmin(const uint32_t* rc, const uint32_t* rp, int lane)
{
    uint32_t cmin = 0xFFFFFFFFu;
    for (int i = lane; i < C::NG * C::ND; i += 32) cmin = min(cmin, rc[i]);
    cmin = __reduce_min_sync(0xffffffffu, cmin);
    uint32_t pmin = 0xFFFFFFFFu;
    for (int i = lane; i < C::NG * C::ND; i += 32)
        if (rc[i] == cmin) pmin = min(pmin, rp[i]);
    pmin = __reduce_min_sync(0xffffffffu, pmin);
    return ((unsigned long long)cmin << 27) | pmin;
}

template <int B, int R, int G, int TBX, int NT, int METRIC>
__global__ void __launch_bounds__(NT) block_match_tma_kernel(const __grid_constant__ CUtensorMap map_prev,
                                                             const __grid_constant__ CUtensorMap map_cur,
                                                             const BlockMatchParams p)
{
    using C = TmaCfg<B, R, G, TBX, NT>;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint32_t* win = reinterpret_cast<uint32_t*>(smem_raw);                         // [4][WIN_H][ROW_WORDS]
    uint32_t* curs = reinterpret_cast<uint32_t*>(smem_raw + 4 * C::COPY_BYTES);    // [B][CUR_ROW_WORDS]
    uint32_t* r_cost = reinterpret_cast<uint32_t*>(smem_raw + 4 * C::COPY_BYTES + C::CUR_BYTES);   // [SLOTS]
    uint32_t* r_pos = r_cost + C::SLOTS;
    uint64_t* bar = reinterpret_cast<uint64_t*>(r_pos + C::SLOTS);
    __shared__ TileLimits lim[TBX];

    const int tid = threadIdx.x;
    const int tile_bx0 = blockIdx.x * TBX;
    const int by = blockIdx.y;
    const int pair = blockIdx.z;
    const int x0 = tile_bx0 * B;
    const int y0 = by * B;

    if (tid == 0) {
        const uint32_t b32 = smem_u32(bar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b32));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b32), "r"(C::TX_BYTES) : "memory");
        const int wx = x0 - C::RA, wy = y0 - R + p.halo_top;   // tensor row 0 = first halo row
        tma_load_3d(smem_u32(win), &map_prev, wx, wy, pair, b32);
        tma_load_3d(smem_u32(curs), &map_cur, x0, y0, pair, b32);
    }
    if (tid < TBX) {
        const int bx = tile_bx0 + tid;
        TileLimits L;
        L.dy_lo = max(-R, -p.halo_top - y0);                          // y0 + dy >= -halo_top
        L.dy_hi = min(R, p.strip_h + p.halo_bottom - B - y0);         // y0 + dy + B <= strip_h + halo_bottom
        L.dx_lo = max(-R, -bx * B);
        L.dx_hi = bx < p.nbx ? min(R, p.w - B - bx * B) : -R - 1;
        L.xadj = 0;
        lim[tid] = L;
    }
    __syncthreads();   // barrier init + limits visible
    mbar_wait(smem_u32(bar), 0);

    // byte-shifted copies 1..3: copy_s[row][k] = bytes [4k+s, 4k+s+4) of the window row
    for (int idx = tid; idx < C::WIN_H_VALID * C::ROW_WORDS; idx += NT) {
        const int k = idx % C::ROW_WORDS;
        const uint32_t lo = win[idx];
        const uint32_t hi = (k + 1 < C::ROW_WORDS) ? win[idx + 1] : 0u;
        win[C::COPY_WORDS + idx] = __funnelshift_r(lo, hi, 8);
        win[2 * C::COPY_WORDS + idx] = __funnelshift_r(lo, hi, 16);
        win[3 * C::COPY_WORDS + idx] = __funnelshift_r(lo, hi, 24);
    }
    __syncthreads();

    tile_items<C, B, R, G, TBX, NT, METRIC, B>(win, curs, lim, r_cost, r_pos, tid);
    __syncthreads();

    const int warp = tid >> 5, lane = tid & 31;
    for (int blk = warp; blk < TBX; blk += NT / 32) {
        const int bx = tile_bx0 + blk;
        const unsigned long long key = block_argmin<C>(r_cost + blk * C::NG * C::ND, r_pos + blk * C::NG * C::ND, lane);
        if (lane == 0 && bx < p.nbx) {
            const size_t out_idx = (size_t)pair * p.nbx * p.nby + (size_t)by * p.nbx + bx;
            write_block_outputs(p, out_idx, key, bx, by);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Work-list variant: exhaustive search of an arbitrary list of blocks (the blocks the pruning pass,
// block_match_prune.cu, could not decide).  Persistent CTAs walk the list TBX blocks at a time; every
// block gets its own (B+2RA) x (B+2R) window box and B x B current box by TMA into a raw staging area,
// and one pass lays them out as the four byte-shifted window copies / the current tile the SAD loop
// expects (same loop, window pitch B+2RA instead of B).
template <int B, int R, int G, int TBX, int NT>
struct ListCfg : TmaCfg<B, R, G, TBX, NT> {
    using Base = TmaCfg<B, R, G, TBX, NT>;
    static constexpr int XADJ_MAX = (B % 16) ? 8 : 0;                      // box origin rounded down to 16 bytes
    static constexpr int WBOX_W = (B + 2 * Base::RA + XADJ_MAX + 15) & ~15;   // bytes per row of one block's window box
    static constexpr int WBOX_WORDS = WBOX_W / 4;
    static constexpr int ROW_WORDS = TBX * WBOX_WORDS;                     // window row of the whole tile
    static constexpr int COPY_BYTES = (Base::WIN_H * ROW_WORDS * 4 + 127) & ~127;
    static constexpr int COPY_WORDS = COPY_BYTES / 4;
    static constexpr int CBOX_W = B < 16 ? 16 : B;                         // TMA inner box >= 16 bytes
    static constexpr int RAW_WIN_BYTES = (WBOX_W * Base::WIN_H_VALID + 127) & ~127;
    static constexpr int RAW_CUR_BYTES = (CBOX_W * B + 127) & ~127;
    static constexpr uint32_t TX_PER_BLOCK = (uint32_t)WBOX_W * Base::WIN_H_VALID + (uint32_t)CBOX_W * B;
    static constexpr size_t SMEM_BYTES = (size_t)TBX * (RAW_WIN_BYTES + RAW_CUR_BYTES) + 4 * (size_t)COPY_BYTES +
                                         Base::CUR_BYTES + (size_t)Base::SLOTS * 8 + 16;
    static_assert(WBOX_W % 16 == 0 && WBOX_W <= 256, "TMA box limits");
};

template <int B, int R, int G, int TBX, int NT, int METRIC>
__global__ void __launch_bounds__(NT) block_match_list_kernel(const __grid_constant__ CUtensorMap map_prev,
                                                              const __grid_constant__ CUtensorMap map_cur,
                                                              const BlockMatchParams p,
                                                              const uint32_t* __restrict__ list,
                                                              const uint32_t* __restrict__ list_count)
{
    using C = ListCfg<B, R, G, TBX, NT>;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* raw_win = smem_raw;                                                        // [TBX][RAW_WIN_BYTES]
    uint8_t* raw_cur = raw_win + TBX * C::RAW_WIN_BYTES;                                // [TBX][RAW_CUR_BYTES]
    uint32_t* win = reinterpret_cast<uint32_t*>(raw_cur + TBX * C::RAW_CUR_BYTES);     // [4][WIN_H][ROW_WORDS]
    uint32_t* curs = win + 4 * C::COPY_WORDS;                                           // [B][CUR_ROW_WORDS]
    uint32_t* r_cost = curs + C::CUR_BYTES / 4;
    uint32_t* r_pos = r_cost + C::SLOTS;
    uint64_t* bar = reinterpret_cast<uint64_t*>(r_pos + C::SLOTS);
    __shared__ TileLimits lim[2][TBX];
    __shared__ uint32_t s_blk[2][TBX];
    static_assert(TBX <= 32, "tile set-up is done by one warp");

    const int tid = threadIdx.x;
    const uint32_t count = *list_count;
    const uint32_t nblk = (uint32_t)p.nbx * p.nby;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // Warp 0 describes a tile (block coordinates, candidate limits) and starts its TMA loads into the raw
    // staging area; called for tile i+1 as soon as the layout pass of tile i has consumed the staging area,
    // so the loads run behind the SAD loop of tile i.
    auto start_tile = [&](uint32_t tile, int slot) {
        if (tid < 32) {
            if (tid < TBX) {
                const uint32_t li = tile * TBX + tid;
                TileLimits L = {0, -1, 0, -1, 0};
                uint32_t gb = 0xFFFFFFFFu;
                if (li < count) {
                    gb = list[li];
                    const uint32_t rem = gb % nblk;
                    const int by = (int)(rem / p.nbx), bx = (int)(rem % p.nbx);
                    const int y0 = by * B;
                    L.dy_lo = max(-R, -p.halo_top - y0);
                    L.dy_hi = min(R, p.strip_h + p.halo_bottom - B - y0);
                    L.dx_lo = max(-R, -bx * B);
                    L.dx_hi = min(R, p.w - B - bx * B);
                    L.xadj = (bx * B) & 15;
                }
                lim[slot][tid] = L;
                s_blk[slot][tid] = gb;
            }
            __syncwarp();
            if (tid == 0) {
                const uint32_t b32 = smem_u32(bar);
                uint32_t nvalid = 0;
                for (int b = 0; b < TBX; b++) nvalid += s_blk[slot][b] != 0xFFFFFFFFu;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b32), "r"(nvalid * C::TX_PER_BLOCK) : "memory");
                for (int b = 0; b < TBX; b++) {
                    const uint32_t gb = s_blk[slot][b];
                    if (gb == 0xFFFFFFFFu) continue;
                    const int pair = (int)(gb / nblk);
                    const uint32_t rem = gb % nblk;
                    const int by = (int)(rem / p.nbx), bx = (int)(rem % p.nbx);
                    const int xa = (bx * B) & ~15;   // TMA boxes start on a 16-byte boundary
                    tma_load_3d(smem_u32(raw_win + b * C::RAW_WIN_BYTES), &map_prev, xa - C::RA, by * B - R + p.halo_top, pair, b32);
                    tma_load_3d(smem_u32(raw_cur + b * C::RAW_CUR_BYTES), &map_cur, xa, by * B, pair, b32);
                }
            }
        }
    };

    uint32_t parity = 0;
    int slot = 0;
    if ((unsigned long long)blockIdx.x * TBX < count) start_tile(blockIdx.x, 0);
    __syncthreads();
    for (uint32_t tile = blockIdx.x; (unsigned long long)tile * TBX < count; tile += gridDim.x, slot ^= 1) {
        mbar_wait(smem_u32(bar), parity);
        parity ^= 1;
        // lay out the four byte-shifted window copies [s][row][b][word] and the current tile [row][b][word]
        for (int idx = tid; idx < TBX * C::WIN_H_VALID * C::WBOX_WORDS; idx += NT) {
            const int b = idx / (C::WIN_H_VALID * C::WBOX_WORDS);
            const int rem = idx - b * (C::WIN_H_VALID * C::WBOX_WORDS);
            const int row = rem / C::WBOX_WORDS, k = rem - row * C::WBOX_WORDS;
            const uint32_t* src = reinterpret_cast<const uint32_t*>(raw_win + b * C::RAW_WIN_BYTES) + rem;
            const uint32_t lo = src[0];
            const uint32_t hi = (k + 1 < C::WBOX_WORDS) ? src[1] : 0u;
            uint32_t* dst = win + row * C::ROW_WORDS + b * C::WBOX_WORDS + k;
            dst[0] = lo;
            dst[C::COPY_WORDS] = __funnelshift_r(lo, hi, 8);
            dst[2 * C::COPY_WORDS] = __funnelshift_r(lo, hi, 16);
            dst[3 * C::COPY_WORDS] = __funnelshift_r(lo, hi, 24);
        }
        for (int idx = tid; idx < TBX * B * C::WCOLS; idx += NT) {
            const int b = idx / (B * C::WCOLS);
            const int rem = idx - b * (B * C::WCOLS);
            const int row = rem / C::WCOLS, k = rem - row * C::WCOLS;
            curs[row * C::CUR_ROW_WORDS + b * C::WCOLS + k] =
                reinterpret_cast<const uint32_t*>(raw_cur + b * C::RAW_CUR_BYTES)[row * (C::CBOX_W / 4) + k + (lim[slot][b].xadj >> 2)];
        }
        __syncthreads();   // staging consumed: the next tile's loads may overwrite it
        const uint32_t next = tile + gridDim.x;
        if ((unsigned long long)next * TBX < count) start_tile(next, slot ^ 1);

        tile_items<C, B, R, G, TBX, NT, METRIC, C::WBOX_W>(win, curs, lim[slot], r_cost, r_pos, tid);
        __syncthreads();

        const int warp = tid >> 5, lane = tid & 31;
        for (int blk = warp; blk < TBX; blk += NT / 32) {
            const uint32_t gb = s_blk[slot][blk];
            if (gb == 0xFFFFFFFFu) continue;
            const unsigned long long key = block_argmin<C>(r_cost + blk * C::NG * C::ND, r_pos + blk * C::NG * C::ND, lane);
            if (lane == 0) {
                const uint32_t rem = gb % nblk;
                write_block_outputs(p, (size_t)gb, key, (int)(rem % p.nbx), (int)(rem / p.nbx));
            }
        }
        __syncthreads();   // results are reused by the next tile
    }
}

template <int B, int R, int G, int TBX, int NT>
int launch_tma(const BlockMatchParams& p, cudaStream_t stream)
{
    using C = TmaCfg<B, R, G, TBX, NT>;
    const int rows_prev = p.halo_top + p.strip_h + p.halo_bottom;
    const uint8_t* prev_base = p.prev - (long long)p.halo_top * p.stride;
    CUtensorMap mp, mc;
    if (!make_map(&mp, prev_base, p.w, rows_prev, p.stride, p.pair_stride, p.n_pairs, C::WIN_W, C::WIN_H_VALID) ||
        !make_map(&mc, p.cur, p.w, p.strip_h, p.stride, p.pair_stride, p.n_pairs, C::CUR_W, B))
        return 1;   // not expressible as a tensor map: the caller falls back to the LDG-staged kernel
    dim3 grid((p.nbx + TBX - 1) / TBX, p.nby, p.n_pairs);
    if (p.metric == OFPSB_METRIC_SAD) {
        auto k = block_match_tma_kernel<B, R, G, TBX, NT, OFPSB_METRIC_SAD>;
        OFPSB_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
        k<<<grid, NT, C::SMEM_BYTES, stream>>>(mp, mc, p);
    } else {
        auto k = block_match_tma_kernel<B, R, G, TBX, NT, OFPSB_METRIC_SSD>;
        OFPSB_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
        k<<<grid, NT, C::SMEM_BYTES, stream>>>(mp, mc, p);
    }
    OFPSB_CUDA_TRY(cudaGetLastError());
    return OFPSB_OK;
}

template <int B, int R, int G, int TBX, int NT>
int launch_list(const BlockMatchParams& p, const uint32_t* d_list, const uint32_t* d_count, int sm_count, cudaStream_t stream)
{
    using C = ListCfg<B, R, G, TBX, NT>;
    const int rows_prev = p.halo_top + p.strip_h + p.halo_bottom;
    const uint8_t* prev_base = p.prev - (long long)p.halo_top * p.stride;
    CUtensorMap mp, mc;
    if (!make_map(&mp, prev_base, p.w, rows_prev, p.stride, p.pair_stride, p.n_pairs, C::WBOX_W, C::WIN_H_VALID) ||
        !make_map(&mc, p.cur, p.w, p.strip_h, p.stride, p.pair_stride, p.n_pairs, C::CBOX_W, B))
        return 1;
    const long long total = (long long)p.nbx * p.nby * p.n_pairs;
    long long tiles = (total + TBX - 1) / TBX;
    const long long cap = (long long)(sm_count > 0 ? sm_count : 148) * 8;
    const unsigned grid = (unsigned)(tiles < cap ? tiles : cap);
    static bool attr_set[2][64] = {};   // per (metric, device): the attribute call costs as much as a launch
    int dev = 0;
    OFPSB_CUDA_TRY(cudaGetDevice(&dev));
    const bool need_attr = dev < 0 || dev >= 64 || !attr_set[p.metric == OFPSB_METRIC_SAD ? 0 : 1][dev];
    if (p.metric == OFPSB_METRIC_SAD) {
        auto k = block_match_list_kernel<B, R, G, TBX, NT, OFPSB_METRIC_SAD>;
        if (need_attr) OFPSB_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
        k<<<grid, NT, C::SMEM_BYTES, stream>>>(mp, mc, p, d_list, d_count);
    } else {
        auto k = block_match_list_kernel<B, R, G, TBX, NT, OFPSB_METRIC_SSD>;
        if (need_attr) OFPSB_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
        k<<<grid, NT, C::SMEM_BYTES, stream>>>(mp, mc, p, d_list, d_count);
    }
    if (dev >= 0 && dev < 64) attr_set[p.metric == OFPSB_METRIC_SAD ? 0 : 1][dev] = true;
    OFPSB_CUDA_TRY(cudaGetLastError());
    return OFPSB_OK;
}

}  // namespace

bool block_match_tma_usable(const BlockMatchParams& p)
{
    const bool aligned = ((reinterpret_cast<uintptr_t>(p.prev) | reinterpret_cast<uintptr_t>(p.cur) | (uintptr_t)p.stride |
                           (uintptr_t)(p.n_pairs > 1 ? p.pair_stride : 0)) & 15) == 0;
    return aligned && p.w <= (1 << 30) && p.stride < (1ll << 32) && get_encode() != nullptr;
}

// Exhaustive search of the blocks in d_list[0 .. *d_count) (global block indices pair*nbx*nby + by*nbx + bx).
// Returns 0 when launched, 1 when no instance applies.
int launch_block_match_list(const BlockMatchParams& p, const uint32_t* d_list, const uint32_t* d_count, int sm_count,
                            cudaStream_t stream)
{
    if (!block_match_tma_usable(p)) return 1;
    // a launch of a few thousand blocks lists a few hundred at most: one block per tile and short dy groups keep the
    // latency of the (single) wave low — the search of a listed block is spread over 198 threads instead of 66
    if (p.block == 16 && p.range == 16 && (long long)p.nbx * p.nby * p.n_pairs <= 40000)
        return launch_list<16, 16, 6, 1, 224>(p, d_list, d_count, sm_count, stream);
    if (p.block == 16 && p.range == 16) return launch_list<16, 16, 17, 4, 288>(p, d_list, d_count, sm_count, stream);
    if (p.block == 16 && p.range == 8) return launch_list<16, 8, 17, 4, 96>(p, d_list, d_count, sm_count, stream);
    if (p.block == 16 && p.range == 32) return launch_list<16, 32, 13, 2, 672>(p, d_list, d_count, sm_count, stream);
    if (p.block == 8 && p.range == 32) return launch_list<8, 32, 22, 2, 416>(p, d_list, d_count, sm_count, stream);
    if (p.block == 8 && p.range == 16) return launch_list<8, 16, 17, 4, 288>(p, d_list, d_count, sm_count, stream);
    if (p.block == 8 && p.range == 8) return launch_list<8, 8, 17, 8, 160>(p, d_list, d_count, sm_count, stream);
    return 1;
}

// Returns OFPSB_OK when launched, 1 when there is no TMA instance for this geometry / the frame
// layout cannot be described by a tensor map (16-byte aligned base and strides), <0 on error.
int launch_block_match_tma(const BlockMatchParams& p, cudaStream_t stream, int variant)
{
    if (!block_match_tma_usable(p)) return 1;
    //                                                       B   R   G  TBX  NT
    if (p.block == 16 && p.range == 16) return variant == 1 ? launch_tma<16, 16, 17, 8, 544>(p, stream)
                                                            : launch_tma<16, 16, 17, 4, 288>(p, stream);
    if (p.block == 16 && p.range == 8) return launch_tma<16, 8, 17, 12, 224>(p, stream);
    if (p.block == 16 && p.range == 32) return launch_tma<16, 32, 13, 4, 672>(p, stream);
    if (p.block == 8 && p.range == 32) return launch_tma<8, 32, 22, 8, 544>(p, stream);
    if (p.block == 8 && p.range == 16) return launch_tma<8, 16, 17, 16, 544>(p, stream);
    if (p.block == 8 && p.range == 8) return launch_tma<8, 8, 17, 16, 288>(p, stream);
    return 1;
}

}  // namespace ofpsb
