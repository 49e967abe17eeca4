// K1, exact pruning front end (successive elimination, Li & Salari 1995) for the SAD block matcher.
//
// An exhaustive 16x16/+-16 SAD search is ~540 abs-diffs per input byte and is bound by the integer
// pipe, ~46x above its HBM time.  The only way towards the memory roofline that keeps
// the exhaustive-search result is to not evaluate candidates that provably cannot win:
//     |sum(cur block) - sum(prev window)|  <=  SAD(cur block, prev window)            (triangle inequality)
// so a candidate whose window-sum bound already exceeds the best exact SAD found so far is out.
//
//   1. window_sum_kernel: S[y][x] = sum of the BxB window of the previous frame at every position
//      (u16, one streaming pass: dp4a row sums, sliding column sums).
//   2. prune_kernel, one warp per block: bound of all (2R+1)^2 candidates from S (one u16 load + one
//      VABSDIFF + one min each), exact SAD at the candidate with the smallest bound key; if that SAD
//      equals its bound the block is decided (nothing can have a smaller key).  Otherwise the
//      candidates whose bound key is still below the best key are counted: a few are evaluated
//      exactly (each can only tighten the best), many send the block to the work list.
//   3. block_match_list_kernel (block_match_tma.cu): exhaustive search of the listed blocks only.
//
// Every step compares the spec's full lexicographic key (cost, dx^2+dy^2, dy, dx), so the output is
// bit-identical to the exhaustive kernels and to the oracle; only the amount of work is data-dependent
// (worst case = bounds + full exhaustive search).
#include "tma_common.cuh"

namespace ofpsb {

namespace {

using namespace bm;
using namespace tma;

constexpr int WS_NT = 128;             // threads per CTA of the window-sum pass
constexpr int WS_COLS = 4 * WS_NT;     // each thread owns 4 adjacent columns

// S[y][x] = sum_{j<B, i<B} prev[y+j][x+i] for 0 <= x <= w-B, 0 <= y <= rows-B (u16: B <= 16).
// One streaming pass: a CTA walks `rs` output rows of a 512-column strip; per input row each thread
// builds 4 horizontal B-sums (dp4a on aligned words, then slide by one byte three times) and updates
// 4 sliding column sums through a B-deep ring in shared memory.
constexpr int WS_RB = 8;                       // input rows per TMA box
constexpr int WS_BOXW = WS_NT + 8;            // u32 elements per box row (>= WS_NT + B/4 + 1, multiple of 4)

// The frame is described to the TMA unit as rows of u32 elements, so one box row can span 544 bytes;
// two stages of WS_RB rows are kept in flight (the load of batch k+2 is issued as soon as batch k has
// been consumed), out-of-frame words arrive as zeros.
template <int B>
__global__ void __launch_bounds__(WS_NT) window_sum_kernel(const __grid_constant__ CUtensorMap map_prev, int w, int rows,
                                                           uint16_t* __restrict__ S, int ws, long long s_plane, int rs)
{
    constexpr int NW = B / 4;
    static_assert(NW + 1 <= 8, "box row too short");
    __shared__ __align__(128) uint32_t rowbuf[2][WS_RB][WS_BOXW];
    __shared__ uint2 ring[B][WS_NT];
    __shared__ __align__(8) uint64_t bars[2];
    const int tid = threadIdx.x;
    const int x_base = blockIdx.x * WS_COLS;
    const int r0 = blockIdx.y * rs;
    uint16_t* out = S + (long long)blockIdx.z * s_plane;
    const int x = x_base + 4 * tid;
    const int r_end = min(r0 + rs + B - 1, rows);
    const int nbatch = (r_end - r0 + WS_RB - 1) / WS_RB;
    constexpr uint32_t TX = WS_RB * WS_BOXW * 4;
    if (tid == 0) {
        for (int i = 0; i < 2; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int kb = 0; kb < 2 && kb < nbatch; kb++) {
            const uint32_t b32 = smem_u32(&bars[kb]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b32), "r"(TX) : "memory");
            tma_load_3d(smem_u32(&rowbuf[kb][0][0]), &map_prev, x_base / 4, r0 + kb * WS_RB, blockIdx.z, b32);
        }
    }
    __syncthreads();
    uint32_t V0 = 0, V1 = 0, V2 = 0, V3 = 0;
    for (int kb = 0; kb < nbatch; kb++) {
        const int st = kb & 1;
        mbar_wait(smem_u32(&bars[st]), (uint32_t)((kb >> 1) & 1));
        const int rb = r0 + kb * WS_RB;
        const int nr = min(WS_RB, r_end - rb);
        for (int q = 0; q < nr; q++) {
            const int r = rb + q;
            const uint32_t* buf = rowbuf[st][q];
            uint32_t wd[NW + 1];
#pragma unroll
            for (int k = 0; k <= NW; k++) wd[k] = buf[tid + k];
            uint32_t H0 = 0;
#pragma unroll
            for (int k = 0; k < NW; k++) H0 = __dp4a(wd[k], 0x01010101u, H0);
            const uint32_t H1 = H0 - (wd[0] & 255u) + (wd[NW] & 255u);
            const uint32_t H2 = H1 - ((wd[0] >> 8) & 255u) + ((wd[NW] >> 8) & 255u);
            const uint32_t H3 = H2 - ((wd[0] >> 16) & 255u) + ((wd[NW] >> 16) & 255u);
            const int slot = (r - r0) & (B - 1);
            uint2 old = make_uint2(0u, 0u);
            if (r - r0 >= B) old = ring[slot][tid];
            V0 += H0 - (old.x & 0xFFFFu);
            V1 += H1 - (old.x >> 16);
            V2 += H2 - (old.y & 0xFFFFu);
            V3 += H3 - (old.y >> 16);
            ring[slot][tid] = make_uint2(H0 | (H1 << 16), H2 | (H3 << 16));
            const int y_out = r - B + 1;
            if (y_out >= r0) {
                uint16_t* o = out + (long long)y_out * ws + x;
                if (x + 3 <= w - B) {
                    *reinterpret_cast<uint2*>(o) = make_uint2((V0 & 0xFFFFu) | (V1 << 16), (V2 & 0xFFFFu) | (V3 << 16));
                } else {
                    if (x <= w - B) o[0] = (uint16_t)V0;
                    if (x + 1 <= w - B) o[1] = (uint16_t)V1;
                    if (x + 2 <= w - B) o[2] = (uint16_t)V2;
                }
            }
        }
        __syncthreads();   // every thread is done with stage st
        if (tid == 0 && kb + 2 < nbatch) {
            const uint32_t b32 = smem_u32(&bars[st]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b32), "r"(TX) : "memory");
            tma_load_3d(smem_u32(&rowbuf[st][0][0]), &map_prev, x_base / 4, r0 + (kb + 2) * WS_RB, blockIdx.z, b32);
        }
    }
}

// Exact SAD of one candidate, one warp: 8 (B=16) or 4 (B=8) bytes of the block per lane, unaligned in x,
// read from the previous-frame window staged in shared memory (pitch `pw` bytes, window coordinates).
template <int B>
__device__ __forceinline__ uint32_t warp_block_sad(const uint8_t* __restrict__ win, int pw, int wx, int wy, uint32_t c0,
                                                   uint32_t c1, int lane)
{
    uint32_t sad = 0;
    if (B == 16) {
        const int row = lane >> 1, xb = wx + 8 * (lane & 1);
        const uint8_t* rp = win + (wy + row) * pw;
        const int xa = xb & ~3, sh = (xb & 3) * 8;
        const uint32_t w0 = *reinterpret_cast<const uint32_t*>(rp + xa);
        const uint32_t w1 = *reinterpret_cast<const uint32_t*>(rp + xa + 4);
        const uint32_t w2 = *reinterpret_cast<const uint32_t*>(rp + xa + 8);   // only its low bytes are used
        sad = sad4_acc(c0, __funnelshift_r(w0, w1, sh), 0);
        sad = sad4_acc(c1, __funnelshift_r(w1, w2, sh), sad);
    } else {   // B == 8: 16 lanes, one word each
        if (lane < 16) {
            const int row = lane >> 1, xb = wx + 4 * (lane & 1);
            const uint8_t* rp = win + (wy + row) * pw;
            const int xa = xb & ~3, sh = (xb & 3) * 8;
            const uint32_t w0 = *reinterpret_cast<const uint32_t*>(rp + xa);
            const uint32_t w1 = *reinterpret_cast<const uint32_t*>(rp + xa + 4);
            sad = sad4_acc(c0, __funnelshift_r(w0, w1, sh), 0);
        }
    }
    return __reduce_add_sync(0xffffffffu, sad);
}

constexpr int PRUNE_WARPS = 8;
constexpr int MAX_SURVIVORS = 8;
constexpr uint32_t S_INVALID = 0x00FFFFFFu;   // stands in for the window sum of an illegal candidate: bound ~2^24
constexpr uint32_t LB_INVALID = 0x00800000u;  // any bound at or above this came from S_INVALID

__device__ __forceinline__ uint32_t pos_code(int dx, int dy, int R)
{
    return ((uint32_t)(dx * dx + dy * dy) << 14) | ((uint32_t)(dy + R) << 7) | (uint32_t)(dx + R);
}

template <int B, int R>
struct PruneCfg {
    static constexpr int ND = 2 * R + 1;
    // S window of a tile of PRUNE_WARPS adjacent blocks: columns x0-R .. x0+(PRUNE_WARPS-1)B+R, rows y0-R .. y0+R
    static constexpr int SW = ((PRUNE_WARPS - 1) * B + ND + 7) & ~7;   // u16 elements per row (TMA: 16-byte multiple)
    static constexpr int S_BYTES = (ND * SW * 2 + 127) & ~127;           // multiples of 128 bytes (TMA destinations)
    static constexpr int RA = (R + 15) & ~15;                             // pixel window origin on a 16-byte boundary (TMA, u8)
    static constexpr int PW = PRUNE_WARPS * B + 2 * RA;                   // previous-frame window of the tile: bytes per row
    static constexpr int PH = B + 2 * R;
    static constexpr int P_BYTES = (PW * PH + 16 + 127) & ~127;           // +16: the last word read may straddle the end
    static constexpr int CW = PRUNE_WARPS * B;                            // current tile: bytes per row
    static constexpr int C_BYTES = (CW * B + 127) & ~127;
    static constexpr uint32_t TX_BYTES = (uint32_t)(ND * SW * 2 + PW * PH + CW * B);
    static_assert(PW % 16 == 0 && PW <= 256 && CW % 16 == 0, "TMA box limits");
    static_assert(SW <= 256 && ND <= 256 && (R % 8) == 0, "TMA box limits / 16-byte aligned window origin");
};

// Window sums of one candidate column group into registers: lanes <-> 32 consecutive dx, one shared-memory
// u16 load per dy at an immediate offset.  `col` = this lane's column in the staged S window.
template <int B, int R>
__device__ __forceinline__ void load_group(const uint16_t* __restrict__ sS, int col, bool interior, bool col_ok, int dy_lo,
                                           int dy_hi, uint32_t (&v)[2 * R + 1])
{
    constexpr int SW = PruneCfg<B, R>::SW;
    if (interior) {   // every candidate of every block of the tile is legal: no predicates
#pragma unroll
        for (int dyi = 0; dyi < 2 * R + 1; dyi++) v[dyi] = sS[dyi * SW + col];
    } else {
#pragma unroll
        for (int dyi = 0; dyi < 2 * R + 1; dyi++) {
            const int dy = dyi - R;
            v[dyi] = (col_ok && dy >= dy_lo && dy <= dy_hi) ? (uint32_t)sS[dyi * SW + col] : S_INVALID;
        }
    }
}

// One block, one warp: bounds from the staged S window, exact SAD of the best-bound candidate, decision.
template <int B, int R>
__device__ __forceinline__ void prune_block(const uint16_t* __restrict__ sS, const uint8_t* __restrict__ sP,
                                            const uint8_t* __restrict__ sC, const BlockMatchParams& p, int bx, int by,
                                            int pair, int x0t, bool tile_full,
                                            int lane, int wib, uint32_t (*s_lb)[MAX_SURVIVORS],
                                            uint32_t (*s_pos)[MAX_SURVIVORS], uint32_t* __restrict__ worklist,
                                            uint32_t* __restrict__ wl_count, unsigned long long* __restrict__ stats)
{
    constexpr int ND = 2 * R + 1;
    // lanes <-> dx for NMAIN groups of 32 columns; a single left-over column (ND % 32 == 1) is walked
    // with lanes <-> dy instead, so no warp iteration runs with one active lane
    constexpr bool HAS_EXTRA = (ND % 32) == 1;
    constexpr int NMAIN = HAS_EXTRA ? ND / 32 : (ND + 31) / 32;
    constexpr int NEXTRA_IT = HAS_EXTRA ? (ND + 31) / 32 : 0;
    constexpr int SW = PruneCfg<B, R>::SW;
    const int x0 = bx * B, y0 = by * B;
    constexpr int PW = PruneCfg<B, R>::PW, CW = PruneCfg<B, R>::CW, XO = PruneCfg<B, R>::RA - R;
    const long long gw = ((long long)pair * p.nby + by) * p.nbx + bx;
    const int dy_lo = max(-R, -p.halo_top - y0), dy_hi = min(R, p.strip_h + p.halo_bottom - B - y0);
    const int dx_lo = max(-R, -x0), dx_hi = min(R, p.w - B - x0);

    // current block -> registers (from the staged tile)
    uint32_t c0 = 0, c1 = 0;
    if (B == 16) {
        const uint2 v = *reinterpret_cast<const uint2*>(sC + (lane >> 1) * CW + wib * B + 8 * (lane & 1));
        c0 = v.x;
        c1 = v.y;
    } else if (lane < 16) {
        c0 = *reinterpret_cast<const uint32_t*>(sC + (lane >> 1) * CW + wib * B + 4 * (lane & 1));
    }
    // the whole tile is interior when no candidate of any of its blocks leaves the frame (CTA-uniform)
    // (with ND < 32 the lanes beyond dx = +R must be masked: only the predicated path does that — ADVICE r1)
    const bool interior = (HAS_EXTRA || ND % 32 == 0) && dy_lo == -R && dy_hi == R && x0t - R >= 0 &&
                          x0t + (PRUNE_WARPS - 1) * B + R <= p.w - B && tile_full;
    const int col0 = wib * B + lane;   // column of dx = lane - R in the staged window

    // ---- pass 1: smallest bound key.  Per lane and column the dy fold uses key = lb << 7 | rank(dy)
    // (rank is an immediate), then the lane keeps its best (lb, position code).
    uint32_t my_lb = 0xFFFFFFFFu, my_pos = 0xFFFFFFFFu;
    uint32_t vex[NEXTRA_IT > 0 ? NEXTRA_IT : 1];
    if (HAS_EXTRA) {   // left-over column dx = +R: lanes <-> dy
#pragma unroll
        for (int t = 0; t < NEXTRA_IT; t++) {
            const int dy = lane + 32 * t - R;
            const bool ok = dy <= R && dy >= dy_lo && dy <= dy_hi && R <= dx_hi;
            vex[t] = ok ? (uint32_t)sS[(dy + R) * SW + wib * B + 2 * R] : S_INVALID;
        }
    }
    uint32_t v[ND];
    load_group<B, R>(sS, col0, interior, lane - R >= dx_lo && lane - R <= dx_hi, dy_lo, dy_hi, v);
    const uint32_t sc = __reduce_add_sync(0xffffffffu, __dp4a(c0, 0x01010101u, __dp4a(c1, 0x01010101u, 0u)));
#pragma unroll
    for (int jg = 0; jg < NMAIN; jg++) {
        const int dx = lane + 32 * jg - R;
        if (jg > 0) load_group<B, R>(sS, col0 + 32 * jg, interior, dx <= R && dx >= dx_lo && dx <= dx_hi, dy_lo, dy_hi, v);
        uint32_t kmin = 0xFFFFFFFFu;
#pragma unroll
        for (int dyi = 0; dyi < ND; dyi++) {
            const int dy = dyi - R;
            kmin = min(kmin, __usad(v[dyi], sc, 0u) * 128u + (uint32_t)(2 * (dy < 0 ? -dy : dy) - (dy < 0 ? 1 : 0)));
        }
        const uint32_t lb = kmin >> 7;
        if (lb < LB_INVALID) {
            const int code = (int)(kmin & 127u);
            const int ady = (code + 1) >> 1;
            const uint32_t pos = pos_code(dx, (code & 1) ? -ady : ady, R);
            if (lb < my_lb || (lb == my_lb && pos < my_pos)) { my_lb = lb; my_pos = pos; }
        }
    }
    if (HAS_EXTRA) {
#pragma unroll
        for (int t = 0; t < NEXTRA_IT; t++) {
            const uint32_t lb = __usad(vex[t], sc, 0u);
            if (lb < LB_INVALID) {
                const uint32_t pos = pos_code(R, lane + 32 * t - R, R);
                if (lb < my_lb || (lb == my_lb && pos < my_pos)) { my_lb = lb; my_pos = pos; }
            }
        }
    }
    const uint32_t lb_min = __reduce_min_sync(0xffffffffu, my_lb);
    const uint32_t pos_min = __reduce_min_sync(0xffffffffu, my_lb == lb_min ? my_pos : 0xFFFFFFFFu);
    if (lb_min == 0xFFFFFFFFu) return;   // no legal candidate (cannot happen: (0,0) is always legal)

    // ---- exact SAD at the candidate with the smallest bound key
    // window origin = (x0t - RA, y0 - R): candidate (dx, dy) of block wib starts at (wib*B + dx + RA, dy + R)
    uint32_t best_cost = warp_block_sad<B>(sP, PW, wib * B + XO + (int)(pos_min & 127u), (int)((pos_min >> 7) & 127u), c0, c1, lane);
    uint32_t best_pos = pos_min;
    unsigned long long evaluated = 1;
    bool resolved = best_cost == lb_min;   // nothing has a smaller (bound, position) key than this exact key

    if (!resolved) {
        // ---- pass 2: candidates whose bound key is still below the best exact key; the first few are recorded
        int survivors = 0;
        const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
        for (int jg = 0; jg < NMAIN; jg++) {
            const int dx = lane + 32 * jg - R;
            if (NMAIN > 1) load_group<B, R>(sS, col0 + 32 * jg, interior, dx <= R && dx >= dx_lo && dx <= dx_hi, dy_lo, dy_hi, v);
#pragma unroll
            for (int dyi = 0; dyi < ND; dyi++) {
                const uint32_t lb = __usad(v[dyi], sc, 0u);
                const uint32_t pos = pos_code(dx, dyi - R, R);
                const bool alive = pos != pos_min && (lb < best_cost || (lb == best_cost && pos < best_pos));
                const unsigned m = __ballot_sync(0xffffffffu, alive);
                if (alive) {
                    const int idx = survivors + __popc(m & lt_mask);
                    if (idx < MAX_SURVIVORS) { s_lb[wib][idx] = lb; s_pos[wib][idx] = pos; }
                }
                survivors += __popc(m);
            }
        }
        if (HAS_EXTRA) {
#pragma unroll
            for (int t = 0; t < NEXTRA_IT; t++) {
                const uint32_t lb = __usad(vex[t], sc, 0u);
                const uint32_t pos = pos_code(R, lane + 32 * t - R, R);
                const bool alive = lb < LB_INVALID && pos != pos_min && (lb < best_cost || (lb == best_cost && pos < best_pos));
                const unsigned m = __ballot_sync(0xffffffffu, alive);
                if (alive) {
                    const int idx = survivors + __popc(m & lt_mask);
                    if (idx < MAX_SURVIVORS) { s_lb[wib][idx] = lb; s_pos[wib][idx] = pos; }
                }
                survivors += __popc(m);
            }
        }
        __syncwarp();
        if (survivors <= MAX_SURVIVORS) {
            // evaluate them exactly; each evaluation can only tighten the best key
            for (int i = 0; i < survivors; i++) {
                const uint32_t clb = s_lb[wib][i], cpos = s_pos[wib][i];
                if (!(clb < best_cost || (clb == best_cost && cpos < best_pos))) continue;   // best tightened meanwhile
                const uint32_t c = warp_block_sad<B>(sP, PW, wib * B + XO + (int)(cpos & 127u), (int)((cpos >> 7) & 127u), c0, c1, lane);
                evaluated++;
                if (c < best_cost || (c == best_cost && cpos < best_pos)) { best_cost = c; best_pos = cpos; }
            }
            resolved = true;
        }
    }
    if (lane == 0) {
        if (resolved) {
            write_block_outputs(p, (size_t)gw, ((unsigned long long)best_cost << 27) | best_pos, bx, by);
        } else {
            worklist[atomicAdd(wl_count, 1u)] = (uint32_t)gw;
        }
        if (stats) {
            atomicAdd(&stats[0], 1ull);
            atomicAdd(&stats[1], resolved ? 1ull : 0ull);
            atomicAdd(&stats[2], evaluated);
        }
    }
}

// One CTA = PRUNE_WARPS adjacent blocks.  Everything the tile needs — the S window, the previous-frame pixel
// window and the current tile — arrives by three TMA box loads issued up front; the body then runs out of
// shared memory only (no dependent global loads), and resident CTAs overlap each other's load latency.
// (A persistent variant walking 16 block rows per CTA with two prefetched stages measured 1.7x SLOWER on
// B200 — 671 vs 399 us for 64 1080p pairs — so tiles stay one per CTA.)
template <int B, int R>
__global__ void __launch_bounds__(PRUNE_WARPS * 32) prune_kernel(const __grid_constant__ CUtensorMap map_S,
                                                                 const __grid_constant__ CUtensorMap map_prev,
                                                                 const __grid_constant__ CUtensorMap map_cur,
                                                                 const BlockMatchParams p,
                                                                 uint32_t* __restrict__ worklist,
                                                                 uint32_t* __restrict__ wl_count,
                                                                 unsigned long long* __restrict__ stats)
{
    using C = PruneCfg<B, R>;
    __shared__ __align__(128) uint8_t smem[C::S_BYTES + C::P_BYTES + C::C_BYTES];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_lb[PRUNE_WARPS][MAX_SURVIVORS], s_pos[PRUNE_WARPS][MAX_SURVIVORS];
    const uint16_t* sS = reinterpret_cast<const uint16_t*>(smem);
    const uint8_t* sP = smem + C::S_BYTES;
    const uint8_t* sC = sP + C::P_BYTES;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int bx = blockIdx.x * PRUNE_WARPS + wib, by = blockIdx.y, pair = blockIdx.z;
    const int x0t = blockIdx.x * PRUNE_WARPS * B, y0 = by * B;
    const bool tile_full = (int)(blockIdx.x + 1) * PRUNE_WARPS <= p.nbx;
    if (threadIdx.x == 0) {
        const uint32_t b32 = smem_u32(&bar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b32));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b32), "r"(C::TX_BYTES) : "memory");
        // row 0 of the S plane / of the previous-frame tensor is the first halo row; out-of-range = zeros (masked)
        tma_load_3d(smem_u32(sS), &map_S, x0t - R, y0 + p.halo_top - R, pair, b32);
        tma_load_3d(smem_u32(sP), &map_prev, x0t - C::RA, y0 + p.halo_top - R, pair, b32);
        tma_load_3d(smem_u32(sC), &map_cur, x0t, y0, pair, b32);
    }
    __syncthreads();
    if (bx >= p.nbx) return;
    mbar_wait(smem_u32(&bar), 0);
    prune_block<B, R>(sS, sP, sC, p, bx, by, pair, x0t, tile_full, lane, wib, s_lb, s_pos, worklist, wl_count, stats);
}

}  // namespace

bool block_match_tma_usable(const BlockMatchParams& p);
int launch_block_match_list(const BlockMatchParams& p, const uint32_t* d_list, const uint32_t* d_count, int sm_count,
                            cudaStream_t stream);

// Exact pruned SAD search.  Returns 0 when launched, 1 when the path does not apply (metric, geometry,
// alignment) — the caller then runs the exhaustive kernels — and < 0 on error.
static int pruned_chunk(const BlockMatchParams& p, BlockMatchScratch& sc, int sm_count, cudaStream_t stream,
                        uint64_t* launches, bool first_chunk)
{
    if (p.metric != OFPSB_METRIC_SAD || (p.w & 3) || !block_match_tma_usable(p)) return 1;
    const bool geom = (p.block == 16 && (p.range == 8 || p.range == 16 || p.range == 32)) ||
                      (p.block == 8 && (p.range == 8 || p.range == 16 || p.range == 32));
    if (!geom) return 1;
    const int B = p.block;
    const int rows_prev = p.halo_top + p.strip_h + p.halo_bottom;
    if (p.w < B || rows_prev < B) return 1;
    const long long total = (long long)p.nbx * p.nby * p.n_pairs;
    if (total >= 0xFFFFFFF0ll) return 1;
    const int ws = (p.w + 7) & ~7;
    const long long s_plane = (long long)ws * rows_prev;
    if (int rc = sc.sums.reserve((size_t)s_plane * p.n_pairs * sizeof(uint16_t))) return rc;
    if (int rc = sc.worklist.reserve((size_t)(total + 8) * sizeof(uint32_t) + 64)) return rc;
    uint32_t* wl_count = sc.worklist.as<uint32_t>();
    unsigned long long* stats = reinterpret_cast<unsigned long long*>(wl_count + 2);
    uint32_t* worklist = wl_count + 16;
    OFPSB_CUDA_TRY(cudaMemsetAsync(wl_count, 0, first_chunk ? 64 : 8, stream));   // counters accumulate over chunks

    const uint8_t* prev_base = p.prev - (long long)p.halo_top * p.stride;
    // rows per CTA of the window-sum pass: enough CTAs to fill the machine, little redundant warm-up
    const int col_ctas = (p.w - B + 1 + WS_COLS - 1) / WS_COLS;
    const int out_rows = rows_prev - B + 1;
    int rs = 128;
    while (rs > 16 && (long long)col_ctas * ((out_rows + rs - 1) / rs) * p.n_pairs < 8ll * (sm_count > 0 ? sm_count : 148)) rs >>= 1;
    dim3 g1(col_ctas, (out_rows + rs - 1) / rs, p.n_pairs);
    CUtensorMap map32;
    if (!make_map_elems(&map32, 4, prev_base, p.w / 4, rows_prev, p.stride, p.pair_stride, p.n_pairs, WS_BOXW, WS_RB)) return 1;
    if (B == 16)
        window_sum_kernel<16><<<g1, WS_NT, 0, stream>>>(map32, p.w, rows_prev, sc.sums.as<uint16_t>(), ws, s_plane, rs);
    else
        window_sum_kernel<8><<<g1, WS_NT, 0, stream>>>(map32, p.w, rows_prev, sc.sums.as<uint16_t>(), ws, s_plane, rs);
    const dim3 g2((p.nbx + PRUNE_WARPS - 1) / PRUNE_WARPS, p.nby, p.n_pairs);
#define OFPSB_PRUNE(BB, RR)                                                                                          \
    do {                                                                                                             \
        using PC = PruneCfg<BB, RR>;                                                                                 \
        CUtensorMap map_s, map_p, map_c;                                                                             \
        if (!make_map_elems(&map_s, 2, sc.sums.ptr, p.w - B + 1, rows_prev - B + 1, (long long)ws * 2, s_plane * 2,  \
                            p.n_pairs, PC::SW, PC::ND) ||                                                            \
            !make_map(&map_p, prev_base, p.w, rows_prev, p.stride, p.pair_stride, p.n_pairs, PC::PW, PC::PH) ||      \
            !make_map(&map_c, p.cur, p.w, p.strip_h, p.stride, p.pair_stride, p.n_pairs, PC::CW, BB))                \
            return 1;                                                                                                \
        prune_kernel<BB, RR><<<g2, PRUNE_WARPS * 32, 0, stream>>>(map_s, map_p, map_c, p, worklist, wl_count,        \
                                                                  sc.collect_stats ? stats : nullptr);              \
    } while (0)
    if (B == 16 && p.range == 16) OFPSB_PRUNE(16, 16);
    else if (B == 16 && p.range == 8) OFPSB_PRUNE(16, 8);
    else if (B == 16 && p.range == 32) OFPSB_PRUNE(16, 32);
    else if (B == 8 && p.range == 32) OFPSB_PRUNE(8, 32);
    else if (B == 8 && p.range == 16) OFPSB_PRUNE(8, 16);
    else OFPSB_PRUNE(8, 8);
#undef OFPSB_PRUNE
    OFPSB_CUDA_TRY(cudaGetLastError());
    const int rc = launch_block_match_list(p, worklist, wl_count, sm_count, stream);
    if (rc != OFPSB_OK) {
        set_error("block_match: work-list kernel unavailable for block=%d range=%d", p.block, p.range);
        return rc < 0 ? rc : OFPSB_E_INVALID;
    }
    if (launches) *launches += 3;
    return OFPSB_OK;
}

// Public entry.  ncu shows 905 MB of DRAM traffic per 64-pair step for 274 MB of algorithmic bytes (the 262 MB
// of window sums round-trip through HBM, the frames are read by all three kernels), so the batch CAN be walked
// in chunks small enough to stay L2-resident ("block_match_chunk_pairs").  Measured on B200 (64 1080p pairs):
// chunks of 4 / 8 / 16 / 32 / 64 pairs -> 20.1 / 17.1 / 15.7 / 14.6 / 14.3 us per pair: the path is bound by
// instruction issue and TMA latency, not by HBM, and smaller launches only add tails — so the default is one chunk.
int launch_block_match_pruned(const BlockMatchParams& p, BlockMatchScratch& sc, int sm_count, cudaStream_t stream,
                              uint64_t* launches)
{
    // the fused four-term SEA kernel covers every tuned geometry; what it declines (or skips on content feedback) goes
    // to the exhaustive kernels.  The round-1 pipeline below runs only on request ("block_match_pruner" = 1): at
    // +-32 it is slower than the exhaustive search on all but noise-free 16x16 content.
    if (sc.pruner != 1) return launch_block_match_sea(p, sc, sm_count, stream, launches);
    int chunk = sc.chunk_pairs;
    if (chunk <= 0) chunk = p.n_pairs;
    if (chunk >= p.n_pairs) return pruned_chunk(p, sc, sm_count, stream, launches, true);
    const size_t nb = (size_t)p.nbx * p.nby;
    for (int first = 0; first < p.n_pairs; first += chunk) {
        BlockMatchParams q = p;
        q.n_pairs = p.n_pairs - first < chunk ? p.n_pairs - first : chunk;
        q.prev = p.prev + (long long)first * p.pair_stride;
        q.cur = p.cur + (long long)first * p.pair_stride;
        if (p.mv_xy) q.mv_xy = p.mv_xy + 2 * nb * first;
        if (p.cost) q.cost = p.cost + nb * first;
        if (p.entries) q.entries = p.entries + nb * first;
        const int rc = pruned_chunk(q, sc, sm_count, stream, launches, first == 0);
        if (rc) return rc;
    }
    return OFPSB_OK;
}

}  // namespace ofpsb
