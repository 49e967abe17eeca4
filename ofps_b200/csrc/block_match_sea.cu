// K1, default SAD path: fused multi-level successive elimination (SEA / MSEA, Li & Salari 1995, Gao et al. 2000).
//
// The exhaustive 16x16/+-16 SAD search is ~540 abs-diffs per input byte and is bound by the integer pipe, ~46x above
// its HBM time.  The way towards the memory roofline that keeps the exhaustive-search RESULT is to not
// evaluate candidates that provably cannot win.  With the block split into four sub-blocks of edge N = B/2,
//     sum_k | sum(cur sub-block k) - sum(prev sub-window k) |   <=   SAD(cur block, prev window)
// (triangle inequality per sub-block), so a candidate whose bound is not below the best exact cost found so far is out.
// The four-term bound is what makes this work on real (noisy) content: on the noisy 1080p stream the 16x16 one-term
// bound of round 1 left 237 of 1089 candidates per block alive, the four-term bound 4 (tools/sea_survivors.py).
//
// One CTA = one 128 x 64 pixel tile of blocks, one launch for the whole batch; nothing but the frames is read from
// global memory and nothing but the results is written (round 1 wrote a u16 window-sum plane per frame to HBM and read it
// back: 3.3x the algorithmic traffic):
//   1. two TMA box loads: the previous-frame window of the tile (tile + R on every side; out-of-frame bytes arrive as
//      zeros) and the current tile;
//   2. the N x N window sums of every window position, in shared memory: one horizontal pass (dp4a on the staged words,
//      four positions per thread) and one vertical sliding pass, in place (u16 pairs, no carries between the halves);
//   3. one warp per block, a warp walks down a column of blocks:
//        a. exact SAD of the zero vector and of a predictor (the winner of the block above; for the first block of a
//           warp the candidate with the smallest bound of the tile's first block, found by the whole CTA);
//        b. best cost 0: only a candidate with a SHORTER vector and bound 0 could still win -> scan the rows
//           |dy| <= sqrt(d2 of the best) for window sums equal to the block's (one load + one compare per candidate);
//        c. otherwise: the bounds of all (2R+1)^2 candidates into registers (lanes <-> dx: 2.5 shared loads + 4
//           VABSDIFF per candidate), the candidate with the smallest bound is evaluated exactly, the candidates whose
//           bound is still below the best are collected (bit masks, no ballots) and evaluated; more than SEA_CAP of
//           them send the block to the work list of the exhaustive kernel (block_match_tma.cu).
// Every comparison uses the spec's lexicographic key (cost, dx^2+dy^2, dy, dx), so the output is bit-identical to the
// exhaustive kernels and to the oracle; only the amount of work is data-dependent.  Predictors change the work, never
// the result: a candidate is dropped only when its (bound, position) key is not below an exact key already found.
#ifndef OFPSB_EMU
#include <cstdlib>
#include "tma_common.cuh"
#else
#include "block_match_common.cuh"
#endif

namespace ofpsb {

namespace {

using namespace bm;
#ifndef OFPSB_EMU
using namespace tma;
#endif

constexpr int SEA_WARPS = 8;
constexpr int SEA_NT = SEA_WARPS * 32;
constexpr int SEA_CAP = 32;                   // survivors a warp evaluates itself; more -> exhaustive work list
constexpr uint32_t SEA_BIG = 0x00FFFFFFu;     // bound of an illegal candidate (real bounds are < 2^16)
constexpr int SEA_TILE_W = 128;   // tile height TH is a template parameter: 64 rows, or 32 for launches that do not fill the machine

template <int B, int R, int TH>
struct SeaCfg {
    static constexpr int N = B / 2;                       // sub-block edge
    static constexpr int ND = 2 * R + 1;
    static constexpr int TBX = SEA_TILE_W / B, TBY = TH / B;
    static constexpr int RA = (R + 15) & ~15;             // window origin on a 16-byte boundary (TMA, u8)
    static constexpr int PW = SEA_TILE_W + 2 * RA;        // previous-frame window: bytes per row
    static constexpr int PH = TH + 2 * R;
    static constexpr int CW = SEA_TILE_W, CH = TH;
    // pitch of the window-sum plane in u16: 84 words per row instead of 80, so that a COLUMN walk (lanes <-> dy: the best's
    // own column in the zero-cost scan, the dx = +R column of the full scan) hits 8 different banks instead of 2
    // (ncu r2: 44 % of the kernel's shared-load wavefronts were bank-conflict replays and the LSU pipe was its co-limiter)
    static constexpr int SP = PW + 8;
    static constexpr int WC = SP / 2;                     // window-sum plane: u32 (= two u16 sums) per row
    static constexpr int OR = PH - N + 1;                 // rows of window sums
    static constexpr int NSEG = SEA_NT / WC;              // row segments of the vertical pass
    static constexpr int SR = (OR + NSEG - 1) / NSEG;     // output rows per segment
    static constexpr int HR = (NSEG * SR + N - 1) > PH ? (NSEG * SR + N - 1) : PH;   // plane rows incl. slack
    static constexpr int P_BYTES = (PW * PH + 16 + 127) & ~127;   // +16: the last words read straddle the end
    static constexpr int S_BYTES = (SP * 2 * HR + 127) & ~127;
    static constexpr int C_BYTES = (CW * CH + 127) & ~127;
    static constexpr int SMEM_BYTES = P_BYTES + S_BYTES + C_BYTES + 128;   // + alignment slack
    static constexpr uint32_t TX_BYTES = (uint32_t)(PW * PH + CW * CH);
    static constexpr bool EXTRA = ND > 32;                // column dx = +R is walked with lanes <-> dy
    static constexpr int NL = EXTRA ? 32 : ND;            // lanes of the lanes <-> dx mapping (dx = lane - R)
    static constexpr int NEX = EXTRA ? (ND + 31) / 32 : 0;
    static constexpr int CPW = TBX / SEA_WARPS;           // block columns per warp
    static_assert(B == 8 || B == 16, "sub-block sums are staged for 8x8 and 16x16 blocks");
    static_assert(ND <= 65, "lanes <-> dx in groups of 32 columns plus the dx = +R column; 7-bit position fields");
    static_assert(TBX % SEA_WARPS == 0 && PW % 16 == 0 && PW <= 256 && PH <= 256, "tile / TMA box limits");
    static_assert(SR >= N && NSEG >= 1, "vertical pass: a segment is at least one window high");
    static_assert(RA + B - R >= 2 * N, "every window-sum column a candidate reads is a full window");
};

__device__ __forceinline__ uint32_t sea_pos(int dx, int dy, int R)
{
    return ((uint32_t)(dx * dx + dy * dy) << 14) | ((uint32_t)(dy + R) << 7) | (uint32_t)(dx + R);
}

// ---- step 2: N x N window sums of the staged window, in shared memory ------------------------------------------------
// horizontal: H[y][x] = sum of N bytes of row y starting at x, four x per thread from two / three aligned words
template <int N, int PW, int PH, int SP>
__device__ __forceinline__ void sea_hpass(const uint8_t* __restrict__ sP, uint32_t* __restrict__ sS, int tid)
{
    // one item = 16 consecutive positions of one row: one 16-byte load + the one / two words behind it
    constexpr int IPR = PW / 16;
    for (int item = tid; item < PH * IPR; item += SEA_NT) {
        const int row = item / IPR, k = item - row * IPR;
        const uint8_t* src = sP + row * PW + 16 * k;
        const uint4 a = *reinterpret_cast<const uint4*>(src);
        uint32_t w[6] = {a.x, a.y, a.z, a.w, 0u, 0u};
        if (N == 8) {
            const uint2 t = *reinterpret_cast<const uint2*>(src + 16);   // last item of the last row: the +16 slack
            w[4] = t.x;
            w[5] = t.y;
        } else {
            w[4] = *reinterpret_cast<const uint32_t*>(src + 16);
        }
        uint32_t m[5];
#pragma unroll
        for (int j = 0; j < 5; j++) m[j] = __dp4a(w[j], 0x01010101u, 0u);
        uint32_t o[8];
#pragma unroll
        for (int g = 0; g < 4; g++) {
            uint32_t h0, h1, h2, h3;
            if (N == 8) {
                h0 = m[g] + m[g + 1];
                h1 = __dp4a(w[g], 0x01010100u, __dp4a(w[g + 2], 0x00000001u, m[g + 1]));
                h2 = __dp4a(w[g], 0x01010000u, __dp4a(w[g + 2], 0x00000101u, m[g + 1]));
                h3 = __dp4a(w[g], 0x01000000u, __dp4a(w[g + 2], 0x00010101u, m[g + 1]));
            } else {
                h0 = m[g];
                h1 = __dp4a(w[g], 0x01010100u, __dp4a(w[g + 1], 0x00000001u, 0u));
                h2 = __dp4a(w[g], 0x01010000u, __dp4a(w[g + 1], 0x00000101u, 0u));
                h3 = __dp4a(w[g], 0x01000000u, __dp4a(w[g + 1], 0x00010101u, 0u));
            }
            o[2 * g] = h0 | (h1 << 16);
            o[2 * g + 1] = h2 | (h3 << 16);
        }
        uint4* dst = reinterpret_cast<uint4*>(sS + row * (SP / 2) + 8 * k);
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
    }
}

// vertical, in place: S[y] = H[y] + ... + H[y+N-1] on u16 pairs (sums < 2^16: no carry or borrow crosses the halves).
// A thread owns one u32 column of one row segment; the N-1 rows below its segment are read before anyone writes.
template <typename C>
__device__ __forceinline__ void sea_vpass(uint32_t* __restrict__ sS, int tid)
{
    constexpr int N = C::N, WC = C::WC, SR = C::SR;
    const bool active = tid < C::NSEG * WC;
    const int seg = tid / WC, col = tid - seg * WC;
    uint32_t* base = sS + (seg * SR) * WC + col;
    uint32_t tail[N - 1];
    if (active) {
#pragma unroll
        for (int j = 0; j < N - 1; j++) tail[j] = base[(SR + j) * WC];
    }
    __syncthreads();
    if (active) {
        uint32_t ring[N], v = 0;
#pragma unroll
        for (int j = 0; j < N; j++) { ring[j] = base[j * WC]; v += ring[j]; }
#pragma unroll
        for (int y = 0; y < SR; y++) {
            base[y * WC] = v;
            const uint32_t next = (y + N < SR) ? base[(y + N) * WC] : tail[(y + N - SR) < (N - 1) ? (y + N - SR) : 0];
            v += next - ring[y % N];
            ring[y % N] = next;
        }
    }
}

// Exact SAD of one candidate, one warp: 8 (B=16) or 4 (B=8) bytes of the block per lane, unaligned in x,
// read from the staged previous-frame window (window coordinates wx, wy).
template <int B, int PW>
__device__ __forceinline__ uint32_t sea_exact(const uint8_t* __restrict__ win, int wx, int wy, uint32_t c0, uint32_t c1, int lane)
{
    uint32_t sad = 0;
    if (B == 16) {
        const int row = lane >> 1, xb = wx + 8 * (lane & 1);
        const uint8_t* rp = win + (wy + row) * PW;
        const int xa = xb & ~3, sh = (xb & 3) * 8;
        const uint32_t w0 = *reinterpret_cast<const uint32_t*>(rp + xa);
        const uint32_t w1 = *reinterpret_cast<const uint32_t*>(rp + xa + 4);
        const uint32_t w2 = *reinterpret_cast<const uint32_t*>(rp + xa + 8);   // only its low bytes are used
        sad = sad4_acc(c0, __funnelshift_r(w0, w1, sh), 0);
        sad = sad4_acc(c1, __funnelshift_r(w1, w2, sh), sad);
    } else if (lane < 16) {   // B == 8: 16 lanes, one word each
        const int row = lane >> 1, xb = wx + 4 * (lane & 1);
        const uint8_t* rp = win + (wy + row) * PW;
        const int xa = xb & ~3, sh = (xb & 3) * 8;
        const uint32_t w0 = *reinterpret_cast<const uint32_t*>(rp + xa);
        const uint32_t w1 = *reinterpret_cast<const uint32_t*>(rp + xa + 4);
        sad = sad4_acc(c0, __funnelshift_r(w0, w1, sh), 0);
    }
    return __reduce_add_sync(0xffffffffu, sad);
}

// K2: block winner -> outputs (av-decoder/src/lib.rs:404-419 convention; same operations as write_block_outputs,
// with frame_norm = (1/W, 1/H) computed once per thread instead of once per block)
template <int R>
__device__ __forceinline__ void sea_write(const BlockMatchParams& p, size_t gb, uint32_t cost, uint32_t pos, int bx, int by,
                                          float nx, float ny)
{
    const int dx = (int)(pos & 127u) - R, dy = (int)((pos >> 7) & 127u) - R;
    if (p.mv_xy) *reinterpret_cast<uint32_t*>(p.mv_xy + 2 * gb) = ((uint32_t)dx & 0xFFFFu) | ((uint32_t)dy << 16);
    if (p.cost) p.cost[gb] = cost;
    if (p.entries) {
        const int src_x = bx * p.block + p.block / 2 + dx;
        const int src_y = p.y_offset + by * p.block + p.block / 2 + dy;
        const float4 e = make_float4(__fmul_rn((float)src_x, nx), __fmul_rn((float)src_y, ny), __fmul_rn((float)dx, -nx),
                                     __fmul_rn((float)dy, -ny));
        *reinterpret_cast<float4*>(p.entries + gb) = e;
    }
}

struct SeaOut {
    uint32_t* worklist;
    uint32_t* wl_count;
    unsigned long long* stats;   // [0] blocks, [1] resolved here, [2] exact evaluations, [3] blocks that ran the full scan
    float nx, ny;                // frame_norm = (1/W, 1/H), IEEE f32 divisions done once on the host
    int prefetch_tiles;          // L2 prefetch distance in tiles (0 = off): about the number of resident CTAs
    // peer-halo mode (spatial tiling over GPUs): the previous-frame tensor holds this rank's own rows only; the halo
    // rows above / below are read straight from the neighbours' HBM (tensor maps on peer-mapped memory)
    int peer, own_rows, up_rows, has_up, has_down;
    // profiling hook (env OFPSB_DEBUG_SEA_STOP, tools/debug_sea_stop.py): 1 = return after the loads, 2 = after the window
    // sums, 3 = after the tile predictor.  Measured on B200, us per 1080p pair: 1.62 / 1.88 / 2.26 of the kernel's 5.36 —
    // the loads are bound by the TMA unit's row rate (u8 and u32 tensor descriptions take the same time; a direct copy by
    // all threads with cp.async is faster alone, 1.25, but slows the whole kernel, 5.72: it takes issue slots from the CTAs
    // that are computing, the TMA box does not)
    int debug_stop;
};

// The current tile is staged as [16-byte column block][row][16 bytes] (transposing tensor map, tma_common.cuh): the rows of
// one block are contiguous, so the loads below are conflict-free.  (Row-major staging put the 16 rows of a block on the
// same four banks: a 16-way conflict on every block's first load.)
// current block -> registers: 8 (B=16) / 4 (B=8, lanes 0..15) bytes per lane, lane = 2 * row + half
template <int B, int CH>
__device__ __forceinline__ void sea_cur_block(const uint8_t* __restrict__ sC, int bxl, int byl, int lane, uint32_t& c0,
                                              uint32_t& c1)
{
    c0 = c1 = 0;
    if (B == 16) {
        const uint2 v = *reinterpret_cast<const uint2*>(sC + (bxl * CH + byl * 16) * 16 + 8 * lane);
        c0 = v.x;
        c1 = v.y;
    } else if (lane < 16) {
        c0 = *reinterpret_cast<const uint32_t*>(sC + ((bxl >> 1) * CH + byl * 8 + (lane >> 1)) * 16 + (bxl & 1) * 8 + 4 * (lane & 1));
    }
}

// the four N x N sub-block sums of every block of the current tile -> s_csum[block] = (C00, C10, C01, C11).
// One thread per 16-byte row chunk (consecutive lanes = consecutive rows = consecutive chunks), rows folded by shuffles.
template <typename C, int B>
__device__ __forceinline__ void sea_cur_sums(const uint8_t* __restrict__ sC, uint32_t* __restrict__ s_csum, int tid)
{
    constexpr int CH = C::CH, NCB = C::CW / 16;
    static_assert((NCB * CH) % 32 == 0 && CH % 16 == 0, "whole warps per pass, whole blocks per column");
    for (int i = tid; i < NCB * CH; i += SEA_NT) {
        const int cb = i / CH, row = i - cb * CH;
        const uint4 w = *reinterpret_cast<const uint4*>(sC + (size_t)i * 16);
        if (B == 16) {   // chunk = one row of block (cb, row / 16): left / right 8 bytes
            uint32_t l = __dp4a(w.x, 0x01010101u, __dp4a(w.y, 0x01010101u, 0u));
            uint32_t r = __dp4a(w.z, 0x01010101u, __dp4a(w.w, 0x01010101u, 0u));
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
                l += __shfl_xor_sync(0xffffffffu, l, o);
                r += __shfl_xor_sync(0xffffffffu, r, o);
            }
            if ((row & 7) == 0) {
                const int blk = (row >> 4) * C::TBX + cb, j = (row >> 3) & 1;
                s_csum[4 * blk + 2 * j] = l;
                s_csum[4 * blk + 2 * j + 1] = r;
            }
        } else {         // chunk = one row of blocks (2 cb, row / 8) and (2 cb + 1, row / 8): four 4-byte pieces
            uint32_t s0 = __dp4a(w.x, 0x01010101u, 0u), s1 = __dp4a(w.y, 0x01010101u, 0u);
            uint32_t s2 = __dp4a(w.z, 0x01010101u, 0u), s3 = __dp4a(w.w, 0x01010101u, 0u);
#pragma unroll
            for (int o = 1; o < 4; o <<= 1) {
                s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                s3 += __shfl_xor_sync(0xffffffffu, s3, o);
            }
            if ((row & 3) == 0) {
                const int blk = (row >> 3) * C::TBX + 2 * cb, j = (row >> 2) & 1;
                s_csum[4 * blk + 2 * j] = s0;
                s_csum[4 * blk + 2 * j + 1] = s1;
                s_csum[4 * (blk + 1) + 2 * j] = s2;
                s_csum[4 * (blk + 1) + 2 * j + 1] = s3;
            }
        }
    }
}

struct SeaResult {   // per-block result, kept by lane `it` of the warp until the tile is written out
    uint32_t cost, pos;
    bool resolved;
};

// Survivors of a full scan (s_list[0 .. total), total <= 32, entries dyi << 7 | dxi): one entry per lane — its bound and
// position stay in registers — and the warp evaluates them in ascending bound order, re-testing every survivor against
// the tightening best: entries that no longer qualify cost nothing.  skip*: positions already evaluated.
template <int B, int R, int N, int PW, int SP>
__device__ __forceinline__ void sea_eval_list(const uint32_t* __restrict__ s_list, int total, const uint16_t* __restrict__ s_blk,
                                              const uint8_t* __restrict__ sP, int wx0, int wy0, uint32_t c0, uint32_t c1, int lane,
                                              uint32_t C00, uint32_t C10, uint32_t C01, uint32_t C11, uint32_t skip0, uint32_t skip1,
                                              uint32_t skip2, uint32_t& bc, uint32_t& bp, unsigned long long& evaluated)
{
    constexpr uint32_t NONE = 0xFFFFFFFFu;
    uint32_t lb = NONE, pos = NONE;
    if (lane < total) {
        const uint32_t e = s_list[lane];
        const int dxi = (int)(e & 127u), dyi = (int)(e >> 7);
        pos = sea_pos(dxi - R, dyi - R, R);
        const uint16_t* q = s_blk + dyi * SP + dxi;   // s_blk: window sum of the block at (dx, dy) = (-R, -R)
        lb = __usad(q[N * SP + N], C11, __usad(q[N * SP], C01, __usad(q[N], C10, __usad(q[0], C00, 0u))));
        if (pos == skip0 || pos == skip1 || pos == skip2) lb = NONE;
    }
    for (;;) {
        const bool cand = lb < bc || (lb == bc && pos < bp);
        const uint32_t k = __reduce_min_sync(0xffffffffu, cand ? lb : NONE);
        if (k == NONE) break;
        const int l = __ffs(__ballot_sync(0xffffffffu, cand && lb == k)) - 1;
        const uint32_t cp = __shfl_sync(0xffffffffu, pos, l);
        const uint32_t c = sea_exact<B, PW>(sP, wx0 + (int)(cp & 127u) - R, wy0 + (int)((cp >> 7) & 127u) - R, c0, c1, lane);
        evaluated++;
        if (c < bc || (c == bc && cp < bp)) { bc = c; bp = cp; }
        if (lane == l) lb = NONE;
    }
}

// One chunk of RS rows of a +-32 full scan (sea_block_wide): row j of the chunk finishes the bound of dy index dyi0 + j
// (started N rows earlier, ring slot j) and starts the bound whose upper sub-blocks it holds (slot (j + N) % RS).
// kc: smallest bound << 7 | j of the chunk; code: bit j set when the bound is not above `thr`.  CHECK: the chunk
// holds rows outside the legal dy range (j - lo_rel > span, unsigned).
template <int N, int SP, int RS, bool CHECK>
__device__ __forceinline__ void sea_scan_chunk(const uint16_t* __restrict__ q, uint32_t (&acc)[RS], uint32_t C00, uint32_t C10,
                                               uint32_t C01, uint32_t C11, uint32_t thr, int lo_rel, uint32_t span, uint32_t& kc,
                                               uint32_t& code)
{
#pragma unroll
    for (int j = 0; j < RS; j++) {
        const uint32_t sa = q[j * SP], sb = q[j * SP + N];
        uint32_t b = __usad(sb, C11, __usad(sa, C01, acc[j]));
        if (CHECK && (uint32_t)(j - lo_rel) > span) b = SEA_BIG;
        kc = min(kc, b * 128u + (uint32_t)j);
        code |= (b <= thr ? 1u : 0u) << j;
        acc[(j + N) % RS] = __usad(sb, C10, __usad(sa, C00, 0u));
    }
}

// ---- step 3: one block, one warp ------------------------------------------------------------------------------------
// sS: window sums (u16, pitch PW), sP: previous-frame window, sC: current tile.  (bxl, byl): block inside the tile.
// pred: position code of the predictor.  Returns the winner (or, unresolved, the best found so far).
template <int B, int R, int TH>
__device__ __forceinline__ SeaResult sea_block(const uint16_t* __restrict__ sS, const uint8_t* __restrict__ sP,
                                               const uint8_t* __restrict__ sC, const uint32_t* __restrict__ s_csum,
                                               const BlockMatchParams& p, int bx, int by, int bxl, int byl, bool interior,
                                               uint32_t pred, int lane, uint32_t* __restrict__ s_list, const SeaOut& out)
{
    using C = SeaCfg<B, R, TH>;
    constexpr int N = C::N, ND = C::ND, PW = C::PW, SP = C::SP;
    constexpr uint32_t NONE = 0xFFFFFFFFu;
    int dy_lo = -R, dy_hi = R, dx_lo = -R, dx_hi = R;        // interior tiles (warp-uniform): every candidate is legal
    if (!interior) {
        const int x0 = bx * B, y0 = by * B;
        dy_lo = max(-R, -p.halo_top - y0);
        dy_hi = min(R, p.strip_h + p.halo_bottom - B - y0);
        dx_lo = max(-R, -x0);
        dx_hi = min(R, p.w - B - x0);
    }
    const int wx0 = bxl * B + C::RA, wy0 = byl * B + R;      // block origin in window coordinates
    const int dx = lane - R;                                  // lanes <-> dx mapping
    const bool lane_in = lane < C::NL && dx >= dx_lo && dx <= dx_hi;

    uint32_t c0, c1;
    sea_cur_block<B, C::CH>(sC, bxl, byl, lane, c0, c1);
    const uint4 csum = *reinterpret_cast<const uint4*>(s_csum + 4 * (byl * C::TBX + bxl));
    const uint32_t C00 = csum.x, C10 = csum.y, C01 = csum.z, C11 = csum.w;

    // ---- a. exact cost of the predictor; of the zero vector too unless the predictor already matches exactly (the
    // zero vector is then one of the shorter candidates step b rules out by their window sums)
    const uint32_t pos00 = sea_pos(0, 0, R);
    int pdx = (int)(pred & 127u) - R, pdy = (int)((pred >> 7) & 127u) - R;
    if (!interior && (pdx < dx_lo || pdx > dx_hi || pdy < dy_lo || pdy > dy_hi)) {
        pred = pos00;
        pdx = pdy = 0;
    }
    uint32_t bc = sea_exact<B, PW>(sP, wx0 + pdx, wy0 + pdy, c0, c1, lane);
    uint32_t bp = pred;
    unsigned long long evaluated = 1;
    uint32_t zpos = pred == pos00 ? pos00 : NONE;             // pos00 once the zero vector has been evaluated
    if (bc != 0 && pred != pos00) {
        const uint32_t c = sea_exact<B, PW>(sP, wx0, wy0, c0, c1, lane);
        evaluated++;
        zpos = pos00;
        if (c < bc || (c == bc && pos00 < bp)) { bc = c; bp = pos00; }
    }
    const uint16_t* scol = sS + byl * B * SP + wx0 - R + lane;   // window sum at (dx = lane - R, dy = -R)
    bool resolved = true, full_scan = false;

    if (bc == 0) {
        // ---- b. a zero-cost match: only a zero-cost candidate with a smaller position code (a shorter vector) wins
        if (bp != pos00) {
            const int d2 = (int)(bp >> 14);
            const int bdx = pdx, bdy = pdy;        // a zero-cost best other than the zero vector is the predictor
            // floor(sqrt(d2)): at least max(|dx|, |dy|), a few steps above it at most
            int r = max(abs(bdx), abs(bdy));
            while ((r + 1) * (r + 1) <= d2) r++;
            const int ya = max(-r, dy_lo), yb = min(r, dy_hi);
            // a window sum is < 2^16: idle lanes never match.  The best's own position always matches, so its column
            // is left out of the row scan and tested on its own below.
            const uint32_t c00l = lane_in && dx != bdx ? C00 : NONE;
            // G rows per vote: one load + one compare per candidate; the other three sub-sums, the position test and
            // the exact cost only where a first sub-sum matches.  Rows past yb are inside the plane (G - 1 <= N) and
            // are masked in the slow path.
            constexpr int G = N >= 8 ? 8 : 4;
            const uint16_t* qg = scol + (ya + R) * SP;
            for (int dyq = ya; dyq <= yb; dyq += G, qg += G * SP) {
                bool any = false;
#pragma unroll
                for (int j = 0; j < G; j++) any |= (uint32_t)qg[j * SP] == c00l;
                if (__ballot_sync(0xffffffffu, any) == 0u) continue;
                uint32_t code = 0;
#pragma unroll
                for (int j = 0; j < G; j++) code |= ((uint32_t)qg[j * SP] == c00l ? 1u : 0u) << j;
                const int nrow = yb - dyq + 1;
                if (nrow < G) code &= (1u << nrow) - 1u;
                unsigned rows = __reduce_or_sync(0xffffffffu, code);
                while (rows) {
                    const int j = __ffs(rows) - 1;
                    rows &= rows - 1;
                    const int dy = dyq + j;
                    const uint16_t* q = qg + j * SP;
                    const uint32_t pos = sea_pos(dx, dy, R);
                    const bool zero = ((code >> j) & 1u) && pos < bp && (uint32_t)q[N] == C10 && (uint32_t)q[N * SP] == C01 &&
                                      (uint32_t)q[N * SP + N] == C11;
                    unsigned m = __ballot_sync(0xffffffffu, zero);
                    while (m) {
                        const int l = __ffs(m) - 1;
                        m &= m - 1;
                        const uint32_t cp = __shfl_sync(0xffffffffu, pos, l);
                        if (cp >= bp) continue;
                        const uint32_t c = sea_exact<B, PW>(sP, wx0 + l - R, wy0 + dy, c0, c1, lane);
                        evaluated++;
                        if (c == 0) bp = cp;
                    }
                }
            }
            {   // the best's own column (its vector may have changed above: `bdx` is the column left out), lanes <-> dy
                const uint16_t* col = sS + (byl * B + R) * SP + wx0 + bdx;
#pragma unroll
                for (int t = 0; t < (2 * R + 1 + 31) / 32; t++) {
                    if (t > 0 && ya + 32 * t > yb) break;
                    const int dy = ya + lane + 32 * t;
                    const uint16_t* q = col + min(dy, yb) * SP;
                    const uint32_t pos = sea_pos(bdx, dy, R);
                    const bool hit = dy <= yb && dy != bdy && (uint32_t)q[0] == C00;
                    if (__ballot_sync(0xffffffffu, hit) == 0u) continue;
                    const bool zero = hit && pos < bp && (uint32_t)q[N] == C10 && (uint32_t)q[N * SP] == C01 &&
                                      (uint32_t)q[N * SP + N] == C11;
                    unsigned m = __ballot_sync(0xffffffffu, zero);
                    while (m) {
                        const int l = __ffs(m) - 1;
                        m &= m - 1;
                        const uint32_t cp = __shfl_sync(0xffffffffu, pos, l);
                        if (cp >= bp) continue;
                        const uint32_t c = sea_exact<B, PW>(sP, wx0 + bdx, wy0 + ya + l + 32 * t, c0, c1, lane);
                        evaluated++;
                        if (c == 0) bp = cp;
                    }
                }
            }
            if (C::EXTRA && R * R <= d2 && R <= dx_hi) {      // column dx = +R: lanes <-> dy
#pragma unroll
                for (int t = 0; t < (C::NEX > 0 ? C::NEX : 1); t++) {
                    const int dyi = lane + 32 * t, dy = dyi - R;
                    const bool ok = dyi < ND && dy >= dy_lo && dy <= dy_hi;
                    const uint16_t* q = sS + (byl * B + (ok ? dyi : 0)) * SP + wx0 + R;
                    const uint32_t pos = sea_pos(R, dy, R);
                    const bool zero = ok && pos < bp && (uint32_t)q[0] == C00 && (uint32_t)q[N] == C10 &&
                                      (uint32_t)q[N * SP] == C01 && (uint32_t)q[N * SP + N] == C11;
                    unsigned m = __ballot_sync(0xffffffffu, zero);
                    while (m) {
                        const int l = __ffs(m) - 1;
                        m &= m - 1;
                        const uint32_t cp = __shfl_sync(0xffffffffu, pos, l);
                        if (cp >= bp) continue;
                        const uint32_t c = sea_exact<B, PW>(sP, wx0 + R, wy0 + l + 32 * t - R, c0, c1, lane);
                        evaluated++;
                        if (c == 0) bp = cp;
                    }
                }
            }
        }
    } else {
        // ---- c. bounds of every candidate into registers: window-sum rows t = 0 .. ND+N-1, row t serves the upper
        // sub-blocks of dy index t and the lower sub-blocks of dy index t-N
        full_scan = true;
        uint32_t b[ND];
#pragma unroll
        for (int t = 0; t < ND + N; t++) {
            const uint32_t sa = scol[t * SP], sb = scol[t * SP + N];
            if (t < ND) b[t] = __usad(sb, C10, __usad(sa, C00, 0u));
            if (t >= N) b[t - N] = __usad(sb, C11, __usad(sa, C01, b[t - N]));
        }
        if (!interior) {
#pragma unroll
            for (int dyi = 0; dyi < ND; dyi++)
                if (!lane_in || dyi - R < dy_lo || dyi - R > dy_hi) b[dyi] = SEA_BIG;
        }
        uint32_t bex[C::NEX > 0 ? C::NEX : 1];
        if (C::EXTRA) {
#pragma unroll
            for (int t = 0; t < C::NEX; t++) {
                const int dyi = lane + 32 * t, dy = dyi - R;
                const bool ok = dyi < ND && dy >= dy_lo && dy <= dy_hi && R <= dx_hi;
                const uint16_t* q = sS + (byl * B + (dyi < ND ? dyi : 0)) * SP + wx0 + R;
                const uint32_t v = __usad(q[N * SP + N], C11, __usad(q[N * SP], C01, __usad(q[N], C10, __usad(q[0], C00, 0u))));
                bex[t] = ok ? v : SEA_BIG;
            }
        }
        // smallest (bound, position) key: per lane fold the dy of its column with key = bound << 7 | rank(dy)
        uint32_t kmin = 0xFFFFFFFFu;
#pragma unroll
        for (int dyi = 0; dyi < ND; dyi++) {
            const int dy = dyi - R;
            kmin = min(kmin, b[dyi] * 128u + (uint32_t)(2 * (dy < 0 ? -dy : dy) - (dy < 0 ? 1 : 0)));
        }
        uint32_t my_lb = SEA_BIG, my_pos = NONE;
        if (lane < C::NL && (kmin >> 7) < SEA_BIG) {
            const int code = (int)(kmin & 127u);
            const int ady = (code + 1) >> 1;
            my_lb = kmin >> 7;
            my_pos = sea_pos(dx, (code & 1) ? -ady : ady, R);
        }
        if (C::EXTRA) {
#pragma unroll
            for (int t = 0; t < C::NEX; t++) {
                const uint32_t pos = sea_pos(R, lane + 32 * t - R, R);
                if (bex[t] < my_lb || (bex[t] == my_lb && bex[t] < SEA_BIG && pos < my_pos)) { my_lb = bex[t]; my_pos = pos; }
            }
        }
        const uint32_t lb_min = __reduce_min_sync(0xffffffffu, my_lb);
        const uint32_t pos_min = __reduce_min_sync(0xffffffffu, my_lb == lb_min ? my_pos : NONE);
        if (lb_min < bc || (lb_min == bc && pos_min < bp)) {
            if (pos_min != zpos && pos_min != pred) {
                const uint32_t c = sea_exact<B, PW>(sP, wx0 + (int)(pos_min & 127u) - R, wy0 + (int)((pos_min >> 7) & 127u) - R, c0, c1, lane);
                evaluated++;
                if (c < bc || (c == bc && pos_min < bp)) { bc = c; bp = pos_min; }
            }
            // candidates whose bound is not above the best cost: one bit per dy (ties are sorted out below)
            uint32_t mlo = 0, mhi = 0;
#pragma unroll
            for (int dyi = 0; dyi < ND; dyi++) {
                if (b[dyi] <= bc) {
                    if (dyi < 32) mlo |= 1u << dyi;
                    else mhi |= 1u << (dyi - 32);
                }
            }
            if (lane >= C::NL) mlo = mhi = 0;
            uint32_t mex = 0;
            if (C::EXTRA) {
#pragma unroll
                for (int t = 0; t < C::NEX; t++)
                    if (bex[t] <= bc) mex |= 1u << t;
            }
            const int mine = __popc(mlo) + __popc(mhi) + __popc(mex);
            const int total = (int)__reduce_add_sync(0xffffffffu, (uint32_t)mine);
            if (total > SEA_CAP) {
                resolved = false;
            } else {
                // exclusive prefix of `mine` over the lanes -> slots of this lane's candidates
                int incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int u = (int)__shfl_up_sync(0xffffffffu, (uint32_t)incl, o);
                    if (lane >= o) incl += u;
                }
                int slot = incl - mine;
                while (mlo) { const int dyi = __ffs(mlo) - 1; mlo &= mlo - 1; s_list[slot++] = ((uint32_t)dyi << 7) | (uint32_t)lane; }
                while (mhi) { const int dyi = 32 + __ffs(mhi) - 1; mhi &= mhi - 1; s_list[slot++] = ((uint32_t)dyi << 7) | (uint32_t)lane; }
                while (mex) { const int t = __ffs(mex) - 1; mex &= mex - 1; s_list[slot++] = ((uint32_t)(lane + 32 * t) << 7) | (uint32_t)(2 * R); }
                __syncwarp();
                sea_eval_list<B, R, N, PW, SP>(s_list, total, sS + byl * B * SP + wx0 - R, sP, wx0, wy0, c0, c1, lane, C00, C10, C01, C11,
                                               zpos, pred, pos_min, bc, bp, evaluated);
                __syncwarp();
            }
        }
    }
    if (out.stats && lane == 0) {
        atomicAdd(&out.stats[0], 1ull);
        atomicAdd(&out.stats[1], resolved ? 1ull : 0ull);
        atomicAdd(&out.stats[2], evaluated);
        atomicAdd(&out.stats[3], full_scan ? 1ull : 0ull);
    }
    SeaResult res;
    res.cost = bc;
    res.pos = bp;
    res.resolved = resolved;
    return res;
}

// ---- step 3 for +-32 (65 x 65 candidates): the lanes <-> dx mapping runs over two groups of 32 columns (dx = lane - R + 32 g)
// plus the dx = +R column, and the bounds of a column no longer fit a warp's registers: the full scan makes two passes
// with a rolling window of N + 1 accumulators (a bound is complete N rows after it was started) — pass 1 folds the
// smallest (bound, position) key, pass 2 recomputes the bounds and collects the survivors.  Everything else is the
// +-16 procedure above.
template <int B, int R, int TH>
__device__ __forceinline__ SeaResult sea_block_wide(const uint16_t* __restrict__ sS, const uint8_t* __restrict__ sP,
                                                    const uint8_t* __restrict__ sC, const uint32_t* __restrict__ s_csum,
                                                    const BlockMatchParams& p, int bx, int by, int bxl, int byl, bool interior,
                                                    uint32_t pred, int lane, uint32_t* __restrict__ s_list, const SeaOut& out)
{
    using C = SeaCfg<B, R, TH>;
    constexpr int N = C::N, ND = C::ND, PW = C::PW, SP = C::SP, NG = (ND - 1) / 32, NEX = C::NEX;
    static_assert(ND == 65 && NG == 2, "two groups of 32 columns and the dx = +R column; 64 + 1 rows in the survivor masks");
    constexpr uint32_t NONE = 0xFFFFFFFFu;
    int dy_lo = -R, dy_hi = R, dx_lo = -R, dx_hi = R;
    if (!interior) {
        const int x0 = bx * B, y0 = by * B;
        dy_lo = max(-R, -p.halo_top - y0);
        dy_hi = min(R, p.strip_h + p.halo_bottom - B - y0);
        dx_lo = max(-R, -x0);
        dx_hi = min(R, p.w - B - x0);
    }
    const int wx0 = bxl * B + C::RA, wy0 = byl * B + R;
    uint32_t c0, c1;
    sea_cur_block<B, C::CH>(sC, bxl, byl, lane, c0, c1);
    const uint4 csum = *reinterpret_cast<const uint4*>(s_csum + 4 * (byl * C::TBX + bxl));
    const uint32_t C00 = csum.x, C10 = csum.y, C01 = csum.z, C11 = csum.w;

    const uint32_t pos00 = sea_pos(0, 0, R);
    int pdx = (int)(pred & 127u) - R, pdy = (int)((pred >> 7) & 127u) - R;
    if (!interior && (pdx < dx_lo || pdx > dx_hi || pdy < dy_lo || pdy > dy_hi)) {
        pred = pos00;
        pdx = pdy = 0;
    }
    uint32_t bc = sea_exact<B, PW>(sP, wx0 + pdx, wy0 + pdy, c0, c1, lane);
    uint32_t bp = pred;
    unsigned long long evaluated = 1;
    uint32_t zpos = pred == pos00 ? pos00 : NONE;
    if (bc != 0 && pred != pos00) {
        const uint32_t c = sea_exact<B, PW>(sP, wx0, wy0, c0, c1, lane);
        evaluated++;
        zpos = pos00;
        if (c < bc || (c == bc && pos00 < bp)) { bc = c; bp = pos00; }
    }
    const uint16_t* scol0 = sS + byl * B * SP + wx0 - R + lane;   // window sum at (dx = lane - R, dy = -R); group g: + 32 g
    bool resolved = true, full_scan = false;

    if (bc == 0) {
        if (bp != pos00) {
            const int d2 = (int)(bp >> 14);
            const int bdx = pdx, bdy = pdy;
            int r = max(abs(bdx), abs(bdy));
            while ((r + 1) * (r + 1) <= d2) r++;
            const int ya = max(-r, dy_lo), yb = min(r, dy_hi);
            constexpr int G = 8;
#pragma unroll
            for (int g = 0; g < NG; g++) {
                if (32 * g - R > r || 32 * g - R + 31 < -r) continue;   // the group holds no column of the disc (warp-uniform)
                const int dx = lane - R + 32 * g;
                const uint32_t c00l = dx >= dx_lo && dx <= dx_hi && dx != bdx ? C00 : NONE;
                const uint16_t* qg = scol0 + 32 * g + (ya + R) * SP;
                for (int dyq = ya; dyq <= yb; dyq += G, qg += G * SP) {
                    bool any = false;
#pragma unroll
                    for (int j = 0; j < G; j++) any |= (uint32_t)qg[j * SP] == c00l;
                    if (__ballot_sync(0xffffffffu, any) == 0u) continue;
                    uint32_t code = 0;
#pragma unroll
                    for (int j = 0; j < G; j++) code |= ((uint32_t)qg[j * SP] == c00l ? 1u : 0u) << j;
                    const int nrow = yb - dyq + 1;
                    if (nrow < G) code &= (1u << nrow) - 1u;
                    unsigned rows = __reduce_or_sync(0xffffffffu, code);
                    while (rows) {
                        const int j = __ffs(rows) - 1;
                        rows &= rows - 1;
                        const int dy = dyq + j;
                        const uint16_t* q = qg + j * SP;
                        const uint32_t pos = sea_pos(dx, dy, R);
                        const bool zero = ((code >> j) & 1u) && pos < bp && (uint32_t)q[N] == C10 && (uint32_t)q[N * SP] == C01 &&
                                          (uint32_t)q[N * SP + N] == C11;
                        unsigned m = __ballot_sync(0xffffffffu, zero);
                        while (m) {
                            const int l = __ffs(m) - 1;
                            m &= m - 1;
                            const uint32_t cp = __shfl_sync(0xffffffffu, pos, l);
                            if (cp >= bp) continue;
                            const uint32_t c = sea_exact<B, PW>(sP, wx0 + l - R + 32 * g, wy0 + dy, c0, c1, lane);
                            evaluated++;
                            if (c == 0) bp = cp;
                        }
                    }
                }
            }
            // the best's own column, and the dx = +R column when the disc reaches it: lanes <-> dy
#pragma unroll
            for (int which = 0; which < 2; which++) {
                const int cdx = which == 0 ? bdx : R;
                if (which == 1 && (R * R > d2 || R > dx_hi || bdx == R)) continue;
                const uint16_t* col = sS + (byl * B + R) * SP + wx0 + cdx;
#pragma unroll
                for (int t = 0; t < NEX; t++) {
                    if (t > 0 && ya + 32 * t > yb) break;
                    const int dy = ya + lane + 32 * t;
                    const uint16_t* q = col + min(dy, yb) * SP;
                    const uint32_t pos = sea_pos(cdx, dy, R);
                    const bool hit = dy <= yb && !(cdx == bdx && dy == bdy) && (uint32_t)q[0] == C00;
                    if (__ballot_sync(0xffffffffu, hit) == 0u) continue;
                    const bool zero = hit && pos < bp && (uint32_t)q[N] == C10 && (uint32_t)q[N * SP] == C01 &&
                                      (uint32_t)q[N * SP + N] == C11;
                    unsigned m = __ballot_sync(0xffffffffu, zero);
                    while (m) {
                        const int l = __ffs(m) - 1;
                        m &= m - 1;
                        const uint32_t cp = __shfl_sync(0xffffffffu, pos, l);
                        if (cp >= bp) continue;
                        const uint32_t c = sea_exact<B, PW>(sP, wx0 + cdx, wy0 + ya + l + 32 * t, c0, c1, lane);
                        evaluated++;
                        if (c == 0) bp = cp;
                    }
                }
            }
        }
    } else {
        full_scan = true;
        // ---- pass 1: per column, the smallest bound (key bound << 7 | dy index) and the candidates whose bound is not
        // above the cost of the best so far.  RS rows per trip of a rolled loop (unrolled, the loads are hoisted and the
        // ring of accumulators spills); a trip whose rows are all legal skips the range test.
        constexpr int RS = N + 1 <= 5 ? 5 : 13;   // ring slots: a divisor of ND not below N + 1
        static_assert(ND % RS == 0 && RS >= N + 1 && RS <= 32, "whole trips, one code bit per row");
        const uint16_t* s_blk = scol0 - lane;
        uint32_t bex[NEX];
        uint32_t my_lb = SEA_BIG, my_pos = NONE;
#pragma unroll
        for (int t = 0; t < NEX; t++) {
            const int dyi = lane + 32 * t, dy = dyi - R;
            const bool ok = dyi < ND && dy >= dy_lo && dy <= dy_hi && R <= dx_hi;
            const uint16_t* q = s_blk + (dyi < ND ? dyi : 0) * SP + 2 * R;
            const uint32_t v = __usad(q[N * SP + N], C11, __usad(q[N * SP], C01, __usad(q[N], C10, __usad(q[0], C00, 0u))));
            bex[t] = ok ? v : SEA_BIG;
            const uint32_t pos = sea_pos(R, dy, R);
            if (bex[t] < my_lb || (bex[t] == my_lb && bex[t] < SEA_BIG && pos < my_pos)) { my_lb = bex[t]; my_pos = pos; }
        }
        const int lo_i = dy_lo + R;
        const uint32_t span = (uint32_t)(dy_hi - dy_lo);
        uint32_t pos_min = NONE;
        int total = 0;
        unsigned long long msk[NG];   // bit dyi of the group's column; the last row (dyi = 2 R) in top
        uint32_t top = 0, mex = 0;
        int mine = 0;
#pragma unroll 1
        for (int pass = 0; pass < 2; pass++) {
            mine = 0;
            top = 0;
#pragma unroll
            for (int g = 0; g < NG; g++) {
                const int dx = lane - R + 32 * g;
                const bool lin = dx >= dx_lo && dx <= dx_hi;
                const uint16_t* q = scol0 + 32 * g;
                uint32_t acc[RS], kmin = NONE, code = 0;
                unsigned long long mk = 0;
#pragma unroll
                for (int t = 0; t < N; t++) acc[t] = __usad(q[t * SP + N], C10, __usad(q[t * SP], C00, 0u));
                q += N * SP;
#pragma unroll 1
                for (int d0 = 0; d0 < ND; d0 += RS, q += RS * SP) {
                    uint32_t kc = NONE;
                    code = 0;
                    if (interior || (d0 >= lo_i && d0 + RS - 1 <= lo_i + (int)span))
                        sea_scan_chunk<N, SP, RS, false>(q, acc, C00, C10, C01, C11, bc, 0, 0u, kc, code);
                    else
                        sea_scan_chunk<N, SP, RS, true>(q, acc, C00, C10, C01, C11, bc, lo_i - d0, span, kc, code);
                    kmin = min(kmin, kc + (uint32_t)d0);
                    mk |= (unsigned long long)code << d0;
                }
                if (lin) {
                    msk[g] = mk;
                    top |= ((code >> (RS - 1)) & 1u) << g;   // the last trip's last row: dy index 2 R
                    if ((kmin >> 7) < SEA_BIG) {
                        const uint32_t lb = kmin >> 7, pos = sea_pos(dx, (int)(kmin & 127u) - R, R);
                        if (lb < my_lb || (lb == my_lb && pos < my_pos)) { my_lb = lb; my_pos = pos; }
                    }
                } else {
                    msk[g] = 0;
                }
                mine += __popcll(msk[g]) + (int)((top >> g) & 1u);
            }
            mex = 0;
#pragma unroll
            for (int t = 0; t < NEX; t++)
                if (bex[t] <= bc) mex |= 1u << t;
            mine += __popc(mex);
            total = (int)__reduce_add_sync(0xffffffffu, (uint32_t)mine);
            if (total <= SEA_CAP || pass == 1) break;
            // too many survivors under the predictor's cost: evaluate the candidate of the smallest bound and count again
            const uint32_t lb_min = __reduce_min_sync(0xffffffffu, my_lb);
            pos_min = __reduce_min_sync(0xffffffffu, my_lb == lb_min ? my_pos : NONE);
            if (lb_min > bc || pos_min == zpos || pos_min == pred) break;
            const uint32_t c = sea_exact<B, PW>(sP, wx0 + (int)(pos_min & 127u) - R, wy0 + (int)((pos_min >> 7) & 127u) - R, c0, c1, lane);
            evaluated++;
            if (c < bc || (c == bc && pos_min < bp)) { bc = c; bp = pos_min; }
            else break;   // nothing tightened: the count stands
        }
        if (total > SEA_CAP) {
            resolved = false;
        } else if (total > 0) {
            int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = (int)__shfl_up_sync(0xffffffffu, (uint32_t)incl, o);
                if (lane >= o) incl += u;
            }
            int slot = incl - mine;
#pragma unroll
            for (int g = 0; g < NG; g++) {
                unsigned long long m = msk[g];
                while (m) {
                    const int dyi = __ffsll((long long)m) - 1;
                    m &= m - 1;
                    s_list[slot++] = ((uint32_t)dyi << 7) | (uint32_t)(lane + 32 * g);
                }
                if ((top >> g) & 1u) s_list[slot++] = ((uint32_t)(2 * R) << 7) | (uint32_t)(lane + 32 * g);
            }
            while (mex) { const int t = __ffs(mex) - 1; mex &= mex - 1; s_list[slot++] = ((uint32_t)(lane + 32 * t) << 7) | (uint32_t)(2 * R); }
            __syncwarp();
            sea_eval_list<B, R, N, PW, SP>(s_list, total, s_blk, sP, wx0, wy0, c0, c1, lane, C00, C10, C01, C11, zpos, pred, pos_min, bc, bp,
                                           evaluated);
            __syncwarp();
        }
    }
    if (out.stats && lane == 0) {
        atomicAdd(&out.stats[0], 1ull);
        atomicAdd(&out.stats[1], resolved ? 1ull : 0ull);
        atomicAdd(&out.stats[2], evaluated);
        atomicAdd(&out.stats[3], full_scan ? 1ull : 0ull);
    }
    SeaResult res;
    res.cost = bc;
    res.pos = bp;
    res.resolved = resolved;
    return res;
}

#ifdef OFPSB_EMU
struct SeaMaps { int unused; };
#else
struct SeaMaps { CUtensorMap prev, cur, prev8, up8, down8; };   // *8: boxes of 8 rows (tiles at a strip seam)
#endif

template <int B, int R, int TH>
__global__ void __launch_bounds__(SEA_NT, 3) sea_kernel(const __grid_constant__ SeaMaps maps, const BlockMatchParams p,
                                                        const SeaOut out)
{
    using C = SeaCfg<B, R, TH>;
    OFPSB_DYN_SMEM(smem);
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_list[SEA_WARPS][SEA_CAP];
    __shared__ uint32_t s_probe[SEA_WARPS];
    __shared__ __align__(16) uint32_t s_csum[4 * C::TBX * C::TBY];
#ifndef OFPSB_EMU
    uint8_t* sP = smem + ((128u - (smem_u32(smem) & 127u)) & 127u);   // TMA destinations: 128-byte aligned
#else
    uint8_t* sP = smem;
#endif
    uint32_t* sS32 = reinterpret_cast<uint32_t*>(sP + C::P_BYTES);
    uint8_t* sC = sP + C::P_BYTES + C::S_BYTES;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // peer-halo mode: the tile rows at the two seams go first (last row, then row 0, 1, ...) so that their loads from
    // the neighbours' memory overlap the rest of the strip instead of forming its tail
    const int tile_row = out.peer ? (int)((blockIdx.y + gridDim.y - 1) % gridDim.y) : (int)blockIdx.y;
    const int tx0 = blockIdx.x * SEA_TILE_W, ty0 = tile_row * TH, pair = blockIdx.z;
    // tensor row 0 of prev = first halo row (halo rows stored with the strip) or first own row (peer-halo mode)
    const int wx = tx0 - C::RA, wy = ty0 - R + (out.peer ? 0 : p.halo_top);
#ifndef OFPSB_EMU
    if (tid == 0) {
        const uint32_t b32 = smem_u32(&bar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b32));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b32), "r"(C::TX_BYTES) : "memory");
        if (out.peer && (wy < 0 || wy + C::PH > out.own_rows)) {
            // a tile at a strip seam: 8 rows at a time, each group from the tensor that owns it — the neighbour's HBM
            // over NVLink for the halo rows, zeros (out-of-tensor box) where the frame ends
            static_assert(C::PH % 8 == 0 && (R % 8) == 0, "seam tiles are loaded in groups of 8 rows");
            for (int g = 0; g < C::PH / 8; g++) {
                const int y = wy + 8 * g;
                const uint32_t dst = smem_u32(sP + g * 8 * C::PW);
                if (y < 0 && out.has_up) tma_load_3d(dst, &maps.up8, wx, out.up_rows + y, pair, b32);
                else if (y >= out.own_rows && out.has_down) tma_load_3d(dst, &maps.down8, wx, y - out.own_rows, pair, b32);
                else tma_load_3d(dst, &maps.prev8, wx, y, pair, b32);
            }
        } else {
            tma_load_3d(smem_u32(sP), &maps.prev, wx, wy, pair, b32);
        }
        tma_load_4d(smem_u32(sC), &maps.cur, 0, ty0, tx0 / 16, pair, b32);
        // the frames stream through the L2 once: pull the boxes of the tile a CTA that starts about one wave later will
        // load into the L2 now, so that its loads do not wait for HBM
        if (out.prefetch_tiles > 0) {
            const unsigned per_pair = gridDim.x * gridDim.y;
            const unsigned t = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x + (unsigned)out.prefetch_tiles;
            const unsigned tz = t / per_pair, trem = t - tz * per_pair;
            if (tz < gridDim.z) {
                const int py = (int)(trem / gridDim.x), px = (int)(trem - (unsigned)py * gridDim.x);
                tma_prefetch_3d(&maps.prev, px * SEA_TILE_W - C::RA, py * TH - R + (out.peer ? 0 : p.halo_top), (int)tz);
                tma_prefetch_4d(&maps.cur, 0, py * TH, px * (SEA_TILE_W / 16), (int)tz);
            }
        }
    }
    __syncthreads();
    mbar_wait(smem_u32(&bar), 0);
#else
    (void)maps;
    (void)bar;
    {   // stand-in for the two TMA box loads: out-of-tensor bytes are zeros
        const int rows_prev = p.halo_top + p.strip_h + p.halo_bottom;
        const uint8_t* pb = p.prev - (long long)p.halo_top * p.stride + (long long)pair * p.pair_stride;
        const uint8_t* cb = p.cur + (long long)pair * p.pair_stride;
        for (int i = tid; i < C::PW * C::PH; i += SEA_NT) {
            const int y = wy + i / C::PW, x = wx + i % C::PW;
            sP[i] = (x >= 0 && x < p.w && y >= 0 && y < rows_prev) ? pb[(long long)y * p.stride + x] : 0;
        }
        for (int i = tid; i < C::CW * C::CH; i += SEA_NT) {   // [column block][row][16]; whole column blocks of the tensor
            const int c16 = i / (C::CH * 16), row = (i / 16) % C::CH, b = i % 16;
            const int y = ty0 + row, x = tx0 + c16 * 16 + b;
            sC[i] = (tx0 / 16 + c16 < (p.w + 15) / 16 && y < p.strip_h) ? cb[(long long)y * p.stride + x] : 0;
        }
    }
    __syncthreads();
#endif
    if (out.debug_stop == 1) return;
    sea_hpass<C::N, C::PW, C::PH, C::SP>(sP, sS32, tid);
    sea_cur_sums<C, B>(sC, s_csum, tid);
    __syncthreads();
    sea_vpass<C>(sS32, tid);
    __syncthreads();

    if (out.debug_stop == 2) return;
    // the whole tile is interior when no candidate of any of its blocks leaves the frame (CTA-uniform)
    const bool interior = tx0 - R >= 0 && tx0 + SEA_TILE_W + R <= p.w && ty0 - R >= -p.halo_top &&
                          ty0 + TH + R <= p.strip_h + p.halo_bottom &&
                          tx0 / B + C::TBX <= p.nbx && ty0 / B + C::TBY <= p.nby;
    const uint16_t* sS = reinterpret_cast<const uint16_t*>(sS32);

    // tile predictor: the candidate of the tile's first block with the smallest four-term bound, found by the whole CTA
    // (a few candidates per thread).  It only seeds the first block of every warp — a predictor changes the work,
    // never the result — and replaces a full scan per warp and tile (ncu: a quarter of all blocks before this).
    {
        constexpr int N = C::N, ND = C::ND, SP = C::SP;
        const uint4 cs = *reinterpret_cast<const uint4*>(s_csum);
        const int dy_lo = max(-R, -p.halo_top - ty0), dy_hi = min(R, p.strip_h + p.halo_bottom - B - ty0);
        const int dx_lo = max(-R, -tx0), dx_hi = min(R, p.w - B - tx0);
        uint32_t kb = 0xFFFFFFFFu;
        for (int idx = tid; idx < ND * ND; idx += SEA_NT) {
            const int dyi = idx / ND, dxi = idx - dyi * ND;
            const uint16_t* q = sS + dyi * SP + C::RA - R + dxi;
            const uint32_t v = __usad(q[N * SP + N], cs.w, __usad(q[N * SP], cs.z, __usad(q[N], cs.y, __usad(q[0], cs.x, 0u))));
            const bool ok = interior || (dxi - R >= dx_lo && dxi - R <= dx_hi && dyi - R >= dy_lo && dyi - R <= dy_hi);
            if (ok) kb = min(kb, (v << 13) | (uint32_t)idx);   // v < 2^16, idx < 65 * 65 < 2^13
        }
        kb = __reduce_min_sync(0xffffffffu, kb);
        if (lane == 0) s_probe[warp] = kb;
    }
    __syncthreads();
    uint32_t pred;
    {
        uint32_t kb = s_probe[0];
#pragma unroll
        for (int i = 1; i < SEA_WARPS; i++) kb = min(kb, s_probe[i]);
        const int idx = (int)(kb & 8191u), dyi = idx / C::ND;
        pred = sea_pos(idx - dyi * C::ND - R, dyi - R, R);
    }
    if (out.debug_stop == 3) return;
    // the warp's blocks; lane `it` keeps the result of block `it` and writes it after the loop (one pass of the output
    // code per warp instead of one per block)
    constexpr int NBW = C::TBY * C::CPW;
    static_assert(NBW <= 32, "one lane per block of the warp");
    uint32_t r_cost = 0, r_pos = 0;
    int r_state = 0;   // 0 = no block, 1 = resolved, 2 = exhaustive work list
    for (int it = 0; it < NBW; it++) {
        const int byl = it / C::CPW, bxl = warp * C::CPW + it % C::CPW;
        const int bx = tx0 / B + bxl, by = ty0 / B + byl;
        if (bx >= p.nbx || by >= p.nby) continue;
        SeaResult res;
        if constexpr (C::ND > 33) res = sea_block_wide<B, R, TH>(sS, sP, sC, s_csum, p, bx, by, bxl, byl, interior, pred, lane, s_list[warp], out);
        else res = sea_block<B, R, TH>(sS, sP, sC, s_csum, p, bx, by, bxl, byl, interior, pred, lane, s_list[warp], out);
        pred = res.pos;
        if (lane == it) {
            r_cost = res.cost;
            r_pos = res.pos;
            r_state = res.resolved ? 1 : 2;
        }
    }
    if (r_state) {
        const int byl = lane / C::CPW, bxl = warp * C::CPW + lane % C::CPW;
        const int bx = tx0 / B + bxl, by = ty0 / B + byl;
        const size_t gb = (size_t)((uint32_t)(pair * p.nby + by) * (uint32_t)p.nbx + (uint32_t)bx);   // < 2^32 (launcher)
        if (r_state == 1) sea_write<R>(p, gb, r_cost, r_pos, bx, by, out.nx, out.ny);
        else out.worklist[atomicAdd(out.wl_count, 1u)] = (uint32_t)gb;
    }
}

#ifndef OFPSB_EMU
template <int B, int R, int TH>
int launch_sea_th(const BlockMatchParams& p, const SeaOut& out, const SeaPeer* peer, cudaStream_t stream)
{
    using C = SeaCfg<B, R, TH>;
    SeaMaps maps;
    if (peer) {
        if (!make_map(&maps.prev, p.prev, p.w, peer->own_rows, p.stride, p.pair_stride, p.n_pairs, C::PW, C::PH) ||
            !make_map(&maps.prev8, p.prev, p.w, peer->own_rows, p.stride, p.pair_stride, p.n_pairs, C::PW, 8))
            return 1;
        maps.up8 = maps.down8 = maps.prev8;
        if (peer->up && !make_map(&maps.up8, peer->up, p.w, peer->up_rows, peer->up_stride, peer->up_pair_stride, p.n_pairs, C::PW, 8))
            return 1;
        if (peer->down && !make_map(&maps.down8, peer->down, p.w, peer->down_rows, peer->down_stride, peer->down_pair_stride, p.n_pairs, C::PW, 8))
            return 1;
    } else {
        const int rows_prev = p.halo_top + p.strip_h + p.halo_bottom;
        const uint8_t* prev_base = p.prev - (long long)p.halo_top * p.stride;
        if (!make_map(&maps.prev, prev_base, p.w, rows_prev, p.stride, p.pair_stride, p.n_pairs, C::PW, C::PH)) return 1;
        maps.prev8 = maps.up8 = maps.down8 = maps.prev;
    }
    if (!make_map_colblocks(&maps.cur, p.cur, p.w, p.strip_h, p.stride, p.pair_stride, p.n_pairs, C::CH, C::CW / 16)) return 1;
    static bool attr_set[64] = {};
    int dev = 0;
    OFPSB_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        OFPSB_CUDA_TRY(cudaFuncSetAttribute(sea_kernel<B, R, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    if (getenv("OFPSB_DEBUG_OCC")) {
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, sea_kernel<B, R, TH>, SEA_NT, C::SMEM_BYTES);
        fprintf(stderr, "sea_kernel<%d,%d,%d>: %d CTAs/SM, %d B dynamic smem\n", B, R, TH, nb, C::SMEM_BYTES);
    }
    const dim3 grid((p.nbx * B + SEA_TILE_W - 1) / SEA_TILE_W, (p.nby * B + TH - 1) / TH, p.n_pairs);
    sea_kernel<B, R, TH><<<grid, SEA_NT, C::SMEM_BYTES, stream>>>(maps, p, out);
    OFPSB_CUDA_TRY(cudaGetLastError());
    return OFPSB_OK;
}
// 64-row tiles amortise the window sums best; a launch of fewer tiles than about two waves of resident CTAs (a single
// 1080p pair, a strip of a tiled frame) is bound by the latency of one CTA instead: half-height tiles halve it.
template <int B, int R>
int launch_sea(const BlockMatchParams& p, const SeaOut& out, const SeaPeer* peer, cudaStream_t stream, int sm_count, int tile_h)
{
    const long long tiles64 = (long long)((p.nbx * B + SEA_TILE_W - 1) / SEA_TILE_W) * ((p.nby * B + 63) / 64) * p.n_pairs;
    const bool small = tile_h == 32 || (tile_h == 0 && tiles64 <= 6ll * (sm_count > 0 ? sm_count : 148));
    if constexpr (R > 16)   // +-32: 64-row tiles leave two CTAs per SM (measured: 4K 8x8 176 vs 149 us per pair, 1080p 16x16 equal)
        return launch_sea_th<B, R, 32>(p, out, peer, stream);
    else return small ? launch_sea_th<B, R, 32>(p, out, peer, stream) : launch_sea_th<B, R, 64>(p, out, peer, stream);
}
#endif

}  // namespace

#ifndef OFPSB_EMU
bool block_match_tma_usable(const BlockMatchParams& p);
int launch_block_match_list(const BlockMatchParams& p, const uint32_t* d_list, const uint32_t* d_count, int sm_count,
                            cudaStream_t stream);

// Fused SEA search of one batch.  Returns 0 when launched, 1 when the path does not apply (metric, geometry,
// alignment) — the caller then tries the other paths — and < 0 on error.
int launch_block_match_sea(const BlockMatchParams& p, BlockMatchScratch& sc, int sm_count, cudaStream_t stream,
                           uint64_t* launches, const SeaPeer* peer, cudaEvent_t before_list)
{
    if (p.metric != OFPSB_METRIC_SAD || !block_match_tma_usable(p)) return 1;
    const bool geom = (p.block == 16 || p.block == 8) && (p.range == 8 || p.range == 16 || p.range == 32);
    if (!geom) return 1;
    const int rows_prev = p.halo_top + p.strip_h + p.halo_bottom;
    if (p.w < p.block || rows_prev < p.block) return 1;
    const long long total = (long long)p.nbx * p.nby * p.n_pairs;
    if (total >= 0xFFFFFFF0ll) return 1;
    if (peer && ((reinterpret_cast<uintptr_t>(peer->up) | reinterpret_cast<uintptr_t>(peer->down) | (uintptr_t)peer->up_stride |
                  (uintptr_t)peer->down_stride) & 15))
        return 1;
    // content feedback from the previous SEA launch (never blocks: an unfinished read-back is simply not used yet)
    if (sc.adaptive && !peer && !sc.collect_stats) {
        if (sc.ev_listed && sc.listed_total > 0 && cudaEventQuery(sc.ev_listed) == cudaSuccess) {
            // the +-32 kernel costs a larger share of the exhaustive search it replaces (4K 8x8: 0.5-0.8 on noisy content)
            // two launches in a row over the threshold (a scene cut in a frame-by-frame stream is ONE bad pair and must not
            // send the next 15 good ones to the exhaustive kernel); the streak survives the skipped launches, so on steadily
            // noisy content one probe launch in 16 keeps checking
            if ((double)sc.h_listed[0] > (p.range > 16 ? 0.3 : 0.6) * (double)sc.listed_total) {
                if (++sc.over_streak >= 2) sc.skip_calls = 15;
            } else {
                sc.over_streak = 0;
            }
            sc.listed_total = 0;
        }
        cudaGetLastError();
        if (sc.skip_calls > 0) {
            sc.skip_calls--;
            return 1;   // the caller runs the exhaustive kernels
        }
    }
    // scratch: [0] work-list count, [2..9] four 64-bit statistics, then the work list
    const size_t head = 16;
    if (int rc = sc.worklist.reserve((head + (size_t)total + 8) * sizeof(uint32_t))) return rc;
    uint32_t* base = sc.worklist.as<uint32_t>();
    SeaOut out;
    out.wl_count = base;
    out.stats = sc.collect_stats ? reinterpret_cast<unsigned long long*>(base + 2) : nullptr;
    out.nx = 1.0f / (float)p.w;          // av-decoder/src/lib.rs:404-405 (host code is built with -ffp-contract=off)
    out.ny = 1.0f / (float)p.full_h;
    out.prefetch_tiles = sc.prefetch_tiles >= 0 ? sc.prefetch_tiles : 3 * (sm_count > 0 ? sm_count : 148);
    out.debug_stop = getenv("OFPSB_DEBUG_SEA_STOP") ? atoi(getenv("OFPSB_DEBUG_SEA_STOP")) : 0;
    out.peer = peer ? 1 : 0;
    out.own_rows = peer ? peer->own_rows : 0;
    out.up_rows = peer ? peer->up_rows : 0;
    out.has_up = peer && peer->up ? 1 : 0;
    out.has_down = peer && peer->down ? 1 : 0;
    if (peer) out.prefetch_tiles = 0;
    out.worklist = base + head;
    OFPSB_CUDA_TRY(cudaMemsetAsync(base, 0, head * sizeof(uint32_t), stream));
    if (sc.profile && sc.ev[0]) OFPSB_CUDA_TRY(cudaEventRecord(sc.ev[0], stream));
    int rc = 1;
    if (p.block == 16 && p.range == 16) rc = launch_sea<16, 16>(p, out, peer, stream, sm_count, sc.tile_h);
    else if (p.block == 16 && p.range == 8) rc = launch_sea<16, 8>(p, out, peer, stream, sm_count, sc.tile_h);
    else if (p.block == 8 && p.range == 16) rc = launch_sea<8, 16>(p, out, peer, stream, sm_count, sc.tile_h);
    else if (p.block == 8 && p.range == 8) rc = launch_sea<8, 8>(p, out, peer, stream, sm_count, sc.tile_h);
    else if (p.block == 8 && p.range == 32) rc = launch_sea<8, 32>(p, out, peer, stream, sm_count, sc.tile_h);
    else if (p.block == 16 && p.range == 32) rc = launch_sea<16, 32>(p, out, peer, stream, sm_count, sc.tile_h);
    if (rc) return rc;
    if (sc.profile && sc.ev[1]) OFPSB_CUDA_TRY(cudaEventRecord(sc.ev[1], stream));
    // the exhaustive kernel reads halo rows stored with the strip: in peer-halo mode they are copied on a side
    // stream while the SEA kernel runs; `before_list` marks that copy
    if (before_list) OFPSB_CUDA_TRY(cudaStreamWaitEvent(stream, before_list, 0));
    rc = launch_block_match_list(p, out.worklist, out.wl_count, sm_count, stream);
    if (rc != OFPSB_OK) {
        set_error("block_match: work-list kernel unavailable for block=%d range=%d", p.block, p.range);
        return rc < 0 ? rc : OFPSB_E_INVALID;
    }
    if (sc.profile && sc.ev[2]) OFPSB_CUDA_TRY(cudaEventRecord(sc.ev[2], stream));
    if (sc.adaptive && !peer && !sc.collect_stats) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(stream, &cap);
        if (cap == cudaStreamCaptureStatusNone) {
            if (!sc.h_listed) {
                if (cudaHostAlloc(reinterpret_cast<void**>(&sc.h_listed), 64, cudaHostAllocDefault) != cudaSuccess ||
                    cudaEventCreateWithFlags(&sc.ev_listed, cudaEventDisableTiming) != cudaSuccess) {
                    cudaGetLastError();
                    sc.adaptive = 0;
                }
            }
            if (sc.adaptive) {
                OFPSB_CUDA_TRY(cudaMemcpyAsync(sc.h_listed, out.wl_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
                OFPSB_CUDA_TRY(cudaEventRecord(sc.ev_listed, stream));
                sc.listed_total = total;
            }
        }
    }
    if (launches) *launches += 2;
    return OFPSB_OK;
}
#endif

}  // namespace ofpsb
