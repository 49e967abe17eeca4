// K7-K9: the dense-flow front end of the reference's cv-decoder — what
// CvDecoder::process_frame does around the third-party optical-flow call:
//
//   K7 frame_convert_kernel   cvtColor(BGR2GRAY) (cv-decoder/src/lib.rs:138) + the RGBA out_frame (:145-153)
//   K8 contrast_mask_kernel   Sobel(1,1,k5) -> threshold(>20) -> dilate(11x11 ellipse), fused (:204-236)
//   K9 flow_cells_kernel      masked dense flow -> down-sampling MotionFieldDensifier (:246-274), bit-exact
//      flow_emit_cells_kernel touched cells in BTreeSet<(x,y)> order -> MotionEntry (:276-289)
//      flow_pixels_*          per-pixel variant (process_fullres = false, :272): ordered stream compaction
//
// All three are one-pass stencils / streams over the image grid, HBM-bound by design: K7 reads 3-4 and
// writes 1 (+4) bytes per pixel, K8 reads 1 and writes 1, K9 reads 8 (+1) bytes per pixel and writes a
// few KB.  Integer results are bit-exact with OpenCV (oracle/cv_front.c is pinned on cv2 outputs); the f32
// sums of K9 follow the reference's raster order with un-fused arithmetic (--fmad=false, explicit _rn ops).
#include "common.cuh"

#include <type_traits>

namespace ofpsb {

namespace {

// cv::borderInterpolate(p, len, BORDER_REFLECT_101)
__device__ __forceinline__ int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

// Global -> shared copies that do not pass through registers (LDGSTS): every copy a thread issues is in flight at
// once, so a staging loop costs one memory latency instead of one per iteration.  Source and destination must be
// aligned to the copy size.
__device__ __forceinline__ void async_copy8(void* smem, const void* gmem)
{
#ifdef OFPSB_EMU
    memcpy(smem, gmem, 8);
#else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
#endif
}
__device__ __forceinline__ void async_copy16(void* smem, const void* gmem)
{
#ifdef OFPSB_EMU
    memcpy(smem, gmem, 16);
#else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
#endif
}
// shared-memory destination as a plain byte address: computed once per step, advanced by adds in the copy loops
#ifdef OFPSB_EMU
typedef uint8_t* smem_addr_t;
__device__ __forceinline__ smem_addr_t smem_addr(void* p) { return static_cast<uint8_t*>(p); }
__device__ __forceinline__ void async_copy16_to(smem_addr_t dst, const void* gmem) { memcpy(dst, gmem, 16); }
#else
typedef uint32_t smem_addr_t;
__device__ __forceinline__ smem_addr_t smem_addr(void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void async_copy16_to(smem_addr_t dst, const void* gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem) : "memory");
}
#endif
__device__ __forceinline__ void async_copy_commit()
{
#ifndef OFPSB_EMU
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int PENDING>   // wait until at most PENDING of this thread's committed groups are still in flight
__device__ __forceinline__ void async_copy_wait()
{
#ifndef OFPSB_EMU
    asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory");
#endif
}

// dot product of four UNSIGNED bytes (a) with four SIGNED bytes (b), plus c
__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c)
{
#ifdef OFPSB_EMU
    for (int i = 0; i < 4; i++) c += (int)((a >> (8 * i)) & 255u) * (int)(int8_t)((b >> (8 * i)) & 255u);
    return c;
#else
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#endif
}

// c + a.lo16 * b.byte0 + a.hi16 * b.byte1 (LO) or ... * b.byte2, b.byte3 (HI), everything unsigned
template <bool HI>
__device__ __forceinline__ uint32_t dp2a_uu(uint32_t a, uint32_t b, uint32_t c)
{
#ifdef OFPSB_EMU
    const int o = HI ? 16 : 0;
    return c + (a & 0xFFFFu) * ((b >> o) & 255u) + (a >> 16) * ((b >> (o + 8)) & 255u);
#else
    uint32_t d;
    if (HI) asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    else asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#endif
}

// ------------------------------------------------------------------------------------------ K7
// OpenCV 4.x 8-bit BGR2GRAY: (B*3735 + G*19235 + R*9798 + 2^14) >> 15.
// One thread = PX adjacent pixels of one row; blockIdx.y walks the rows (no index division).  VEC (PX = 16): rows of
// src and gray start on 16-byte boundaries, so the 48 / 64 source bytes are three / four 16-byte loads — enough
// bytes in flight per thread to cover the HBM latency — the 16 gray bytes one 16-byte store and the RGBA pixels
// four.  Otherwise (PX = 4) bytes are gathered one at a time; the ragged row end always takes that path.
template <int CH, int PX, bool VEC>
__global__ void __launch_bounds__(128) frame_convert_kernel(const uint8_t* __restrict__ src, int w, int h, int stride,
                                                            int rgb_order, uint8_t* __restrict__ gray, int gray_stride,
                                                            uint32_t* __restrict__ rgba, int rgba_vec)
{
    constexpr int NW = PX * CH / 4;   // source words per thread
    const int x = PX * (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (x >= w) return;
    const int nv = min(PX, w - x);
    for (int y = blockIdx.y; y < h; y += gridDim.y) {
        const uint8_t* s = src + (size_t)y * stride + (size_t)x * CH;
        uint32_t wd[NW];
        if (VEC && nv == PX) {
#pragma unroll
            for (int k = 0; k < NW / 4; k++) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(s) + k);
                wd[4 * k] = v.x; wd[4 * k + 1] = v.y; wd[4 * k + 2] = v.z; wd[4 * k + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < NW; k++) wd[k] = 0u;
#pragma unroll
            for (int i = 0; i < PX * CH; i++)
                if (i < nv * CH) wd[i >> 2] |= (uint32_t)__ldg(s + i) << (8 * (i & 3));
        }
        // pixel k as one word (c0, c1, c2, x): its three channels start at byte k*CH of the thread's words
        auto pixel = [&](int k) -> uint32_t {
            if (CH == 4) return wd[k];
            const int i = (3 * k) >> 2, sh = 8 * ((3 * k) & 3);
            return sh == 0 ? wd[i] : __funnelshift_r(wd[i], wd[i + 1 < NW ? i + 1 : i], sh);
        };
        // luma weights as 16-bit pairs for dp2a: (c0, c1) and (c2, -)
        const uint32_t k01 = rgb_order ? (9798u | (19235u << 16)) : (3735u | (19235u << 16));
        const uint32_t k2 = rgb_order ? 3735u : 9798u;
        if (gray) {
            uint32_t g[PX / 4];
#pragma unroll
            for (int q = 0; q < PX / 4; q++) {
                g[q] = 0u;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t p = pixel(4 * q + j);
                    g[q] |= (dp2a_uu<true>(k2, p, dp2a_uu<false>(k01, p, 16384u)) >> 15) << (8 * j);
                }
            }
            uint8_t* o = gray + (size_t)y * gray_stride + x;
            if (VEC && nv == PX) {
                if (PX == 16) *reinterpret_cast<uint4*>(o) = make_uint4(g[0], g[PX / 4 > 1 ? 1 : 0], g[PX / 4 > 2 ? 2 : 0], g[PX / 4 > 3 ? 3 : 0]);
                else *reinterpret_cast<uint32_t*>(o) = g[0];
            } else {
#pragma unroll
                for (int k = 0; k < PX; k++)
                    if (k < nv) o[k] = (uint8_t)(g[k >> 2] >> (8 * (k & 3)));
            }
        }
        if (rgba) {   // RGBA::from_rgb_slice(&[bgr[2], bgr[1], bgr[0]]): (c2, c1, c0, 255) as bytes 0..3
            uint32_t* o = rgba + (size_t)y * w + x;
            uint32_t e[PX];
#pragma unroll
            for (int k = 0; k < PX; k++) e[k] = __byte_perm(pixel(k), 0xFFu, 0x4012);
            if (VEC && PX % 4 == 0 && rgba_vec && nv == PX) {
#pragma unroll
                for (int q = 0; q < PX / 4; q++) reinterpret_cast<uint4*>(o)[q] = make_uint4(e[4 * q], e[4 * q + 1], e[4 * q + 2], e[4 * q + 3]);
            } else {
#pragma unroll
                for (int k = 0; k < PX; k++)
                    if (k < nv) o[k] = e[k];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ K7b
// imgproc::resize(INTER_LINEAR) of the 8-bit frame when "Process Fullres" is off (cv-decoder/src/lib.rs:127-135):
// OpenCV's fixed-point bilinear path (INTER_RESIZE_COEF_BITS = 11), see oracle/cv_front.c for the derivation.
// The per-column / per-row taps are recomputed by every thread (two f64 operations) instead of being tabulated.
__device__ __forceinline__ void resize_tap(int d, double scale, int src, int& s, int& a0, int& a1)
{
    float f = (float)__dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5);
    s = (int)floorf(f);
    f = __fsub_rn(f, (float)s);
    if (s < 0) { f = 0.0f; s = 0; }
    if (s >= src - 1) { f = 0.0f; s = src - 1; }
    a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.0f, f), 2048.0f));   // cvRound: half to even
    a1 = __float2int_rn(__fmul_rn(f, 2048.0f));
}

template <int CH>
__global__ void __launch_bounds__(256) frame_resize_kernel(const uint8_t* __restrict__ src, int sw, int sh, int stride,
                                                           uint8_t* __restrict__ dst, int dw, int dh, int dst_stride,
                                                           double scale_x, double scale_y)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    int x0, a0, a1, y0, b0, b1;
    resize_tap(x, scale_x, sw, x0, a0, a1);
    resize_tap(y, scale_y, sh, y0, b0, b1);
    const int x1 = min(x0 + 1, sw - 1), y1 = min(y0 + 1, sh - 1);
    const uint8_t* r0 = src + (size_t)y0 * stride;
    const uint8_t* r1 = src + (size_t)y1 * stride;
    uint8_t* o = dst + (size_t)y * dst_stride + (size_t)x * CH;
#pragma unroll
    for (int c = 0; c < CH; c++) {
        const int h0 = (int)__ldg(r0 + x0 * CH + c) * a0 + (int)__ldg(r0 + x1 * CH + c) * a1;
        const int h1 = (int)__ldg(r1 + x0 * CH + c) * a0 + (int)__ldg(r1 + x1 * CH + c) * a1;
        const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
        o[c] = (uint8_t)min(max(v, 0), 255);
    }
}

// ------------------------------------------------------------------------------------------ K8
// One CTA = a 256 x 32 tile of the mask.  (1) the gray tile with a 7-pixel apron (2 for the 5x5 Sobel + 5 for
// the 11x11 dilation) is staged in shared memory in 16-byte chunks: chunks inside the frame by asynchronous
// copies (row reflection only changes the source row), chunks crossing the left / right frame border by byte
// gathers with REFLECT_101; (2) one warp
// per 32 columns walks down the rows: horizontal derivative from 4 shared bytes, vertical derivative over a
// 5-row register window, `> 20` and a ballot turn 32 pixels into one word of the threshold bit-plane;
// (3) tiles at the frame border copy the bits of out-of-frame positions from their reflections (the Sobel of a
// reflected position is NOT the reflection of the Sobel: the mixed derivative changes sign, so bits are
// copied, not recomputed); (4) the dilation is bit-parallel: OR the rows that share a half-width, one
// log-step run-OR per half-width on a 64-bit window, 32 output pixels per thread.
constexpr int CM_TW = 256, CM_TH = 32, CM_NT = 288;
constexpr int CM_GW = CM_TW + 32;   // gray columns t0-16 .. t0+TW+15 (apron 7, widened to whole 16-byte chunks)
constexpr int CM_GH = CM_TH + 14;   // gray rows ty0-7 .. ty0+TH+6
constexpr int CM_BW = 9;            // words per threshold row: bit b <-> column t0-8+b (columns t0-5 .. t0+TW+4 are read)
constexpr int CM_BH = CM_TH + 10;   // threshold rows ty0-5 .. ty0+TH+4
constexpr int CM_QG = 8 * CM_BW;    // 4-pixel column groups per threshold row (72)
constexpr int CM_QN = (8 + CM_TW + 5 + 3) / 4;   // groups that hold a column some output pixel reads (68)
constexpr int CM_SEG = 11;          // threshold rows per row segment
static_assert(CM_NT == 4 * CM_QG && 4 * CM_SEG >= CM_BH, "thread = (column group, one of 4 row segments)");
static_assert(2 + CM_QN - 1 + 1 < CM_GW / 4, "the last group's right neighbour word lies inside the gray tile");

__global__ void __launch_bounds__(CM_NT) contrast_mask_kernel(const uint8_t* __restrict__ gray, int w, int h, int stride,
                                                              uint8_t* __restrict__ mask, int mask_stride)
{
    __shared__ __align__(16) uint8_t G[CM_GH][CM_GW];
    __shared__ uint32_t T[CM_BH][CM_BW];
    const int tid = threadIdx.x;
    const int t0 = blockIdx.x * CM_TW, ty0 = blockIdx.y * CM_TH;
    const int gx0 = t0 - 16, gy0 = ty0 - 7;

    // (1) gray tile: whole 16-byte chunks inside the frame by asynchronous copies (a reflected ROW only changes the
    //     source row); the bytes of chunks that cross the left / right frame border are gathered one per thread
    {
        constexpr int CPR = CM_GW / 16;   // chunks per row
        const bool aligned = ((reinterpret_cast<uintptr_t>(gray) | (uintptr_t)stride) & 15u) == 0;
        const bool rows_inside = gy0 >= 0 && gy0 + CM_GH <= h;
        // chunk k is copied whole iff it lies inside the frame: k in [k_lo, k_hi)
        const int k_lo = aligned ? max(0, (-gx0 + 15) >> 4) : CPR;
        const int k_hi = aligned ? min(CPR, (w - gx0) >> 4) : 0;
        for (int i = tid; i < CM_GH * CPR; i += CM_NT) {
            const int r = i / CPR, k = i - r * CPR;
            if (k >= k_lo && k < k_hi) {
                const int gy = rows_inside ? gy0 + r : reflect101(gy0 + r, h);
                async_copy16(&G[r][16 * k], gray + (size_t)gy * stride + gx0 + 16 * k);
            }
        }
        async_copy_commit();
        // Bytes the copies left out: in-frame pixels of partially covered chunks, and the two reflected columns on
        // either side of the frame that the 5-tap row filter of an in-frame pixel reads.  (Threshold bits of
        // out-of-frame columns are copied in step 3, never computed, so nothing further out is needed.)
        const int c_lo = max(0, -2 - gx0), c_hi = min(CM_GW, w + 2 - gx0);
        const int a_hi = k_hi > k_lo ? min(16 * k_lo, c_hi) : c_hi;        // segment A: [c_lo, a_hi)
        const int b_lo = k_hi > k_lo ? max(16 * k_hi, c_lo) : c_hi;        // segment B: [b_lo, c_hi)
        const int lane = tid & 31;
        for (int r = tid >> 5; r < CM_GH; r += CM_NT / 32) {
            const uint8_t* row = gray + (size_t)reflect101(gy0 + r, h) * stride;
            for (int c = c_lo + lane; c < a_hi; c += 32) G[r][c] = __ldg(row + reflect101(gx0 + c, w));
            for (int c = b_lo + lane; c < c_hi; c += 32) G[r][c] = __ldg(row + reflect101(gx0 + c, w));
        }
        async_copy_wait<0>();
    }
    __syncthreads();

    // (2) Sobel(dx=1, dy=1, ksize 5) = [-1,-2,0,2,1]^T x [-1,-2,0,2,1], then `> 20`, four pixels per thread:
    //     thread = (column group q of 4 adjacent pixels, row segment): 72 groups x 4 segments of 11 threshold rows.
    //     Horizontal pass: the five taps of a pixel are one dp4a over the four bytes starting two to its left
    //     (weights -1,-2,0,2) plus one dp4a that picks the fifth byte; +1024 keeps it positive.  Two such values
    //     share a register as 16-bit lanes, so the vertical pass is three integer operations per pixel PAIR
    //     (minuend and subtrahend are sums of positives: no borrow crosses the lanes), and the constant puts
    //     `s >= 21` into bit 15 of each lane.  The four bits of a thread and the nibbles of eight adjacent
    //     threads (xor-butterfly) make one word of the threshold bit-plane.
    {
        const uint32_t* Gw = reinterpret_cast<const uint32_t*>(&G[0][0]);
        constexpr int GWW = CM_GW / 4;                  // words per gray row
        const int q = tid % CM_QG, seg = tid / CM_QG;
        // word of the group's first pixel (column t0-8+4q) in a gray row.  Groups past the last column that any
        // output pixel reads (q >= CM_QN) are clamped into the tile: their bits (>= 272) are never looked at.
        const int wq = 2 + min(q, CM_QN - 1);
        const int X0 = t0 - 8 + 4 * q;
        uint32_t colmask = 0;   // out-of-frame columns must stay 0: step (3) ORs their reflections in
#pragma unroll
        for (int j = 0; j < 4; j++) colmask |= (uint32_t)(X0 + j >= 0 && X0 + j < w) << j;
        const int tr0 = CM_SEG * seg;
        const uint32_t* gcol = Gw + wq;
        const bool writer = (tid & 7) == 0;
        uint32_t p[5][2];                               // packed horizontal derivatives of gray rows gr-4 .. gr
#pragma unroll
        for (int it = 0; it < CM_SEG + 4; it++) {
            const int gr = min(tr0 + it, CM_GH - 1);    // gray row of this iteration (rows past the tile only feed unused threshold rows)
#pragma unroll
            for (int k = 0; k < 4; k++) { p[k][0] = p[k + 1][0]; p[k][1] = p[k + 1][1]; }
            const uint32_t* gw_ = gcol + gr * GWW;
            const uint32_t W0 = gw_[-1], W1 = gw_[0], W2 = gw_[1];
            const int h0 = dp4a_us(W1, 0x00010000, dp4a_us(__funnelshift_r(W0, W1, 16), 0x0200FEFF, 1024));
            const int h1 = dp4a_us(W1, 0x01000000, dp4a_us(__funnelshift_r(W0, W1, 24), 0x0200FEFF, 1024));
            const int h2 = dp4a_us(W2, 0x00000001, dp4a_us(W1, 0x0200FEFF, 1024));
            const int h3 = dp4a_us(W2, 0x00000100, dp4a_us(__funnelshift_r(W1, W2, 8), 0x0200FEFF, 1024));
            p[4][0] = __byte_perm((uint32_t)h0, (uint32_t)h1, 0x5410);
            p[4][1] = __byte_perm((uint32_t)h2, (uint32_t)h3, 0x5410);
            if (it >= 4) {                              // centre = gray row gr-2 = threshold row tr (frame row ty0-5+tr)
                const int tr = tr0 + it - 4;
                // lane = s + 32747, s = h[gr] - h[gr-4] + 2 (h[gr-1] - h[gr-3]): bit 15 <=> s > 20.  Rows outside the
                // frame need no masking: step (4) reflects row indices and never reads them.
                const uint32_t slo = (p[4][0] + 2u * p[3][0] + 0x7FEB7FEBu) - (p[0][0] + 2u * p[1][0]);
                const uint32_t shi = (p[4][1] + 2u * p[3][1] + 0x7FEB7FEBu) - (p[0][1] + 2u * p[1][1]);
                const uint32_t v = ((slo >> 15) & 0x10001u) | (((shi >> 15) & 0x10001u) << 2);
                uint32_t word = ((v | (v >> 15)) & colmask) << (4 * (tid & 7));   // OR over the 8 lanes of the group
                word |= __shfl_xor_sync(0xffffffffu, word, 1);
                word |= __shfl_xor_sync(0xffffffffu, word, 2);
                word |= __shfl_xor_sync(0xffffffffu, word, 4);
                if (writer && tr < CM_BH) T[tr][q >> 3] = word;
            }
        }
    }
    __syncthreads();

    // (3) dilate's BORDER_REFLECT_101: rows are reflected by index in (4); out-of-frame COLUMNS take the bit of
    //     their reflection here (at most five on either side of the frame)
    if (t0 == 0 || w < t0 + CM_TW + 5) {
        for (int i = tid; i < CM_BH * 10; i += CM_NT) {
            const int br = i / 10, k = i - br * 10;
            const int X = k < 5 ? k - 5 : w + (k - 5);
            const int b = X - (t0 - 8), qb = reflect101(X, w) - (t0 - 8);
            if (b < 0 || b >= 32 * CM_BW || qb < 0 || qb >= 32 * CM_BW) continue;   // not in this tile / read by no pixel of it
            if ((T[br][qb >> 5] >> (qb & 31)) & 1u) atomicOr(&T[br][b >> 5], 1u << (b & 31));
        }
        __syncthreads();
    }

    // (4) dilation by the 11x11 ellipse: row half-widths 0,3,4,5,5,5,5,5,4,3,0
    if (tid < (CM_TW / 32) * CM_TH) {
        const int orow = tid >> 3, ow = tid & 7;
        const int y = ty0 + orow, x0 = t0 + 32 * ow;
        if (y < h && x0 < w) {
            // 64-bit window of the threshold row of frame row y+dy (reflected into the frame): bit i <-> column x0-8+i
            const bool rows_inside = ty0 - 5 >= 0 && ty0 + CM_TH + 5 <= h;   // tile-uniform: no row of the window is reflected
            auto win = [&](int dy) -> unsigned long long {
                int py = y + dy;
                if (!rows_inside && (py < 0 || py >= h)) py = reflect101(py, h);
                const uint32_t* t = T[py - (ty0 - 5)];
                return (unsigned long long)t[ow] | ((unsigned long long)t[ow + 1] << 32);
            };
            const unsigned long long v5 = win(-2) | win(-1) | win(0) | win(1) | win(2);
            const unsigned long long v4 = win(-3) | win(3), v3 = win(-4) | win(4), v0 = win(-5) | win(5);
            // run-OR of n = 2k+1 consecutive bits, then >> (8-k) centres it on the output pixel
            const unsigned long long a2 = v5 | (v5 >> 1), a4 = a2 | (a2 >> 2), a8 = a4 | (a4 >> 4);
            uint32_t m = (uint32_t)((a8 | (a4 >> 7)) >> 3);                    // k = 5: 11 bits
            const unsigned long long b2 = v4 | (v4 >> 1), b4 = b2 | (b2 >> 2), b8 = b4 | (b4 >> 4);
            m |= (uint32_t)((b8 | (v4 >> 8)) >> 4);                            // k = 4: 9 bits
            const unsigned long long c2 = v3 | (v3 >> 1), c4 = c2 | (c2 >> 2);
            m |= (uint32_t)((c4 | (c4 >> 3)) >> 5);                            // k = 3: 7 bits
            m |= (uint32_t)(v0 >> 8);                                          // k = 0
            uint8_t* o = mask + (size_t)y * mask_stride + x0;
            if (x0 + 32 <= w && ((reinterpret_cast<uintptr_t>(mask) | (uintptr_t)mask_stride) & 15u) == 0) {
                uint32_t e[8];
#pragma unroll
                for (int k = 0; k < 8; k++) e[k] = ((((m >> (4 * k)) & 15u) * 0x00204081u) & 0x01010101u) * 0xFFu;
                reinterpret_cast<uint4*>(o)[0] = make_uint4(e[0], e[1], e[2], e[3]);
                reinterpret_cast<uint4*>(o)[1] = make_uint4(e[4], e[5], e[6], e[7]);
            } else {
                for (int j = 0; j < 32 && x0 + j < w; j++) o[j] = ((m >> j) & 1u) ? 255 : 0;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ K9
// MotionFieldDensifier::add_vector_weighted's cell lookup (motion_field.rs:170-176) for the positions
// cv-decoder generates, pos = (x + 0.5) * (1/W) (cv-decoder/src/lib.rs:262-264).  Every factor is a
// monotone f32 operation, so the pixels of one cell column are a contiguous run of x (likewise rows): the
// cell grid is separable and K9 can fold whole pixel rectangles.  The host checks that no position reaches
// 0 or 1, which keeps nalgebra's all-components clamp the identity.
__device__ __forceinline__ int cell_coord(int x, float inv, float gm1)
{
    const float pos = __fmul_rn(__fadd_rn((float)x, 0.5f), inv);
    const float v = roundf(__fmul_rn(pos, gm1));
    return v > 0.0f ? (int)v : 0;
}

// First pixel coordinate whose cell index is >= c (len if there is none): binary search on the monotone cell_coord.
__device__ __forceinline__ int cell_start(int c, int len, float inv, float gm1)
{
    int lo = 0, hi = len;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cell_coord(mid, inv, gm1) >= c) hi = mid;
        else lo = mid + 1;
    }
    return lo;
}

constexpr float F32_EPSILON = 1.1920928955078125e-07f;
constexpr int FC_NC = 32;         // cells per CTA = one warp: adjacent cells of one cell row, one lane each
constexpr int FC_CAP = 2048;      // most pixels of one row in a step (wider spans are walked in chunks)
constexpr int FC_STEP_PX = 2048;  // rows are grouped into steps of about this many pixels
constexpr int FC_MAX_NST = 8;     // most staging buffers in the ring
constexpr int FC_UNR = 8;         // pixels per fold group
constexpr int FC_SMEM_PAD = 128;  // the fold may load up to FC_UNR - 1 pixels past the end of the last buffer
constexpr int FC_SMEM_BYTES = 72 * 1024;   // dynamic shared memory per CTA (3 CTAs per SM): the ring takes as many
                                           // buffers of the step's size as fit, at least 3 (FC_CAP * 8 + mask < 24 KB)

struct __align__(16) CellRec { float mx, my; uint32_t touched, pad; };

// One CTA = two warps = FC_NC adjacent cells of one cell row.  The pixel rows of those cells stream through a ring of
// 3-8 shared-memory buffers (as many steps of ~2048 pixels as fit into 72 KB).  Warp 1 (the stager) issues the copies
// of step s+nst-1 (ASYNC: 16-byte asynchronous copies, two flow pixels / sixteen mask bytes each, starting at the
// aligned address below the first pixel) while warp 0 folds step s; one __syncthreads per step hands a landed buffer
// over and takes the folded one back.  The fold is a single warp's dependent instruction stream (one instruction every
// ~4 cycles, see profiles/r1_cv_front_ncu.md), so everything that is not an add of the reference's sum is kept off it.
// Lane t folds the pixels of cell t in raster order — rows top to bottom, columns left to right — which is the
// order in which the reference's loop reaches that cell, so sums and counts are bit-identical:
//   counts += 1.0 (from f32::EPSILON), sum = motion * 1.0 + sum, motion = flow .* (1/W, 1/H).
template <bool ASYNC>
__global__ void __launch_bounds__(64) flow_cells_kernel(const float* __restrict__ flow, long long flow_stride,
                                                        const uint8_t* __restrict__ mask, long long mask_stride,
                                                        int w, int h, int gw, int gh, CellRec* __restrict__ cells,
                                                        uint32_t* __restrict__ colcount)
{
    OFPSB_DYN_SMEM(dyn);
    const int lane = threadIdx.x & 31;
    const bool stager = threadIdx.x >= 32;   // warp 1 keeps the copies in flight, warp 0 folds
    const int cy = blockIdx.y, c0 = blockIdx.x * FC_NC;
    const float nx = __fdiv_rn(1.0f, (float)w), ny = __fdiv_rn(1.0f, (float)h);
    const float gxm1 = (float)(gw - 1), gym1 = (float)(gh - 1);
    const int cell = c0 + lane;
    const bool owner = !stager && cell < gw;
    const int xa = cell_start(min(cell, gw), w, nx, gxm1), xb = cell_start(min(cell + 1, gw), w, nx, gxm1);
    const int px0 = __shfl_sync(0xffffffffu, xa, 0), px1 = __shfl_sync(0xffffffffu, xb, 31);
    const int y0 = cell_start(cy, h, ny, gym1), y1 = cell_start(cy + 1, h, ny, gym1);
    float sx = 0.0f, sy = 0.0f, cnt = F32_EPSILON;
    const int span = px1 - px0, rows = y1 - y0;
    if (span > 0 && rows > 0) {
        const int cw = min(span, FC_CAP - 4);          // pixels per row of one step
        const int pitchf = (cw + 3) & ~1;              // row pitch of the flow buffer (pixels; even, >= cw + 2)
        const int pitchm = (cw + 30) & ~15;            // row pitch of the mask buffer (bytes; multiple of 16, >= cw + 15)
        const int rg = max(1, FC_STEP_PX / pitchf);    // rows per step; > 1 only when cw == span
        const int fbytes = rg * pitchf * 8, sbytes = fbytes + rg * pitchm;   // one staging buffer: flow rows, then mask rows
        const int nst = min(FC_MAX_NST, FC_SMEM_BYTES / sbytes);             // ring depth: nst - 1 steps in flight
        const int ncs = (span + cw - 1) / cw, nrs = (rows + rg - 1) / rg, nsteps = ncs * nrs;
        // step s = (row group ri, column chunk ci), buffer s % nst — kept incrementally for the staging side (p_*) and
        // the folding side (c_*): no divisions on the warp's critical path
        int p_s = 0, p_ri = 0, p_ci = 0, p_buf = 0;
        auto stage = [&]() {   // issue the copies of the next step into its buffer, commit one group
            if (p_s < nsteps) {
                const int r0 = y0 + p_ri * rg, nr = min(rg, y1 - r0);
                const int cx0 = px0 + p_ci * cw, nc = min(cw, px1 - cx0);
                uint8_t* buf = dyn + p_buf * sbytes;
                if (ASYNC) {
                    const int ax0 = cx0 & ~1, nf = (cx0 + nc - ax0 + 1) >> 1;
                    const char* g = reinterpret_cast<const char*>(flow + (long long)r0 * flow_stride + 2ll * ax0) + 16 * lane;
                    smem_addr_t d = smem_addr(buf) + 16 * lane;
                    for (int r = 0; r < nr; r++, g += flow_stride * 4, d += pitchf * 8) {
                        const char* gk = g;
                        smem_addr_t dk = d;
                        for (int k = lane; k < nf; k += 32, gk += 512, dk += 512) async_copy16_to(dk, gk);
                    }
                    if (mask) {
                        const int am0 = cx0 & ~15, nm = (cx0 + nc - am0 + 15) >> 4;
                        const uint8_t* m = mask + (long long)r0 * mask_stride + am0 + 16 * lane;
                        smem_addr_t dm = smem_addr(buf + fbytes) + 16 * lane;
                        for (int r = 0; r < nr; r++, m += mask_stride, dm += pitchm) {
                            const uint8_t* mk = m;
                            smem_addr_t dk = dm;
                            for (int k = lane; k < nm; k += 32, mk += 512, dk += 512) async_copy16_to(dk, mk);
                        }
                    }
                } else {
                    float2* sf = reinterpret_cast<float2*>(buf);
                    uint8_t* sm = buf + fbytes;
                    for (int r = 0; r < nr; r++) {
                        const float* frow = flow + (long long)(r0 + r) * flow_stride + 2ll * cx0;
                        for (int c = lane; c < nc; c += 32) sf[r * pitchf + c] = make_float2(__ldg(frow + 2 * c), __ldg(frow + 2 * c + 1));
                        if (mask) {
                            const uint8_t* mrow = mask + (long long)(r0 + r) * mask_stride + cx0;
                            for (int c = lane; c < nc; c += 32) sm[r * pitchm + c] = __ldg(mrow + c);
                        }
                    }
                }
                p_s++;
                if (++p_ci == ncs) { p_ci = 0; p_ri++; }
                if (++p_buf == nst) p_buf = 0;
            }
            async_copy_commit();   // one group per call (possibly empty) keeps wait_group's count uniform
        };
        // Warp 1 runs nst - 1 steps ahead of warp 0; one __syncthreads per step hands a landed buffer over and takes
        // the folded one back.  At the barrier that ends iteration s the copies of step s + 1 have landed
        // (wait_group leaves the nst - 2 younger groups in flight) and step s has been folded, so its buffer —
        // (s + nst) % nst — is the one the stager fills next.
        auto wait_landed = [&]() {
            switch (nst) {
                case 3: async_copy_wait<1>(); break;
                case 4: async_copy_wait<2>(); break;
                case 5: async_copy_wait<3>(); break;
                case 6: async_copy_wait<4>(); break;
                case 7: async_copy_wait<5>(); break;
                default: async_copy_wait<6>(); break;
            }
        };
        if (stager) {
            for (int s = 0; s < nst - 1; s++) stage();
            wait_landed();   // step 0
        }
        __syncthreads();
        int c_ri = 0, c_ci = 0, c_buf = 0;
        for (int s = 0; s < nsteps; s++) {
            if (stager) {
                stage();         // step s + nst - 1, into the buffer folded in the previous iteration
                wait_landed();   // step s + 1
            }
            if (owner) {
                const int r0 = y0 + c_ri * rg, nr = min(rg, y1 - r0);
                const int cx0 = px0 + c_ci * cw, nc = min(cw, px1 - cx0);
                const int a = max(xa, cx0) - cx0, b = min(xb, cx0 + nc) - cx0;
                const float2* sf = reinterpret_cast<const float2*>(dyn + c_buf * sbytes) + (ASYNC ? (cx0 & 1) : 0);
                const uint8_t* sm = dyn + c_buf * sbytes + fbytes + (ASYNC ? (cx0 & 15) : 0);
                for (int r = 0; r < nr; r++) {
                    const float2* fr = sf + r * pitchf;
                    const uint8_t* mr = sm + r * pitchm;
                    // Groups of FC_UNR pixels: all loads of a group are issued before the first add, so the warp waits
                    // for shared memory once per group; slots past the cell's last pixel are loaded (the buffers are
                    // padded) but predicated off, which also removes any remainder loop.
                    auto group = [&](int c, auto tail, auto slots) {   // tail: the last, partial group of the row checks c + j < b
                        constexpr bool TAIL = decltype(tail)::value;
                        constexpr int N = decltype(slots)::value;   // 8, or 4 for a short tail
                        float2 f[N];
                        uint32_t m[N];
#pragma unroll
                        for (int j = 0; j < N; j++) {
                            f[j] = fr[c + j];
                            m[j] = mask ? (uint32_t)mr[c + j] : 1u;   // `*mask < 0.1` -> skip (cv-decoder:258)
                        }
#pragma unroll
                        for (int j = 0; j < N; j++) OFPSB_KEEP_LOADED(f[j].x, f[j].y);
#pragma unroll
                        for (int j = 0; j < N; j++)
                            if ((!TAIL || c + j < b) && m[j] != 0u) {
                                cnt = __fadd_rn(cnt, 1.0f);
                                sx = __fadd_rn(__fmul_rn(f[j].x, nx), sx);
                                sy = __fadd_rn(__fmul_rn(f[j].y, ny), sy);
                            }
                    };
                    int c = a;
                    for (; c + FC_UNR <= b; c += FC_UNR) group(c, std::false_type{}, std::integral_constant<int, FC_UNR>{});
                    if (c + FC_UNR / 2 < b) group(c, std::true_type{}, std::integral_constant<int, FC_UNR>{});
                    else if (c < b) group(c, std::true_type{}, std::integral_constant<int, FC_UNR / 2>{});
                }
            }
            if (++c_ci == ncs) { c_ci = 0; c_ri++; }
            if (++c_buf == nst) c_buf = 0;
            __syncthreads();   // step s + 1 is visible to the folding warp, buffer s % nst is free for step s + nst
        }
    }
    if (owner) {
        CellRec rec;
        rec.mx = __fdiv_rn(sx, cnt);   // MotionField::from(densifier): sum ./ counts
        rec.my = __fdiv_rn(sy, cnt);
        rec.touched = cnt > 0.5f ? 1u : 0u;   // counts start at f32::EPSILON and grow by 1.0 per vector
        rec.pad = 0u;
        cells[(size_t)cell * gh + cy] = rec;   // column-major: the order flow_emit_cells_kernel walks
        if (rec.touched) atomicAdd(&colcount[cell], 1u);   // touched cells per column (integer: order-free)
    }
}

// Block-wide exclusive scan helper for the ordered compactions (1024 threads).
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* warp_tot, uint32_t* total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    __syncthreads();   // warp_tot may still be read from the previous call
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
    for (int k = 0; k < nw; k++) {
        const uint32_t t = warp_tot[k];
        if (k < wid) base += t;
        tot += t;
    }
    *total = tot;
    return base + inc - v;
}

// Touched cells in (x, y) lexicographic order (BTreeSet<(usize, usize)>, cv-decoder:243, 276) -> entries:
// pos = (x + 0.5, y + 0.5) .* (1/gw, 1/gh), motion = cell mean.  One warp per cell column: the entries of column x
// start at the number of touched cells in the columns before it (colcount, summed per CTA), and inside the column a
// ballot ranks the touched cells, so a warp writes its entries to consecutive addresses.
constexpr int FE_WARPS = 8;
__global__ void __launch_bounds__(32 * FE_WARPS) flow_emit_cells_kernel(const CellRec* __restrict__ cells,
                                                                        const uint32_t* __restrict__ colcount, int gw, int gh,
                                                                        ofps_mv* __restrict__ out, unsigned long long cap,
                                                                        unsigned long long* __restrict__ n_out)
{
    __shared__ unsigned long long s_part[FE_WARPS];
    __shared__ unsigned long long s_base;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int xfirst = blockIdx.x * FE_WARPS;
    // touched cells in the columns before this CTA's first one (the last CTA also totals everything for *n_out)
    const int upto = blockIdx.x == gridDim.x - 1 ? gw : xfirst;
    unsigned long long before = 0, all = 0;
    for (int i = threadIdx.x; i < upto; i += blockDim.x) {
        const uint32_t c = __ldg(colcount + i);
        all += c;
        if (i < xfirst) before += c;
    }
    for (int pass = 0; pass < 2; pass++) {   // block sum of `before`, then of `all`
        unsigned long long v = pass == 0 ? before : all;
        uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
        for (int d = 16; d > 0; d >>= 1) {
            const uint32_t lo2 = __shfl_xor_sync(0xffffffffu, lo, d), hi2 = __shfl_xor_sync(0xffffffffu, hi, d);
            v = (((unsigned long long)hi << 32) | lo) + (((unsigned long long)hi2 << 32) | lo2);
            lo = (uint32_t)v;
            hi = (uint32_t)(v >> 32);
        }
        __syncthreads();   // s_part is reused by the second pass
        if (lane == 0) s_part[wid] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long t = 0;
            for (int k = 0; k < FE_WARPS; k++) t += s_part[k];
            if (pass == 0) s_base = t;
            else if (blockIdx.x == gridDim.x - 1) *n_out = t;
        }
    }
    __syncthreads();
    const int x = xfirst + wid;
    if (x >= gw) return;
    unsigned long long pos = s_base;
    for (int k = 0; k < wid; k++) pos += __ldg(colcount + xfirst + k);
    const float gx = __fdiv_rn(1.0f, (float)gw), gy = __fdiv_rn(1.0f, (float)gh);
    const uint4* col = reinterpret_cast<const uint4*>(cells) + (size_t)x * gh;   // records are column-major
    for (int y0 = 0; y0 < gh; y0 += 32) {
        const int y = y0 + lane;
        uint4 rec = make_uint4(0u, 0u, 0u, 0u);
        if (y < gh) rec = __ldg(col + y);
        const unsigned m = __ballot_sync(0xffffffffu, (rec.z & 1u) != 0);
        if (rec.z & 1u) {
            const unsigned long long p = pos + __popc(m & ((1u << lane) - 1u));
            if (p < cap) {
                float4 e;
                e.x = __fmul_rn(__fadd_rn((float)x, 0.5f), gx);
                e.y = __fmul_rn(__fadd_rn((float)y, 0.5f), gy);
                e.z = __uint_as_float(rec.x);
                e.w = __uint_as_float(rec.y);
                *reinterpret_cast<float4*>(out + p) = e;
            }
        }
        pos += __popc(m);
    }
}

// Per-pixel variant (process_fullres = false): every kept pixel becomes an entry, raster order.
constexpr int FP_NT = 256;
constexpr int FP_TILE = 4096;   // pixels per CTA

__global__ void __launch_bounds__(FP_NT) flow_pixels_count_kernel(const uint8_t* __restrict__ mask, long long mask_stride,
                                                                  int w, long long npix, uint32_t* __restrict__ tile_count)
{
    __shared__ uint32_t s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const long long p0 = (long long)blockIdx.x * FP_TILE;
    uint32_t c = 0;
    for (int i = threadIdx.x; i < FP_TILE; i += FP_NT) {
        const long long p = p0 + i;
        if (p < npix) {
            const long long y = p / w;
            c += __ldg(mask + y * mask_stride + (p - y * w)) != 0;
        }
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) tile_count[blockIdx.x] = s_cnt;
}

// Exclusive scan of the tile counts in place (one CTA), total -> *n_out.
__global__ void __launch_bounds__(1024) tile_scan_kernel(uint32_t* __restrict__ tile_count, long long ntiles,
                                                         unsigned long long* __restrict__ tile_base,
                                                         unsigned long long* __restrict__ n_out)
{
    __shared__ uint32_t warp_tot[32];
    unsigned long long base = 0;
    for (long long k0 = 0; k0 < ntiles; k0 += blockDim.x) {
        const long long k = k0 + threadIdx.x;
        const uint32_t v = k < ntiles ? tile_count[k] : 0u;
        uint32_t tot;
        const uint32_t ex = block_excl_scan(v, warp_tot, &tot);
        if (k < ntiles) tile_base[k] = base + ex;
        base += tot;
    }
    if (threadIdx.x == 0) *n_out = base;
}

__global__ void __launch_bounds__(FP_NT) flow_pixels_emit_kernel(const float* __restrict__ flow, long long flow_stride,
                                                                 const uint8_t* __restrict__ mask, long long mask_stride,
                                                                 int w, int h, long long npix,
                                                                 const unsigned long long* __restrict__ tile_base,
                                                                 ofps_mv* __restrict__ out, unsigned long long cap)
{
    __shared__ uint32_t warp_tot[32];
    const float nx = __fdiv_rn(1.0f, (float)w), ny = __fdiv_rn(1.0f, (float)h);
    const long long p0 = (long long)blockIdx.x * FP_TILE;
    unsigned long long base = mask ? tile_base[blockIdx.x] : (unsigned long long)p0;
    for (int i0 = 0; i0 < FP_TILE; i0 += FP_NT) {
        const long long p = p0 + i0 + threadIdx.x;
        int x = 0, y = 0;
        uint32_t keep = 0;
        if (p < npix) {
            y = (int)(p / w);
            x = (int)(p - (long long)y * w);
            keep = mask ? (__ldg(mask + (long long)y * mask_stride + x) != 0) : 1u;
        }
        uint32_t rank, tot;
        if (mask) {
            rank = block_excl_scan(keep, warp_tot, &tot);
        } else {
            rank = threadIdx.x;
            tot = FP_NT;
        }
        if (keep && base + rank < cap) {
            const float* f = flow + (long long)y * flow_stride + 2ll * x;
            ofps_mv e;
            e.px = __fmul_rn(__fadd_rn((float)x, 0.5f), nx);
            e.py = __fmul_rn(__fadd_rn((float)y, 0.5f), ny);
            e.mx = __fmul_rn(__ldg(f), nx);
            e.my = __fmul_rn(__ldg(f + 1), ny);
            out[base + rank] = e;
        }
        base += tot;
    }
}

}  // namespace

// ------------------------------------------------------------------------------------ launchers
int launch_frame_convert(const uint8_t* d_src, int w, int h, int stride, int channels, int rgb_order, uint8_t* d_gray,
                         int gray_stride, uint8_t* d_rgba, cudaStream_t stream, uint64_t* launches)
{
    if (w <= 0 || h <= 0 || (channels != 3 && channels != 4) || stride < w * channels || (d_gray && gray_stride < w)) {
        set_error("frame_convert: invalid geometry (w=%d h=%d stride=%d channels=%d)", w, h, stride, channels);
        return OFPSB_E_INVALID;
    }
    if (!d_gray && !d_rgba) return OFPSB_OK;
    const bool vec = ((reinterpret_cast<uintptr_t>(d_src) | (uintptr_t)stride) & 15u) == 0 &&
                     (!d_gray || ((reinterpret_cast<uintptr_t>(d_gray) | (uintptr_t)gray_stride) & 15u) == 0);
    const int rgba_vec = d_rgba && (reinterpret_cast<uintptr_t>(d_rgba) & 15u) == 0 && (w & 3) == 0;
    uint32_t* rgba = reinterpret_cast<uint32_t*>(d_rgba);
    const int px = vec ? 16 : 4;
    const dim3 grid((unsigned)(((w + px - 1) / px + 127) / 128), (unsigned)(h < 65535 ? h : 65535));   // small CTAs: even waves
    if (channels == 3) {
        if (vec) OFPSB_LAUNCH((frame_convert_kernel<3, 16, true>), grid, 128, stream, d_src, w, h, stride, rgb_order, d_gray, gray_stride, rgba, rgba_vec);
        else OFPSB_LAUNCH((frame_convert_kernel<3, 4, false>), grid, 128, stream, d_src, w, h, stride, rgb_order, d_gray, gray_stride, rgba, rgba_vec);
    } else {
        if (vec) OFPSB_LAUNCH((frame_convert_kernel<4, 16, true>), grid, 128, stream, d_src, w, h, stride, rgb_order, d_gray, gray_stride, rgba, rgba_vec);
        else OFPSB_LAUNCH((frame_convert_kernel<4, 4, false>), grid, 128, stream, d_src, w, h, stride, rgb_order, d_gray, gray_stride, rgba, rgba_vec);
    }
    OFPSB_CUDA_TRY(cudaGetLastError());
    if (launches) *launches += 1;
    return OFPSB_OK;
}

int launch_frame_resize(const uint8_t* d_src, int sw, int sh, int stride, int channels, uint8_t* d_dst, int dw, int dh,
                        int dst_stride, cudaStream_t stream, uint64_t* launches)
{
    if (sw <= 0 || sh <= 0 || dw <= 0 || dh <= 0 || (channels != 3 && channels != 4) || stride < sw * channels ||
        dst_stride < dw * channels) {
        set_error("frame_resize: invalid geometry (%dx%d -> %dx%d, channels=%d)", sw, sh, dw, dh, channels);
        return OFPSB_E_INVALID;
    }
    if (dw > sw || dh > sh) {   // OpenCV takes a different path when enlarging; only reductions are pinned (and used)
        set_error("frame_resize: %dx%d -> %dx%d enlarges the frame; cv-decoder only reduces", sw, sh, dw, dh);
        return OFPSB_E_INVALID;
    }
    const dim3 block(32, 8), grid((unsigned)((dw + 31) / 32), (unsigned)((dh + 7) / 8));
    const double sx = (double)sw / (double)dw, sy = (double)sh / (double)dh;
    if (channels == 3) OFPSB_LAUNCH(frame_resize_kernel<3>, grid, block, stream, d_src, sw, sh, stride, d_dst, dw, dh, dst_stride, sx, sy);
    else OFPSB_LAUNCH(frame_resize_kernel<4>, grid, block, stream, d_src, sw, sh, stride, d_dst, dw, dh, dst_stride, sx, sy);
    OFPSB_CUDA_TRY(cudaGetLastError());
    if (launches) *launches += 1;
    return OFPSB_OK;
}

int launch_contrast_mask(const uint8_t* d_gray, int w, int h, int stride, uint8_t* d_mask, int mask_stride,
                         cudaStream_t stream, uint64_t* launches)
{
    if (w <= 0 || h <= 0 || stride < w || mask_stride < w) {
        set_error("contrast_mask: invalid geometry (w=%d h=%d stride=%d mask_stride=%d)", w, h, stride, mask_stride);
        return OFPSB_E_INVALID;
    }
    const dim3 grid((unsigned)((w + CM_TW - 1) / CM_TW), (unsigned)((h + CM_TH - 1) / CM_TH));
    if (grid.y > 65535u) {
        set_error("contrast_mask: frame too tall (%d rows)", h);
        return OFPSB_E_INVALID;
    }
    OFPSB_LAUNCH(contrast_mask_kernel, grid, CM_NT, stream, d_gray, w, h, stride, d_mask, mask_stride);
    OFPSB_CUDA_TRY(cudaGetLastError());
    if (launches) *launches += 1;
    return OFPSB_OK;
}

// true when every generated position (x+0.5)/W stays strictly inside (0,1) in f32 (see cell_coord)
static bool positions_strictly_inside(int len)
{
    const float inv = 1.0f / (float)len;
    volatile float hi = ((float)(len - 1) + 0.5f) * inv, lo = 0.5f * inv;
    return lo > 0.0f && hi < 1.0f;
}

int launch_flow_entries(const float* d_flow, size_t flow_stride, const uint8_t* d_mask, size_t mask_stride, int w, int h,
                        size_t gw, size_t gh, ofps_mv* d_entries, size_t cap, unsigned long long* d_count,
                        FlowScratch& scratch, cudaStream_t stream, uint64_t* launches)
{
    if (w <= 0 || h <= 0 || flow_stride < 2 * (size_t)w || (d_mask && mask_stride < (size_t)w)) {
        set_error("flow_entries: invalid geometry (w=%d h=%d flow_stride=%zu mask_stride=%zu)", w, h, flow_stride, mask_stride);
        return OFPSB_E_INVALID;
    }
    if ((gw == 0) != (gh == 0) || gw > 32768 || gh > 32768 || w > (1 << 20) || h > (1 << 20)) {
        set_error("flow_entries: invalid motion-field size %zux%zu for a %dx%d frame", gw, gh, w, h);
        return OFPSB_E_INVALID;
    }
    if (!positions_strictly_inside(w) || !positions_strictly_inside(h)) {
        set_error("flow_entries: %dx%d frame: pixel-centre positions reach 0 or 1 in f32", w, h);
        return OFPSB_E_INVALID;
    }
    if (reinterpret_cast<uintptr_t>(d_entries) & 15u) {
        set_error("flow_entries: the entry buffer must be 16-byte aligned");
        return OFPSB_E_INVALID;
    }
    const long long npix = (long long)w * h;
    if (gw == 0) {   // per pixel
        const long long ntiles = (npix + FP_TILE - 1) / FP_TILE;
        if (d_mask) {
            if (int rc = scratch.tiles.reserve((size_t)ntiles * (4 + 8))) return rc;
            unsigned long long* tile_base = scratch.tiles.as<unsigned long long>();
            uint32_t* tile_count = reinterpret_cast<uint32_t*>(tile_base + ntiles);
            OFPSB_LAUNCH(flow_pixels_count_kernel, (unsigned)ntiles, FP_NT, stream, d_mask, (long long)mask_stride, w, npix, tile_count);
            OFPSB_LAUNCH(tile_scan_kernel, 1, 1024, stream, tile_count, ntiles, tile_base, d_count);
            OFPSB_LAUNCH(flow_pixels_emit_kernel, (unsigned)ntiles, FP_NT, stream, d_flow, (long long)flow_stride, d_mask,
                                                                           (long long)mask_stride, w, h, npix, tile_base,
                                                                           d_entries, cap);
            if (launches) *launches += 3;
        } else {
            const unsigned long long n = (unsigned long long)npix;
            OFPSB_CUDA_TRY(cudaMemcpyAsync(d_count, &n, sizeof(n), cudaMemcpyHostToDevice, stream));
            OFPSB_LAUNCH(flow_pixels_emit_kernel, (unsigned)ntiles, FP_NT, stream, d_flow, (long long)flow_stride, nullptr, 0, w, h,
                                                                           npix, nullptr, d_entries, cap);
            if (launches) *launches += 1;
        }
        OFPSB_CUDA_TRY(cudaGetLastError());
        return OFPSB_OK;
    }
    const int igw = (int)gw, igh = (int)gh;
    if (int rc = scratch.cells.reserve(gw * gh * sizeof(CellRec))) return rc;
    if (int rc = scratch.bounds.reserve(gw * sizeof(uint32_t))) return rc;
    CellRec* cells = scratch.cells.as<CellRec>();
    uint32_t* colcount = scratch.bounds.as<uint32_t>();
    OFPSB_CUDA_TRY(cudaMemsetAsync(colcount, 0, gw * sizeof(uint32_t), stream));
    const dim3 grid((unsigned)((igw + FC_NC - 1) / FC_NC), (unsigned)igh);
    // 16-byte asynchronous staging needs 16-byte aligned flow rows (and mask rows, when there is a mask)
    const bool async = ((reinterpret_cast<uintptr_t>(d_flow) & 15u) | (flow_stride & 3u)) == 0 &&
                       (!d_mask || ((reinterpret_cast<uintptr_t>(d_mask) & 15u) | (mask_stride & 15u)) == 0);
#ifndef OFPSB_EMU
    static bool smem_opt_in_dev[64] = {};   // > 48 KB of dynamic shared memory needs the opt-in, once per device and kernel
    int dev = 0;
    OFPSB_CUDA_TRY(cudaGetDevice(&dev));
    bool& smem_opt_in = smem_opt_in_dev[dev & 63];
    if (!smem_opt_in) {
        OFPSB_CUDA_TRY(cudaFuncSetAttribute(flow_cells_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FC_SMEM_BYTES + FC_SMEM_PAD));
        OFPSB_CUDA_TRY(cudaFuncSetAttribute(flow_cells_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FC_SMEM_BYTES + FC_SMEM_PAD));
        smem_opt_in = true;
    }
#endif
    if (async)
        OFPSB_LAUNCH_SMEM(flow_cells_kernel<true>, grid, 64, FC_SMEM_BYTES + FC_SMEM_PAD, stream, d_flow, (long long)flow_stride, d_mask,
                          (long long)mask_stride, w, h, igw, igh, cells, colcount);
    else
        OFPSB_LAUNCH_SMEM(flow_cells_kernel<false>, grid, 64, FC_SMEM_BYTES + FC_SMEM_PAD, stream, d_flow, (long long)flow_stride, d_mask,
                          (long long)mask_stride, w, h, igw, igh, cells, colcount);
    OFPSB_LAUNCH(flow_emit_cells_kernel, (unsigned)((igw + FE_WARPS - 1) / FE_WARPS), 32 * FE_WARPS, stream, cells, colcount, igw, igh,
                 d_entries, cap, d_count);
    OFPSB_CUDA_TRY(cudaGetLastError());
    if (launches) *launches += 2;
    return OFPSB_OK;
}

}  // namespace ofpsb
