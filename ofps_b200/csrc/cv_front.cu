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
__device__ __forceinline__ void async_copy_wait()
{
#ifndef OFPSB_EMU
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
#endif
}

// ------------------------------------------------------------------------------------------ K7
// OpenCV 4.x 8-bit BGR2GRAY: (B*3735 + G*19235 + R*9798 + 2^14) >> 15.
__device__ __forceinline__ uint32_t gray_of(uint32_t c0, uint32_t c1, uint32_t c2, bool rgb_order)
{
    const uint32_t b = rgb_order ? c2 : c0, r = rgb_order ? c0 : c2;
    return (b * 3735u + c1 * 19235u + r * 9798u + 16384u) >> 15;
}

// One thread = 4 adjacent pixels of one row.  VEC: rows start on 4-byte boundaries (src and gray), so the
// 12 / 16 source bytes are three / four aligned words and the four gray bytes one word store.
template <int CH, bool VEC>
__global__ void __launch_bounds__(256) frame_convert_kernel(const uint8_t* __restrict__ src, int w, int h, int stride,
                                                            int rgb_order, uint8_t* __restrict__ gray, int gray_stride,
                                                            uint32_t* __restrict__ rgba)
{
    const int qpr = (w + 3) >> 2;   // pixel quads per row
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (long long)qpr * h) return;
    const int y = (int)(id / qpr), x = 4 * (int)(id % qpr);
    const uint8_t* s = src + (size_t)y * stride + (size_t)x * CH;
    uint32_t px[4][3];
    const int nv = min(4, w - x);
    if (VEC && nv == 4) {
        const uint32_t* s4 = reinterpret_cast<const uint32_t*>(s);
        if (CH == 3) {
            const uint32_t a = __ldg(s4), b = __ldg(s4 + 1), c = __ldg(s4 + 2);
            px[0][0] = a & 255u;         px[0][1] = (a >> 8) & 255u;  px[0][2] = (a >> 16) & 255u;
            px[1][0] = a >> 24;          px[1][1] = b & 255u;         px[1][2] = (b >> 8) & 255u;
            px[2][0] = (b >> 16) & 255u; px[2][1] = b >> 24;          px[2][2] = c & 255u;
            px[3][0] = (c >> 8) & 255u;  px[3][1] = (c >> 16) & 255u; px[3][2] = c >> 24;
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t a = __ldg(s4 + k);
                px[k][0] = a & 255u; px[k][1] = (a >> 8) & 255u; px[k][2] = (a >> 16) & 255u;
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++)
#pragma unroll
            for (int c = 0; c < 3; c++) px[k][c] = k < nv ? (uint32_t)__ldg(s + k * CH + c) : 0u;
    }
    if (gray) {
        uint32_t g[4];
#pragma unroll
        for (int k = 0; k < 4; k++) g[k] = gray_of(px[k][0], px[k][1], px[k][2], rgb_order != 0);
        uint8_t* o = gray + (size_t)y * gray_stride + x;
        if (VEC && nv == 4) {
            *reinterpret_cast<uint32_t*>(o) = g[0] | (g[1] << 8) | (g[2] << 16) | (g[3] << 24);
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (k < nv) o[k] = (uint8_t)g[k];
        }
    }
    if (rgba) {   // RGBA::from_rgb_slice(&[bgr[2], bgr[1], bgr[0]]): r | g << 8 | b << 16 | 255 << 24 (little endian)
        uint32_t* o = rgba + (size_t)y * w + x;
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (k < nv) o[k] = px[k][2] | (px[k][1] << 8) | (px[k][0] << 16) | 0xFF000000u;
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
constexpr int CM_BC = CM_TW + 10;   // threshold columns t0-5 .. t0+TW+4: bit b <-> column t0-5+b
constexpr int CM_BW = 9;            // words per threshold row (288 bits >= 266)
constexpr int CM_BH = CM_TH + 10;   // threshold rows ty0-5 .. ty0+TH+4
static_assert(CM_NT == 32 * CM_BW, "one warp per threshold word column");

__global__ void __launch_bounds__(CM_NT) contrast_mask_kernel(const uint8_t* __restrict__ gray, int w, int h, int stride,
                                                              uint8_t* __restrict__ mask, int mask_stride)
{
    __shared__ __align__(16) uint8_t G[CM_GH][CM_GW];
    __shared__ uint32_t T[CM_BH][CM_BW];
    const int tid = threadIdx.x;
    const int t0 = blockIdx.x * CM_TW, ty0 = blockIdx.y * CM_TH;
    const int gx0 = t0 - 16, gy0 = ty0 - 7;

    // (1) gray tile
    {
        constexpr int CPR = CM_GW / 16;   // chunks per row
        const bool aligned = ((reinterpret_cast<uintptr_t>(gray) | (uintptr_t)stride) & 15u) == 0;
        for (int i = tid; i < CM_GH * CPR; i += CM_NT) {
            const int r = i / CPR, k = i - r * CPR;
            const int x = gx0 + 16 * k;
            const uint8_t* row = gray + (size_t)reflect101(gy0 + r, h) * stride;
            if (aligned && x >= 0 && x + 16 <= w) {
                async_copy16(&G[r][16 * k], row + x);
            } else {
#pragma unroll 4
                for (int j = 0; j < 16; j++) G[r][16 * k + j] = __ldg(row + reflect101(x + j, w));
            }
        }
        async_copy_wait();
    }
    __syncthreads();

    // (2) Sobel(dx=1, dy=1, ksize 5) = [-1,-2,0,2,1]^T x [-1,-2,0,2,1], threshold, ballot
    {
        const int wc = tid >> 5, lane = tid & 31;
        const int b = 32 * wc + lane;
        const int px = t0 - 5 + b;
        const int c = b + 11;   // column of px in G
        const bool col_ok = b < CM_BC && px >= 0 && px < w;
        int h0 = 0, h1 = 0, h2 = 0, h3 = 0;   // horizontal derivative at gray rows r-4 .. r-1
        for (int r = 0; r < CM_GH; r++) {
            int hr = 0;
            if (col_ok) {
                const uint8_t* g = &G[r][c];
                hr = ((int)g[2] - (int)g[-2]) + 2 * ((int)g[1] - (int)g[-1]);
            }
            if (r >= 4) {   // centre = gray row r-2 = frame row ty0-5+(r-4)
                const int s = (hr - h0) + 2 * (h3 - h1);
                const int py = ty0 - 5 + (r - 4);
                const bool on = col_ok && py >= 0 && py < h && s > 20;
                const unsigned word = __ballot_sync(0xffffffffu, on);
                if (lane == 0) T[r - 4][wc] = word;
            }
            h0 = h1; h1 = h2; h2 = h3; h3 = hr;
        }
    }
    __syncthreads();

    // (3) out-of-frame threshold positions take the bit of their reflection (dilate's BORDER_REFLECT_101)
    if (t0 - 5 < 0 || t0 + CM_TW + 5 > w || ty0 - 5 < 0 || ty0 + CM_TH + 5 > h) {
        for (int i = tid; i < CM_BH * CM_BC; i += CM_NT) {
            const int br = i / CM_BC, b = i % CM_BC;
            const int px = t0 - 5 + b, py = ty0 - 5 + br;
            if (px >= 0 && px < w && py >= 0 && py < h) continue;
            const int qb = reflect101(px, w) - (t0 - 5), qr = reflect101(py, h) - (ty0 - 5);
            if (qb < 0 || qb >= CM_BC || qr < 0 || qr >= CM_BH) continue;   // read by no in-frame pixel of this tile
            if ((T[qr][qb >> 5] >> (qb & 31)) & 1u) atomicOr(&T[br][b >> 5], 1u << (b & 31));
        }
        __syncthreads();
    }

    // (4) dilation by the 11x11 ellipse: row half-widths 0,3,4,5,5,5,5,5,4,3,0
    if (tid < (CM_TW / 32) * CM_TH) {
        const int orow = tid >> 3, ow = tid & 7;
        const int y = ty0 + orow, x0 = t0 + 32 * ow;
        if (y < h && x0 < w) {
            // 64-bit window of threshold row dy: bit i <-> column x0-5+i
            auto win = [&](int dy) -> unsigned long long {
                const uint32_t* t = T[orow + 5 + dy];
                return (unsigned long long)t[ow] | ((unsigned long long)t[ow + 1] << 32);
            };
            const unsigned long long v5 = win(-2) | win(-1) | win(0) | win(1) | win(2);
            const unsigned long long v4 = win(-3) | win(3), v3 = win(-4) | win(4), v0 = win(-5) | win(5);
            // run-OR of n = 2k+1 consecutive bits, then >> (5-k) centres it on the output pixel
            const unsigned long long a2 = v5 | (v5 >> 1), a4 = a2 | (a2 >> 2), a8 = a4 | (a4 >> 4);
            uint32_t m = (uint32_t)(a8 | (a4 >> 7));                           // k = 5: 11 bits
            const unsigned long long b2 = v4 | (v4 >> 1), b4 = b2 | (b2 >> 2), b8 = b4 | (b4 >> 4);
            m |= (uint32_t)((b8 | (v4 >> 8)) >> 1);                            // k = 4: 9 bits
            const unsigned long long c2 = v3 | (v3 >> 1), c4 = c2 | (c2 >> 2);
            m |= (uint32_t)((c4 | (c4 >> 3)) >> 2);                            // k = 3: 7 bits
            m |= (uint32_t)(v0 >> 5);                                          // k = 0
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

// starts[c] = first pixel coordinate whose cell index is >= c, c = 0..g (starts[g] = len).
__global__ void cell_bounds_kernel(int w, int h, int gw, int gh, int* __restrict__ xs, int* __restrict__ ys)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= w) {
        const float inv = __fdiv_rn(1.0f, (float)w), gm1 = (float)(gw - 1);
        const int c = i < w ? cell_coord(i, inv, gm1) : gw;
        const int cp = i > 0 ? cell_coord(i - 1, inv, gm1) : -1;
        for (int k = cp + 1; k <= c && k <= gw; k++) xs[k] = i;
    } else if (i - (w + 1) <= h) {
        const int j = i - (w + 1);
        const float inv = __fdiv_rn(1.0f, (float)h), gm1 = (float)(gh - 1);
        const int c = j < h ? cell_coord(j, inv, gm1) : gh;
        const int cp = j > 0 ? cell_coord(j - 1, inv, gm1) : -1;
        for (int k = cp + 1; k <= c && k <= gh; k++) ys[k] = j;
    }
}

constexpr float F32_EPSILON = 1.1920928955078125e-07f;
constexpr int FC_NT = 64;         // threads per CTA (all stage pixels; the first FC_NC fold)
constexpr int FC_NC = 16;         // cells per CTA (of one cell row): small CTAs, many resident per SM, so the
                                  // sequential folds of some overlap the loads of others
constexpr int FC_CAP = 2048;      // pixels staged per step (16 KB of flow + 2 KB of mask)

struct CellRec { float mx, my; uint32_t touched, pad; };

// One CTA = FC_NC adjacent cells of one cell row.  The pixel rectangle of those cells is staged through
// shared memory with coalesced loads (several image rows per step when the span is narrow); thread t then
// folds the pixels of cell t in raster order — rows top to bottom, columns left to right — which is the
// order in which the reference's loop reaches that cell, so sums and counts are bit-identical:
//   counts += 1.0 (from f32::EPSILON), sum = motion * 1.0 + sum, motion = flow .* (1/W, 1/H).
template <bool ASYNC>   // ASYNC: flow rows are 8-byte aligned -> staged by asynchronous 8-byte copies
__global__ void __launch_bounds__(FC_NT) flow_cells_kernel(const float* __restrict__ flow, long long flow_stride,
                                                           const uint8_t* __restrict__ mask, long long mask_stride,
                                                           int w, int h, int gw, int gh, const int* __restrict__ xs,
                                                           const int* __restrict__ ys, CellRec* __restrict__ cells)
{
    __shared__ float2 sflow[FC_CAP];
    __shared__ uint8_t smask[FC_CAP];
    const int tid = threadIdx.x;
    const int cy = blockIdx.y, c0 = blockIdx.x * FC_NC, c1 = min(c0 + FC_NC, gw);
    const int y0 = ys[cy], y1 = ys[cy + 1];
    const int px0 = xs[c0], px1 = xs[c1];
    const int cell = c0 + tid;
    const bool owner = tid < FC_NC && cell < c1;
    const int xa = owner ? xs[cell] : 0, xb = owner ? xs[cell + 1] : 0;
    const float nx = __fdiv_rn(1.0f, (float)w), ny = __fdiv_rn(1.0f, (float)h);
    float sx = 0.0f, sy = 0.0f, cnt = F32_EPSILON;
    uint32_t hits = 0;
    const int span = px1 - px0;
    if (span > 0 && y1 > y0) {
        const int cw = min(span, FC_CAP);
        const int rg = max(1, FC_CAP / cw);   // rows per step; > 1 only when the whole span fits (cw == span)
        for (int r0 = y0; r0 < y1; r0 += rg) {
            const int nr = min(rg, y1 - r0);
            for (int cx0 = px0; cx0 < px1; cx0 += cw) {
                const int nc = min(cw, px1 - cx0);
                for (int r = 0; r < nr; r++) {
                    const float* frow = flow + (long long)(r0 + r) * flow_stride + 2ll * cx0;
                    float2* srow = sflow + r * cw;
                    for (int c = tid; c < nc; c += FC_NT) {
                        if (ASYNC) async_copy8(srow + c, frow + 2 * c);
                        else srow[c] = make_float2(__ldg(frow + 2 * c), __ldg(frow + 2 * c + 1));
                    }
                }
                if (mask)
                    for (int r = 0; r < nr; r++) {
                        const uint8_t* mrow = mask + (long long)(r0 + r) * mask_stride + cx0;
#pragma unroll 4
                        for (int c = tid; c < nc; c += FC_NT) smask[r * cw + c] = __ldg(mrow + c);
                    }
                if (ASYNC) async_copy_wait();
                __syncthreads();
                if (owner) {
                    const int a = max(xa, cx0) - cx0, b = min(xb, cx0 + nc) - cx0;
                    for (int r = 0; r < nr; r++)
                        for (int c = a; c < b; c++) {
                            if (mask && smask[r * cw + c] == 0) continue;   // `*mask < 0.1` -> skip (cv-decoder:258)
                            const float2 f = sflow[r * cw + c];
                            cnt = __fadd_rn(cnt, 1.0f);
                            sx = __fadd_rn(__fmul_rn(f.x, nx), sx);
                            sy = __fadd_rn(__fmul_rn(f.y, ny), sy);
                            hits++;
                        }
                }
                __syncthreads();
            }
        }
    }
    if (owner) {
        CellRec rec;
        rec.mx = __fdiv_rn(sx, cnt);   // MotionField::from(densifier): sum ./ counts
        rec.my = __fdiv_rn(sy, cnt);
        rec.touched = hits ? 1u : 0u;
        rec.pad = 0u;
        cells[(size_t)cy * gw + cell] = rec;
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
// pos = (x + 0.5, y + 0.5) .* (1/gw, 1/gh), motion = cell mean.  One CTA walks the column-major index.
__global__ void __launch_bounds__(1024) flow_emit_cells_kernel(const CellRec* __restrict__ cells, int gw, int gh,
                                                               ofps_mv* __restrict__ out, unsigned long long cap,
                                                               unsigned long long* __restrict__ n_out)
{
    __shared__ uint32_t warp_tot[32];
    const float gx = __fdiv_rn(1.0f, (float)gw), gy = __fdiv_rn(1.0f, (float)gh);
    const long long total = (long long)gw * gh;
    unsigned long long base = 0;
    for (long long k0 = 0; k0 < total; k0 += blockDim.x) {
        const long long k = k0 + threadIdx.x;
        int x = 0, y = 0;
        CellRec rec = {0.f, 0.f, 0u, 0u};
        if (k < total) {
            x = (int)(k / gh);
            y = (int)(k - (long long)x * gh);
            rec = cells[(size_t)y * gw + x];
        }
        uint32_t tot;
        const uint32_t rank = block_excl_scan(rec.touched, warp_tot, &tot);
        if (rec.touched && base + rank < cap) {
            ofps_mv e;
            e.px = __fmul_rn(__fadd_rn((float)x, 0.5f), gx);
            e.py = __fmul_rn(__fadd_rn((float)y, 0.5f), gy);
            e.mx = rec.mx;
            e.my = rec.my;
            out[base + rank] = e;
        }
        base += tot;
    }
    if (threadIdx.x == 0) *n_out = base;
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
    const bool vec = ((reinterpret_cast<uintptr_t>(d_src) | (uintptr_t)stride) & 3u) == 0 &&
                     (!d_gray || ((reinterpret_cast<uintptr_t>(d_gray) | (uintptr_t)gray_stride) & 3u) == 0);
    const long long quads = (long long)((w + 3) / 4) * h;
    const unsigned grid = (unsigned)((quads + 255) / 256);
    uint32_t* rgba = reinterpret_cast<uint32_t*>(d_rgba);
    if (channels == 3) {
        if (vec) OFPSB_LAUNCH((frame_convert_kernel<3, true>), grid, 256, stream, d_src, w, h, stride, rgb_order, d_gray, gray_stride, rgba);
        else OFPSB_LAUNCH((frame_convert_kernel<3, false>), grid, 256, stream, d_src, w, h, stride, rgb_order, d_gray, gray_stride, rgba);
    } else {
        if (vec) OFPSB_LAUNCH((frame_convert_kernel<4, true>), grid, 256, stream, d_src, w, h, stride, rgb_order, d_gray, gray_stride, rgba);
        else OFPSB_LAUNCH((frame_convert_kernel<4, false>), grid, 256, stream, d_src, w, h, stride, rgb_order, d_gray, gray_stride, rgba);
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
    if (int rc = scratch.bounds.reserve((gw + gh + 2) * sizeof(int))) return rc;
    if (int rc = scratch.cells.reserve(gw * gh * sizeof(CellRec))) return rc;
    int* xs = scratch.bounds.as<int>();
    int* ys = xs + gw + 1;
    CellRec* cells = scratch.cells.as<CellRec>();
    OFPSB_LAUNCH(cell_bounds_kernel, (unsigned)((w + h + 2 + 255) / 256), 256, stream, w, h, igw, igh, xs, ys);
    const dim3 grid((unsigned)((igw + FC_NC - 1) / FC_NC), (unsigned)igh);
    if (((reinterpret_cast<uintptr_t>(d_flow) & 7u) | (flow_stride & 1u)) == 0)
        OFPSB_LAUNCH(flow_cells_kernel<true>, grid, FC_NT, stream, d_flow, (long long)flow_stride, d_mask, (long long)mask_stride, w,
                     h, igw, igh, xs, ys, cells);
    else
        OFPSB_LAUNCH(flow_cells_kernel<false>, grid, FC_NT, stream, d_flow, (long long)flow_stride, d_mask, (long long)mask_stride, w,
                     h, igw, igh, xs, ys, cells);
    OFPSB_LAUNCH(flow_emit_cells_kernel, 1, 1024, stream, cells, igw, igh, d_entries, cap, d_count);
    OFPSB_CUDA_TRY(cudaGetLastError());
    if (launches) *launches += 3;
    return OFPSB_OK;
}

}  // namespace ofpsb
