/*
 * cv_front.c — CPU ORACLE (test infrastructure, NOT product code), part 2:
 * the dense-flow front end of the reference's `cv-decoder`:
 * everything `CvDecoder::process_frame` does around the third-party optical-flow call —
 * colour -> luma, the contrast mask, and the conversion of a dense flow image into
 * MotionEntry values (optionally through the down-sampling MotionFieldDensifier).
 *
 * The arithmetic of the first two lives in OpenCV (system library, not under /root/reference,
 * no version pin beyond the crate major: cv-decoder/Cargo.toml:22).  It is restated here from the
 * published algorithms and PINNED against OpenCV itself: tests/golden/make_golden_cv.py runs
 * cv2 4.13 (present in the build container) with the reference's exact call parameters and
 * commits inputs + outputs as tests/golden/golden_cv_v1.npz; tests/test_oracle_cv_front.py
 * checks this file against those vectors bit for bit.
 */
#include "ofps_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* cv::borderInterpolate(p, len, BORDER_REFLECT_101) — gfedcb|abcdefgh|gfedcba */
static int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * (len - 1) - p;
    }
    return p;
}

/* cv-decoder/src/lib.rs:138  imgproc::cvt_color(frame, gray, COLOR_BGR2GRAY).
 * OpenCV 4.x, 8-bit: gray = (B*3735 + G*19235 + R*9798 + 2^14) >> 15  (BY15/GY15/RY15;
 * verified against cv2 4.13 over all 2^24 colours).  channels = 3 or 4 (4th ignored);
 * rgb_order != 0 swaps the roles of channel 0 and 2 (COLOR_RGB2GRAY). */
void orc_bgr_to_gray(const uint8_t *src, int w, int h, int stride, int channels, int rgb_order,
                     uint8_t *gray, int gray_stride)
{
    for (int y = 0; y < h; y++) {
        const uint8_t *s = src + (size_t)y * stride;
        uint8_t *g = gray + (size_t)y * gray_stride;
        for (int x = 0; x < w; x++) {
            unsigned c0 = s[x * channels + 0], c1 = s[x * channels + 1], c2 = s[x * channels + 2];
            unsigned b = rgb_order ? c2 : c0, r = rgb_order ? c0 : c2;
            g[x] = (uint8_t)((b * 3735u + c1 * 19235u + r * 9798u + 16384u) >> 15);
        }
    }
}

/* cv-decoder/src/lib.rs:145-153: out_frame.push(RGBA::from_rgb_slice(&[bgr[2], bgr[1], bgr[0]])),
 * ofps/src/decoder.rs:19-26 (a = 255). */
void orc_bgr_to_rgba(const uint8_t *src, int w, int h, int stride, int channels, uint8_t *rgba)
{
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const uint8_t *s = src + (size_t)y * stride + (size_t)x * channels;
            uint8_t *d = rgba + 4 * ((size_t)y * w + x);
            d[0] = s[2]; d[1] = s[1]; d[2] = s[0]; d[3] = 255;
        }
}

/* cv-decoder/src/lib.rs:90-118: motion-field size from the frame size, the aspect-ratio scale and
 * the "Width"/"Height" properties (usize arithmetic, truncating division). */
void orc_mfield_size(size_t frame_w, size_t frame_h, size_t ar_x, size_t ar_y, size_t max_w, size_t max_h,
                     size_t *dx, size_t *dy)
{
    size_t r0 = frame_w * ar_x, r1 = frame_h * ar_y;
    size_t w = max_w < frame_w ? max_w : frame_w;
    size_t h = max_h < frame_h ? max_h : frame_h;
    size_t wb0 = w, wb1 = r0 ? w * r1 / r0 : 0;
    size_t hb0 = r1 ? h * r0 / r1 : 0, hb1 = h;
    if (wb0 < hb0) { *dx = wb0; *dy = wb1; } else { *dx = hb0; *dy = hb1; }
}

/* cv-decoder/src/lib.rs:127-135: imgproc::resize(frame_tmp, frame, (dx, dy), 0, 0, INTER_LINEAR) on 8-bit pixels
 * ("Process Fullres" off).  OpenCV's fixed-point bilinear path (resizeGeneric_, HResizeLinear / VResizeLinear,
 * INTER_RESIZE_COEF_BITS = 11), restated and verified bit for bit against cv2 4.13:
 *   scale = (double)src/dst;  fx = (float)((d + 0.5)*scale - 0.5);  s = floor(fx);  fx -= s;
 *   s < 0 -> (s, fx) = (0, 0);  s >= src-1 -> (s, fx) = (src-1, 0);
 *   a0 = cvRound((1-fx)*2048), a1 = cvRound(fx*2048)            (f32 products, round half to even, as short)
 *   row[d] = S[s]*a0 + S[s+1]*a1                                  (int32; S[s+1] clamped, its weight is 0 there)
 *   dst = ( ((b0*(row0 >> 4)) >> 16) + ((b1*(row1 >> 4)) >> 16) + 2 ) >> 2
 * (An exact 2x reduction is routed to INTER_AREA by OpenCV, which gives the same values.) */
static void resize_coeffs(int src, int dst, int *ofs, int *a0, int *a1)
{
    double scale = (double)src / (double)dst;
    for (int d = 0; d < dst; d++) {
        float fx = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(fx);
        fx -= (float)s;
        if (s < 0) { fx = 0.0f; s = 0; }
        if (s >= src - 1) { fx = 0.0f; s = src - 1; }
        ofs[d] = s;
        a0[d] = (int)lrintf((1.0f - fx) * 2048.0f);
        a1[d] = (int)lrintf(fx * 2048.0f);
    }
}

void orc_resize_linear(const uint8_t *src, int sw, int sh, int stride, int channels, uint8_t *dst, int dw, int dh)
{
    if (sw <= 0 || sh <= 0 || dw <= 0 || dh <= 0) return;
    int *xo = (int *)malloc(sizeof(int) * 3 * (size_t)dw), *yo = (int *)malloc(sizeof(int) * 3 * (size_t)dh);
    resize_coeffs(sw, dw, xo, xo + dw, xo + 2 * dw);
    resize_coeffs(sh, dh, yo, yo + dh, yo + 2 * dh);
    for (int y = 0; y < dh; y++) {
        const uint8_t *r0 = src + (size_t)yo[y] * stride;
        const uint8_t *r1 = src + (size_t)(yo[y] + 1 < sh ? yo[y] + 1 : sh - 1) * stride;
        int b0 = yo[dh + y], b1 = yo[2 * dh + y];
        for (int x = 0; x < dw; x++) {
            int x0 = xo[x], x1 = x0 + 1 < sw ? x0 + 1 : sw - 1;
            int a0 = xo[dw + x], a1 = xo[2 * dw + x];
            for (int c = 0; c < channels; c++) {
                int h0 = r0[x0 * channels + c] * a0 + r0[x1 * channels + c] * a1;
                int h1 = r1[x0 * channels + c] * a0 + r1[x1 * channels + c] * a1;
                int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
                dst[((size_t)y * dw + x) * channels + c] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
            }
        }
    }
    free(xo); free(yo);
}

/* cv-decoder/src/lib.rs:204-236: the contrast mask of the Farneback path.
 *   sobel  = Sobel(gray, CV_32F, dx=1, dy=1, ksize=5, scale 1, delta 0, BORDER_DEFAULT)
 *            separable kernel [-1,-2,0,2,1] in both directions (cv::getDerivKernels(1,1,5)),
 *            BORDER_DEFAULT = BORDER_REFLECT_101; exact in integers (|sobel| <= 36*255);
 *   thresh = threshold(sobel, 20, 255, THRESH_BINARY): sobel > 20 ? 255 : 0   (strict, signed);
 *   mask   = dilate(thresh, getStructuringElement(MORPH_ELLIPSE, 11x11, anchor (5,5)), BORDER_DEFAULT):
 *            row half-widths 0,3,4,5,5,5,5,5,4,3,0 (dx = round(5*sqrt(1 - dy^2/25))); out-of-frame taps
 *            read the REFLECTED threshold image (borderType is REFLECT_101 here, not dilate's usual
 *            constant border).
 * mask: w*h bytes, 0 or 255 (the reference keeps it as f32 and tests `mask < 0.1`, :258).
 * sobel_out (optional): w*h int32. */
static const int ELLIPSE_HW[6] = { 5, 5, 5, 4, 3, 0 };   /* by |dy| */

void orc_contrast_mask(const uint8_t *gray, int w, int h, int stride, uint8_t *mask, int32_t *sobel_out)
{
    static const int K[5] = { -1, -2, 0, 2, 1 };
    if (w <= 0 || h <= 0) return;
    uint8_t *th = (uint8_t *)malloc((size_t)w * h);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int acc = 0;
            for (int i = -2; i <= 2; i++) {
                if (!K[i + 2]) continue;
                const uint8_t *row = gray + (size_t)reflect101(y + i, h) * stride;
                int hsum = 0;
                for (int j = -2; j <= 2; j++)
                    if (K[j + 2]) hsum += K[j + 2] * (int)row[reflect101(x + j, w)];
                acc += K[i + 2] * hsum;
            }
            if (sobel_out) sobel_out[(size_t)y * w + x] = acc;
            th[(size_t)y * w + x] = acc > 20;
        }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int v = 0;
            for (int dy = -5; dy <= 5 && !v; dy++) {
                const uint8_t *row = th + (size_t)reflect101(y + dy, h) * w;
                int hw = ELLIPSE_HW[dy < 0 ? -dy : dy];
                for (int dx = -hw; dx <= hw; dx++)
                    if (row[reflect101(x + dx, w)]) { v = 1; break; }
            }
            mask[(size_t)y * w + x] = v ? 255 : 0;
        }
    free(th);
}

/* cv-decoder/src/lib.rs:238-291: dense flow image -> MotionEntry list.
 *   frame_norm = (1/W, 1/H) in f32 (:238-241); for y, x in raster order, skipping pixels whose mask
 *   is 0 when a mask is given (:254-260): pos = (x+0.5, y+0.5) .* frame_norm, motion = flow .* frame_norm.
 *   gw == 0 (process_fullres = false): push every kept pixel (:272).
 *   gw > 0  (process_fullres = true):  add_vector into a gw x gh MotionFieldDensifier, remember the touched
 *   cells in a BTreeSet<(x,y)> (:270); then for the touched cells in (x, y) lexicographic order push
 *   pos = (x+0.5, y+0.5) .* (1/gw, 1/gh), motion = mean of the cell (:276-289).
 * flow: h rows of w (fx, fy) pairs, row pitch flow_stride floats.  Returns the number of entries
 * (written up to cap). */
typedef struct { size_t x, y; } cellxy;
static int cell_cmp(const void *a, const void *b)
{
    const cellxy *p = (const cellxy *)a, *q = (const cellxy *)b;
    if (p->x != q->x) return p->x < q->x ? -1 : 1;
    if (p->y != q->y) return p->y < q->y ? -1 : 1;
    return 0;
}

size_t orc_flow_entries(const float *flow, size_t flow_stride, const uint8_t *mask, size_t mask_stride, int w, int h,
                        size_t gw, size_t gh, orc_mv *out, size_t cap)
{
    if (w <= 0 || h <= 0) return 0;
    const float nx = 1.0f / (float)w, ny = 1.0f / (float)h;
    size_t n = 0;
    if (gw == 0 || gh == 0) {
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                if (mask && mask[(size_t)y * mask_stride + x] == 0) continue;
                const float *f = flow + (size_t)y * flow_stride + 2 * (size_t)x;
                if (n < cap) {
                    out[n].px = ((float)x + 0.5f) * nx;
                    out[n].py = ((float)y + 0.5f) * ny;
                    out[n].mx = f[0] * nx;
                    out[n].my = f[1] * ny;
                }
                n++;
            }
        return n;
    }
    size_t cells = gw * gh;
    float *sums = (float *)malloc(sizeof(float) * 2 * cells);
    float *counts = (float *)malloc(sizeof(float) * 2 * cells);
    uint8_t *touched = (uint8_t *)calloc(cells, 1);
    orc_densifier_init(sums, counts, gw, gh);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            if (mask && mask[(size_t)y * mask_stride + x] == 0) continue;
            const float *f = flow + (size_t)y * flow_stride + 2 * (size_t)x;
            size_t cx, cy;
            orc_densifier_add(sums, counts, gw, gh, ((float)x + 0.5f) * nx, ((float)y + 0.5f) * ny, f[0] * nx, f[1] * ny,
                              1.0f, &cx, &cy);
            touched[cy * gw + cx] = 1;
        }
    float *field = (float *)malloc(sizeof(float) * 2 * cells);
    orc_densifier_finish(sums, counts, gw, gh, field);
    cellxy *pts = (cellxy *)malloc(sizeof(cellxy) * cells);
    size_t np = 0;
    for (size_t c = 0; c < cells; c++)
        if (touched[c]) { pts[np].x = c % gw; pts[np].y = c / gw; np++; }
    qsort(pts, np, sizeof(cellxy), cell_cmp);
    const float gx = 1.0f / (float)gw, gy = 1.0f / (float)gh;
    for (size_t i = 0; i < np; i++) {
        if (n < cap) {
            size_t c = pts[i].y * gw + pts[i].x;
            out[n].px = ((float)pts[i].x + 0.5f) * gx;
            out[n].py = ((float)pts[i].y + 0.5f) * gy;
            out[n].mx = field[2 * c];
            out[n].my = field[2 * c + 1];
        }
        n++;
    }
    free(sums); free(counts); free(touched); free(field); free(pts);
    return n;
}
