/*
 * ofps_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference hot path of h33p/ofps, used only as
 * the checker in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  Nothing in ofps_b200/ may import, link or call it.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference checkout).  The reference is Rust; no rustc exists in the build
 * container, so the reference itself cannot be compiled (oracle/_ref is
 * therefore absent) and its linear algebra dependency nalgebra 0.30 (semver
 * range, Cargo.lock git-ignored) is not vendored: its formulas are restated
 * from the published algorithm.
 *
 * PARITY PINNING
 *   pinned   : StandardCamera + solve_ypr_given  — against the reference's own
 *              tests (almeida-estimator/src/lib.rs:257-372: 50x50 grid, camera
 *              (1.0, 90deg), 32 rotations, err < 0.1*rot) and the point_angle
 *              doctest (ofps/src/camera.rs:144-148), re-expressed in
 *              tests/test_oracle_reference_kats.py.
 *   UNPINNED : MotionFieldDensifier, BlockMotionDetection::detect_motion,
 *              interpolate_empty_cells (no reference test or fixture exists);
 *              solve_ypr_ransac (reference RNG is rand::thread_rng(), not
 *              reproducible; only the statistical reference test constrains it);
 *              block_match (the reference contains no SAD search at all — this
 *              oracle IS the specification, see orc_block_match).
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off -fno-fast-math: Rust never
 * contracts a*b+c into an FMA).
 */
#ifndef OFPS_ORACLE_H
#define OFPS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ofps/src/decoder.rs:40  MotionEntry = (Point2<f32> pos, Vector2<f32> motion),
 * flattened in the order motion-extract writes it (motion-extract/src/main.rs:27-29). */
typedef struct { float px, py, mx, my; } orc_mv;

/* ------------------------------------------------------------------ densifier */
/* MotionFieldDensifier::new (ofps/src/motion_field.rs:133-138): sums = 0,
 * counts = f32::EPSILON.  Layout: cell-major pairs [x0,y0,x1,y1,...], cell = y*w+x. */
void orc_densifier_init(float *sums, float *counts, size_t w, size_t h);
/* add_vector_weighted (motion_field.rs:164-178) + add_vector_idx (:141-147). */
void orc_densifier_add(float *sums, float *counts, size_t w, size_t h,
                       float px, float py, float mx, float my, float weight,
                       size_t *cell_x, size_t *cell_y);
/* impl From<MotionFieldDensifier> for MotionField (motion_field.rs:297-308). */
void orc_densifier_finish(const float *sums, const float *counts, size_t w, size_t h, float *field);
/* add_vector over a slice + finish; counts_out optional (2*w*h). */
void orc_densify(const orc_mv *mv, size_t n, size_t w, size_t h, float *field, float *counts_out);
/* interpolate_empty_cells (motion_field.rs:193-294), in place on sums/counts. */
void orc_interpolate_empty_cells(float *sums, float *counts, size_t w, size_t h);
/* flow-extract dense pipeline (flow-extract/src/main.rs:72-83): densify -> interpolate -> finish. */
void orc_flow_field(const orc_mv *mv, size_t n, size_t w, size_t h, float *field);

/* ------------------------------------------------------------------- detector */
/* block_dim = ceil(1/(sqrt(min_size)/subdivide)) (block-motion-detector/src/lib.rs:53-54). */
size_t orc_block_dim(float min_size, size_t subdivide);
/* BlockMotionDetection::detect_motion (block-motion-detector/src/lib.rs:49-119).
 * Returns 1 = Some, 0 = None, <0 = error.  field: dim*dim*2 floats (zeroed when None);
 * mean_field (optional): the full densified field before island filtering. */
int orc_detect_block_motion(const orc_mv *mv, size_t n, float min_size, size_t subdivide,
                            float target_motion, size_t *area, size_t *dim,
                            float *field, size_t field_cap_cells, float *mean_field);

/* ------------------------------------------------------ camera (f32 and f64) */
typedef struct {
    float aspect, fov_y;
    float proj[4];      /* m00, m11, m22, m23 (m32 = -1, m33 = 0) */
    float inv_proj[16]; /* row-major 4x4 */
    float unproj[16];   /* view^T * inv_proj, row-major (camera.rs:54 with rotate's view, :91-96) */
} orc_camera_f;
typedef struct {
    double aspect, fov_y;
    double proj[4];
    double inv_proj[16];
    double unproj[16];
} orc_camera_d;

void orc_camera_new_f(orc_camera_f *c, float aspect, float fov_y_deg);   /* camera.rs:26-35 */
void orc_camera_new_d(orc_camera_d *c, double aspect, double fov_y_deg);
/* generic unproject/project with an arbitrary view matrix (row-major 4x4) (camera.rs:45-55, 72-81) */
void orc_camera_unproject_f(const orc_camera_f *c, float x, float y, const float *inv_view, float out[3]);
void orc_camera_project_f(const orc_camera_f *c, const float world[3], const float *view, float out[2]);
void orc_camera_unproject_d(const orc_camera_d *c, double x, double y, const double *inv_view, double out[3]);
void orc_camera_project_d(const orc_camera_d *c, const double world[3], const double *view, double out[2]);
/* delta (camera.rs:89-117); rot = row-major 4x4 */
void orc_camera_delta_f(const orc_camera_f *c, float x, float y, const float *rot, float out[2]);
void orc_camera_delta_d(const orc_camera_d *c, double x, double y, const double *rot, double out[2]);
/* point_angle (camera.rs:150-161), radians */
void orc_camera_point_angle_f(const orc_camera_f *c, float x, float y, float out[2]);
void orc_camera_point_angle_d(const orc_camera_d *c, double x, double y, double out[2]);

/* ---------------------------------------------------------- almeida estimator */
/* solve_ypr_given (almeida-estimator/src/lib.rs:123-200).  quat = (w,i,j,k). */
void orc_almeida_lsq_f(const orc_mv *mv, size_t n, float aspect, float fov_y_deg, float quat[4]);
/* f64 twin: same algorithm evaluated in double (inputs are the f32 entries). */
void orc_almeida_lsq_d(const orc_mv *mv, size_t n, double aspect, double fov_y_deg, double quat[4]);
/* solve_ypr_ransac (almeida:202-251) with the seeded counter RNG described in
 * ofps_oracle.c (the reference's thread_rng is not reproducible).
 * best_count/best_iter optional. */
void orc_almeida_ransac_f(const orc_mv *mv, size_t n, float aspect, float fov_y_deg,
                          size_t num_iters, float inlier_angle_deg, size_t num_samples,
                          uint64_t seed, float quat[4], size_t *best_count, size_t *best_iter);
/* RNG exposed for index-parity tests: j-th element of the keyed permutation of [0,n). */
uint64_t orc_perm_index(uint64_t seed, uint64_t iter, uint64_t stream, uint64_t j, uint64_t n);

/* helpers for the reference test matrix (almeida:257-306) */
void orc_quat_from_euler_f(float roll, float pitch, float yaw, float q[4]);
void orc_quat_from_euler_d(double roll, double pitch, double yaw, double q[4]);
void orc_look_at_rh_view_d(const double q[4], double view[16]);   /* calc_view (almeida:280-286), eye = 0 */
double orc_quat_angle_to_d(const double a[4], const double b[4]); /* radians */

/* --------------------------------------------------------------- block match */
/* Exhaustive block matcher — SPECIFICATION (no reference counterpart):
 *   blocks BxB at (bx*B, by*B) in cur, only full blocks (bx < W/B, by < H/B);
 *   candidates (dx,dy) in [-R,R]^2 whose BxB window lies fully inside prev;
 *   cost = SAD (metric 0) or SSD (metric 1);
 *   winner = lexicographic min of (cost, dx*dx+dy*dy, dy, dx);
 *   outputs per block (raster order): mv_xy[2*i] = dx, mv_xy[2*i+1] = dy, cost[i],
 *   entries[i] following av-decoder/src/lib.rs:404-419 with dst = block centre,
 *   src = dst + (dx,dy): pos = src * (1/W, 1/H), motion = (dx,dy) * -(1/W, 1/H).
 * threads <= 1: single thread; otherwise OpenMP over blocks.  Returns nblocks or <0. */
long orc_block_match(const uint8_t *prev, const uint8_t *cur, int w, int h, int stride,
                     int block, int range, int metric,
                     int16_t *mv_xy, uint32_t *cost, orc_mv *entries, int threads);
/* Same contract, SIMD (psadbw) SAD inner loop where available; used as the timed CPU baseline. */
long orc_block_match_fast(const uint8_t *prev, const uint8_t *cur, int w, int h, int stride,
                          int block, int range, int metric,
                          int16_t *mv_xy, uint32_t *cost, orc_mv *entries, int threads);
int orc_max_threads(void);

/* ------------------------------------------------- cv-decoder dense-flow front end (cv_front.c) */
/* PINNED against OpenCV itself (cv2 4.13 run with the reference's call parameters;
 * tests/golden/make_golden_cv.py -> tests/golden/golden_cv_v1.npz). */
/* cvtColor(COLOR_BGR2GRAY) on 8-bit pixels (cv-decoder/src/lib.rs:138). */
void orc_bgr_to_gray(const uint8_t *src, int w, int h, int stride, int channels, int rgb_order,
                     uint8_t *gray, int gray_stride);
/* out_frame: BGR(A) -> RGBA, a = 255 (cv-decoder/src/lib.rs:145-153; ofps/src/decoder.rs:19-26). */
void orc_bgr_to_rgba(const uint8_t *src, int w, int h, int stride, int channels, uint8_t *rgba);
/* motion-field size from frame size / aspect scale / max size (cv-decoder/src/lib.rs:90-118). */
void orc_mfield_size(size_t frame_w, size_t frame_h, size_t ar_x, size_t ar_y, size_t max_w, size_t max_h,
                     size_t *dx, size_t *dy);
/* resize(INTER_LINEAR) on 8-bit interleaved pixels (cv-decoder/src/lib.rs:127-135); dst = dw*dh*channels. */
void orc_resize_linear(const uint8_t *src, int sw, int sh, int stride, int channels, uint8_t *dst, int dw, int dh);
/* Sobel(1,1,k5) -> threshold(>20) -> dilate(11x11 ellipse), all BORDER_REFLECT_101
 * (cv-decoder/src/lib.rs:204-236).  mask: w*h bytes 0/255; sobel_out optional (w*h int32). */
void orc_contrast_mask(const uint8_t *gray, int w, int h, int stride, uint8_t *mask, int32_t *sobel_out);
/* dense flow (+ optional mask) -> MotionEntry list, per pixel (gw == 0) or through the gw x gh
 * down-sampling densifier, touched cells in BTreeSet<(x,y)> order (cv-decoder/src/lib.rs:238-291). */
size_t orc_flow_entries(const float *flow, size_t flow_stride, const uint8_t *mask, size_t mask_stride, int w, int h,
                        size_t gw, size_t gh, orc_mv *out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
