/*
 * ofps_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 * See ofps_oracle.h for the pinning statement.  Reference paths are relative to
 * the h33p/ofps checkout.
 */
#include "ofps_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

/* ======================================================================== */
/* MotionFieldDensifier (ofps/src/motion_field.rs:121-190, 297-308)          */
/* ======================================================================== */

void orc_densifier_init(float *sums, float *counts, size_t w, size_t h)
{
    /* motion_field.rs:133-138 */
    for (size_t i = 0; i < 2 * w * h; i++) { sums[i] = 0.0f; counts[i] = FLT_EPSILON; }
}

/* nalgebra::clamp(val, min, max) on Point2 (motion_field.rs:170).  nalgebra's
 * generic clamp is `if val > min { if val < max { val } else { max } } else { min }`
 * and Point/Matrix `>` / `<` hold only when EVERY component compares so.  Hence a
 * point with any component <= 0 collapses to (0,0) and one with any component
 * >= 1 (all > 0) collapses to (1,1).  [nalgebra 0.30 not vendored: restated from
 * its published source; strictly-inside points — everything the decoders emit —
 * are unaffected by this choice.] */
static void clamp_point(float *x, float *y)
{
    if (*x > 0.0f && *y > 0.0f) {
        if (*x < 1.0f && *y < 1.0f) return;
        *x = 1.0f; *y = 1.0f;
    } else {
        *x = 0.0f; *y = 0.0f;
    }
}

/* `f32 as usize`: saturating, NaN -> 0 */
static size_t f32_as_usize(float v)
{
    if (!(v > 0.0f)) return 0;
    if (v >= 18446744073709551616.0f) return SIZE_MAX;
    return (size_t)v;
}

static void densifier_add_idx(float *sums, float *counts, size_t idx, float mx, float my, float weight)
{
    /* motion_field.rs:141-147 */
    counts[2 * idx + 0] += weight;
    counts[2 * idx + 1] += weight;
    sums[2 * idx + 0] = mx * weight + sums[2 * idx + 0];
    sums[2 * idx + 1] = my * weight + sums[2 * idx + 1];
}

void orc_densifier_add(float *sums, float *counts, size_t w, size_t h,
                       float px, float py, float mx, float my, float weight,
                       size_t *cell_x, size_t *cell_y)
{
    /* motion_field.rs:164-178; f32::round is half-away-from-zero = roundf */
    clamp_point(&px, &py);
    size_t x = f32_as_usize(roundf(px * (float)(w - 1)));
    size_t y = f32_as_usize(roundf(py * (float)(h - 1)));
    densifier_add_idx(sums, counts, y * w + x, mx, my, weight);
    if (cell_x) *cell_x = x;
    if (cell_y) *cell_y = y;
}

void orc_densifier_finish(const float *sums, const float *counts, size_t w, size_t h, float *field)
{
    /* motion_field.rs:297-308: component_div_assign */
    for (size_t i = 0; i < 2 * w * h; i++) field[i] = sums[i] / counts[i];
}

void orc_densify(const orc_mv *mv, size_t n, size_t w, size_t h, float *field, float *counts_out)
{
    size_t cells = w * h;
    float *sums = (float *)malloc(sizeof(float) * 2 * (cells ? cells : 1));
    float *counts = counts_out ? counts_out : (float *)malloc(sizeof(float) * 2 * (cells ? cells : 1));
    orc_densifier_init(sums, counts, w, h);
    for (size_t i = 0; i < n; i++)
        orc_densifier_add(sums, counts, w, h, mv[i].px, mv[i].py, mv[i].mx, mv[i].my, 1.0f, NULL, NULL);
    orc_densifier_finish(sums, counts, w, h, field);
    free(sums);
    if (!counts_out) free(counts);
}

/* ------------------------------------------------------------------------ */
/* interpolate_empty_cells (motion_field.rs:193-294)                          */
/* The reference keeps a BTreeSet ordered by (neighbors = -#filled, idx) and  */
/* always takes the first element.  Restated with a lazy binary heap on the   */
/* same key: a popped entry is stale unless it matches the cell's current key.*/
/* ------------------------------------------------------------------------ */
typedef struct { long neighbors; size_t idx; } interp_cell;

static int interp_less(const interp_cell *a, const interp_cell *b)
{
    if (a->neighbors != b->neighbors) return a->neighbors < b->neighbors;
    return a->idx < b->idx;
}

typedef struct { interp_cell *d; size_t n, cap; } interp_heap;

static void heap_push(interp_heap *hp, interp_cell c)
{
    if (hp->n == hp->cap) { hp->cap = hp->cap ? hp->cap * 2 : 1024; hp->d = (interp_cell *)realloc(hp->d, hp->cap * sizeof(interp_cell)); }
    size_t i = hp->n++;
    hp->d[i] = c;
    while (i > 0) {
        size_t p = (i - 1) / 2;
        if (!interp_less(&hp->d[i], &hp->d[p])) break;
        interp_cell t = hp->d[i]; hp->d[i] = hp->d[p]; hp->d[p] = t;
        i = p;
    }
}

static interp_cell heap_pop(interp_heap *hp)
{
    interp_cell top = hp->d[0];
    hp->d[0] = hp->d[--hp->n];
    size_t i = 0;
    for (;;) {
        size_t l = 2 * i + 1, r = l + 1, m = i;
        if (l < hp->n && interp_less(&hp->d[l], &hp->d[m])) m = l;
        if (r < hp->n && interp_less(&hp->d[r], &hp->d[m])) m = r;
        if (m == i) break;
        interp_cell t = hp->d[i]; hp->d[i] = hp->d[m]; hp->d[m] = t;
        i = m;
    }
    return top;
}

static const int INTERP_NB[6][2] = { { -1, 0 }, { 0, -1 }, { -1, -1 }, { 1, 0 }, { 0, 1 }, { 1, 1 } }; /* :208 */

static long interp_calc_counts(const float *counts, size_t w, size_t h, size_t i)
{
    /* motion_field.rs:210-228 */
    long cnt = 0;
    long x = (long)(i % w), y = (long)(i / w);
    for (int k = 0; k < 6; k++) {
        long nx = x + INTERP_NB[k][0], ny = y + INTERP_NB[k][1];
        if (nx >= 0 && nx < (long)w && ny >= 0 && ny < (long)h && counts[2 * ((size_t)nx + (size_t)ny * w)] > 0.1f)
            cnt++;
    }
    return cnt;
}

void orc_interpolate_empty_cells(float *sums, float *counts, size_t w, size_t h)
{
    size_t cells = w * h;
    if (cells == 0) return;
    /* key[i]: current key of a queued cell, LONG_MIN when not queued */
    long *key = (long *)malloc(sizeof(long) * cells);
    interp_heap hp = { 0, 0, 0 };
    size_t queued = 0;
    for (size_t i = 0; i < cells; i++) {
        if (counts[2 * i] < 0.5f) {                            /* :235 */
            key[i] = -interp_calc_counts(counts, w, h, i);
            interp_cell c = { key[i], i };
            heap_push(&hp, c);
            queued++;
        } else {
            key[i] = 1; /* not queued (valid keys are <= 0) */
        }
    }
    if (queued == cells) { free(key); free(hp.d); return; }   /* :243-245 */

    while (hp.n > 0) {
        interp_cell cell = heap_pop(&hp);
        if (key[cell.idx] != cell.neighbors) continue;         /* stale heap entry */
        key[cell.idx] = 1;                                     /* taken out of the set (:247) */
        size_t i = cell.idx;
        long x = (long)(i % w), y = (long)(i / w);
        int added = 0;
        for (int k = 0; k < 6; k++) {                          /* :255-267 */
            long ox = INTERP_NB[k][0], oy = INTERP_NB[k][1];
            long nx = x + ox, ny = y + oy;
            if (nx >= 0 && nx < (long)w && ny >= 0 && ny < (long)h) {
                size_t idx = (size_t)nx + (size_t)ny * w;
                float cnt = counts[2 * idx];
                if (cnt > 0.1f) {
                    float scale = 1.0f - sqrtf((float)(ox * ox + oy * oy)) * 0.5f;
                    float inv_cnt = 1.0f / cnt;
                    /* `scale * inv_cnt * column`: (scale*inv_cnt) then scalar * vector */
                    float s = scale * inv_cnt;
                    densifier_add_idx(sums, counts, i, s * sums[2 * idx], s * sums[2 * idx + 1], scale);
                    added = 1;
                }
            }
        }
        if (!added) {
            /* :269-270 re-insert unchanged; cannot make progress (the reference would spin):
             * unreachable when at least one cell is filled, kept as a guard. */
            break;
        }
        for (int k = 0; k < 6; k++) {                          /* :273-289 */
            long nx = x + INTERP_NB[k][0], ny = y + INTERP_NB[k][1];
            if (nx >= 0 && nx < (long)w && ny >= 0 && ny < (long)h) {
                size_t idx = (size_t)nx + (size_t)ny * w;
                long old_key = -interp_calc_counts(counts, w, h, idx) + 1;
                if (key[idx] == old_key) {
                    key[idx] = old_key - 1;
                    interp_cell c = { key[idx], idx };
                    heap_push(&hp, c);
                }
            }
        }
    }
    free(key);
    free(hp.d);
}

void orc_flow_field(const orc_mv *mv, size_t n, size_t w, size_t h, float *field)
{
    /* flow-extract/src/main.rs:72-83 */
    size_t cells = w * h;
    float *sums = (float *)malloc(sizeof(float) * 2 * (cells ? cells : 1));
    float *counts = (float *)malloc(sizeof(float) * 2 * (cells ? cells : 1));
    orc_densifier_init(sums, counts, w, h);
    for (size_t i = 0; i < n; i++)
        orc_densifier_add(sums, counts, w, h, mv[i].px, mv[i].py, mv[i].mx, mv[i].my, 1.0f, NULL, NULL);
    orc_interpolate_empty_cells(sums, counts, w, h);
    orc_densifier_finish(sums, counts, w, h, field);
    free(sums);
    free(counts);
}

/* ======================================================================== */
/* BlockMotionDetection::detect_motion (block-motion-detector/src/lib.rs)    */
/* ======================================================================== */

size_t orc_block_dim(float min_size, size_t subdivide)
{
    /* :53-54 */
    float block_width = sqrtf(min_size) / (float)subdivide;
    return f32_as_usize(ceilf(1.0f / block_width));
}

int orc_detect_block_motion(const orc_mv *mv, size_t n, float min_size, size_t subdivide,
                            float target_motion, size_t *area_out, size_t *dim_out,
                            float *field, size_t field_cap_cells, float *mean_field)
{
    size_t dim = orc_block_dim(min_size, subdivide);
    if (dim_out) *dim_out = dim;
    if (area_out) *area_out = 0;
    if (dim == 0 || dim > 65535 || dim * dim > field_cap_cells) return -1;
    size_t cells = dim * dim;

    /* :57-61 densify with unit weights, then divide */
    float *mf = (float *)malloc(sizeof(float) * 2 * cells);
    orc_densify(mv, n, dim, dim, mf, NULL);
    if (mean_field) memcpy(mean_field, mf, sizeof(float) * 2 * cells);

    /* :63-68 map[y][x] = |mean| >= target; magnitude = sqrt(x*x + y*y), no FMA */
    unsigned char *map = (unsigned char *)calloc(cells, 1);
    for (size_t i = 0; i < cells; i++) {
        float mx = mf[2 * i], my = mf[2 * i + 1];
        float mag = sqrtf(mx * mx + my * my);
        map[i] = mag >= target_motion;
    }

    /* :70-112 flood fill, keep strictly larger island */
    size_t biggest_area = 0;
    float *best = NULL;
    float *mf2 = (float *)malloc(sizeof(float) * 2 * cells);
    size_t *stack = (size_t *)malloc(sizeof(size_t) * (cells + 1));
    for (size_t y = 0; y < dim; y++) {
        for (size_t x = 0; x < dim; x++) {
            if (!map[y * dim + x]) continue;
            size_t area = 0;
            memset(mf2, 0, sizeof(float) * 2 * cells);
            map[y * dim + x] = 0;            /* seed cleared, never set_motion'd (:80) */
            size_t sp = 0;
            stack[sp++] = y * dim + x;
            while (sp > 0) {
                size_t cur = stack[--sp];
                long cx = (long)(cur % dim), cy = (long)(cur / dim);
                area++;
                /* neighbor_offs: ox outer, oy inner (:86) */
                for (long ox = -1; ox <= 1; ox++)
                    for (long oy = -1; oy <= 1; oy++) {
                        long nx = cx + ox, ny = cy + oy;
                        if (nx < 0 || nx >= (long)dim || ny < 0 || ny >= (long)dim) continue;
                        size_t ni = (size_t)ny * dim + (size_t)nx;
                        if (map[ni]) {
                            mf2[2 * ni] = mf[2 * ni];
                            mf2[2 * ni + 1] = mf[2 * ni + 1];
                            stack[sp++] = ni;
                            map[ni] = 0;
                        }
                    }
            }
            if (area > biggest_area) {       /* :106-109 */
                biggest_area = area;
                if (!best) best = (float *)malloc(sizeof(float) * 2 * cells);
                memcpy(best, mf2, sizeof(float) * 2 * cells);
            }
        }
    }
    int some = 0;
    /* :114-118 */
    if ((float)biggest_area / (float)(dim * dim) >= min_size && best) {
        some = 1;
        memcpy(field, best, sizeof(float) * 2 * cells);
        if (area_out) *area_out = biggest_area;
    } else {
        memset(field, 0, sizeof(float) * 2 * cells);
    }
    free(mf); free(map); free(mf2); free(stack); free(best);
    return some;
}

/* ======================================================================== */
/* StandardCamera + solve_ypr_given, f32 and f64 instantiations              */
/* ======================================================================== */

#define REAL float
#define CAMERA orc_camera_f
#define SUF(n) n##_f
#define TAN tanf
#define ATAN atanf
#define SIN sinf
#define COS cosf
#define FABS fabsf
#define DEG2RAD 0.017453292519943295f   /* f32::to_radians multiplies by PI/180 rounded to f32 */
#define PI_R 3.14159265358979323846f
#include "camera_almeida.inc"
#undef REAL
#undef CAMERA
#undef SUF
#undef TAN
#undef ATAN
#undef SIN
#undef COS
#undef FABS
#undef DEG2RAD
#undef PI_R

#define REAL double
#define CAMERA orc_camera_d
#define SUF(n) n##_d
#define TAN tan
#define ATAN atan
#define SIN sin
#define COS cos
#define FABS fabs
#define DEG2RAD 0.017453292519943295
#define PI_R 3.14159265358979323846
#include "camera_almeida.inc"
#undef REAL
#undef CAMERA
#undef SUF
#undef TAN
#undef ATAN
#undef SIN
#undef COS
#undef FABS
#undef DEG2RAD
#undef PI_R

void orc_almeida_lsq_f(const orc_mv *mv, size_t n, float aspect, float fov_y_deg, float quat[4])
{
    orc_camera_f cam;
    orc_camera_new_f(&cam, aspect, fov_y_deg);
    solve_ypr_given_f(mv, NULL, n, &cam, quat);
}

void orc_almeida_lsq_d(const orc_mv *mv, size_t n, double aspect, double fov_y_deg, double quat[4])
{
    orc_camera_d cam;
    orc_camera_new_d(&cam, aspect, fov_y_deg);
    solve_ypr_given_d(mv, NULL, n, &cam, quat);
}

/* ------------------------------------------------------------------------ */
/* Seeded RNG for RANSAC.  The reference draws from rand::thread_rng() via    */
/* SliceRandom::choose_multiple (almeida:212-222): distinct elements, not     */
/* reproducible.  Contract used by the oracle AND the CUDA path: the j-th      */
/* element drawn by iteration `iter` on stream `s` (0 = the 3 model samples,  */
/* 1 = the scoring subset) is P(j), where P is a keyed pseudo-random           */
/* permutation of [0,n): a 4-round balanced Feistel network over              */
/* 2*ceil(bits(n)/2) bits with cycle walking, round function splitmix64.      */
/* Distinctness is by construction, as with choose_multiple.                  */
/* ------------------------------------------------------------------------ */
static uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

uint64_t orc_perm_index(uint64_t seed, uint64_t iter, uint64_t stream, uint64_t j, uint64_t n)
{
    if (n <= 1) return 0;
    unsigned bits = 0;
    while (((uint64_t)1 << bits) < n) bits++;
    unsigned hb = (bits + 1) / 2;
    if (hb == 0) hb = 1;
    uint64_t mask = ((uint64_t)1 << hb) - 1;
    uint64_t key = splitmix64(seed ^ splitmix64(iter * 2 + stream + 0x0F95B200ull));
    uint64_t v = j;
    do {
        uint64_t l = (v >> hb) & mask, r = v & mask;
        for (unsigned round = 0; round < 4; round++) {
            uint64_t f = splitmix64(key + ((uint64_t)round << 32) + r) & mask;
            uint64_t nl = r, nr = l ^ f;
            l = nl; r = nr;
        }
        v = (l << hb) | r;
    } while (v >= n);
    return v;
}

void orc_almeida_ransac_f(const orc_mv *mv, size_t n, float aspect, float fov_y_deg,
                          size_t num_iters, float inlier_angle_deg, size_t num_samples,
                          uint64_t seed, float quat[4], size_t *best_count_out, size_t *best_iter_out)
{
    orc_camera_f cam;
    orc_camera_new_f(&cam, aspect, fov_y_deg);
    float target_delta = inlier_angle_deg * 0.017453292519943295f;   /* :210 */
    size_t k = num_samples < n ? num_samples : n;
    size_t *inliers = (size_t *)malloc(sizeof(size_t) * (k ? k : 1));
    size_t *best_inliers = (size_t *)malloc(sizeof(size_t) * (k ? k : 1));
    size_t best_count = 0, best_iter = 0;

    for (size_t it = 0; it < num_iters; it++) {               /* :214 */
        size_t s3 = n < 3 ? n : 3;
        size_t samples[3];
        for (size_t j = 0; j < s3; j++) samples[j] = (size_t)orc_perm_index(seed, it, 0, j, n);
        float fit[4];
        solve_ypr_given_f(mv, samples, s3, &cam, fit);        /* :217 */
        /* mat = fit.inverse().to_homogeneous() (:224) */
        float inv[4] = { fit[0], -fit[1], -fit[2], -fit[3] };
        float mat[16];
        quat_to_mat4_f(inv, mat);
        size_t cnt = 0;
        for (size_t j = 0; j < k; j++) {                      /* :226-239 */
            size_t e = (size_t)orc_perm_index(seed, it, 1, j, n);
            float d[2], ang[2];
            orc_camera_delta_f(&cam, mv[e].px, mv[e].py, mat, d);
            float sx = mv[e].px + d[0], sy = mv[e].py + d[1];
            float vx = mv[e].mx - d[0], vy = mv[e].my - d[1];
            orc_camera_point_angle_f(&cam, sx, sy, ang);
            float cx = vx * cosf(ang[0]), cy = vy * cosf(ang[1]);
            if (cx * cx + cy * cy <= target_delta * target_delta) inliers[cnt++] = e;
        }
        if (cnt > best_count) {                               /* :241-243 */
            best_count = cnt; best_iter = it;
            memcpy(best_inliers, inliers, sizeof(size_t) * cnt);
        }
    }
    if (best_count >= 3) {                                    /* :246-250 */
        solve_ypr_given_f(mv, best_inliers, best_count, &cam, quat);
    } else {
        quat[0] = 1; quat[1] = quat[2] = quat[3] = 0;
    }
    if (best_count_out) *best_count_out = best_count;
    if (best_iter_out) *best_iter_out = best_iter;
    free(inliers);
    free(best_inliers);
}

/* calc_view (almeida:280-286) with eye = origin: Matrix4::look_at_rh(0, rot*(0,-1,0), rot*(0,0,1)) */
void orc_look_at_rh_view_d(const double q[4], double view[16])
{
    double r[16];
    quat_to_mat4_d(q, r);
    /* dir = R*(0,-1,0), up = R*(0,0,1) */
    double dir[3] = { -r[1], -r[5], -r[9] };
    double up[3] = { r[2], r[6], r[10] };
    /* Rotation3::look_at_rh(dir, up) = face_towards(-dir, up).inverse() */
    double z[3] = { -dir[0], -dir[1], -dir[2] };
    double zn = sqrt(z[0] * z[0] + z[1] * z[1] + z[2] * z[2]);
    for (int i = 0; i < 3; i++) z[i] /= zn;
    double x[3] = { up[1] * z[2] - up[2] * z[1], up[2] * z[0] - up[0] * z[2], up[0] * z[1] - up[1] * z[0] };
    double xn = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    for (int i = 0; i < 3; i++) x[i] /= xn;
    double y[3] = { z[1] * x[2] - z[2] * x[1], z[2] * x[0] - z[0] * x[2], z[0] * x[1] - z[1] * x[0] };
    double yn = sqrt(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]);
    for (int i = 0; i < 3; i++) y[i] /= yn;
    /* face_towards has columns [x y z]; the inverse (transpose) has them as rows */
    for (int i = 0; i < 16; i++) view[i] = 0;
    for (int c = 0; c < 3; c++) { view[0 * 4 + c] = x[c]; view[1 * 4 + c] = y[c]; view[2 * 4 + c] = z[c]; }
    view[15] = 1;
}

double orc_quat_angle_to_d(const double a[4], const double b[4])
{
    /* (a.inverse() * b).angle() = 2*atan2(|imag|, |w|) */
    double ia[4] = { a[0], -a[1], -a[2], -a[3] }, d[4];
    quat_mul_d(ia, b, d);
    double im = sqrt(d[1] * d[1] + d[2] * d[2] + d[3] * d[3]);
    return 2.0 * atan2(im, fabs(d[0]));
}

/* ======================================================================== */
/* Exhaustive block matcher (specification; see header)                      */
/* ======================================================================== */

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static inline uint32_t block_cost_plain(const uint8_t *c, const uint8_t *p, int stride, int block, int metric)
{
    uint32_t acc = 0;
    for (int y = 0; y < block; y++) {
        const uint8_t *cr = c + (size_t)y * stride, *pr = p + (size_t)y * stride;
        for (int x = 0; x < block; x++) {
            int d = (int)cr[x] - (int)pr[x];
            acc += metric == 0 ? (uint32_t)(d < 0 ? -d : d) : (uint32_t)(d * d);
        }
    }
    return acc;
}

static inline uint32_t block_cost_fast(const uint8_t *c, const uint8_t *p, int stride, int block, int metric)
{
#if defined(__SSE2__)
    if (metric == 0 && block == 16) {
        __m128i acc = _mm_setzero_si128();
        for (int y = 0; y < 16; y++) {
            __m128i a = _mm_loadu_si128((const __m128i *)(c + (size_t)y * stride));
            __m128i b = _mm_loadu_si128((const __m128i *)(p + (size_t)y * stride));
            acc = _mm_add_epi64(acc, _mm_sad_epu8(a, b));
        }
        return (uint32_t)(_mm_cvtsi128_si32(acc) + _mm_cvtsi128_si32(_mm_srli_si128(acc, 8)));
    }
    if (metric == 0 && block == 8) {
        __m128i acc = _mm_setzero_si128();
        for (int y = 0; y < 8; y++) {
            __m128i a = _mm_loadl_epi64((const __m128i *)(c + (size_t)y * stride));
            __m128i b = _mm_loadl_epi64((const __m128i *)(p + (size_t)y * stride));
            acc = _mm_add_epi64(acc, _mm_sad_epu8(a, b));
        }
        return (uint32_t)_mm_cvtsi128_si32(acc);
    }
#endif
    return block_cost_plain(c, p, stride, block, metric);
}

static void match_one_block(const uint8_t *prev, const uint8_t *cur, int w, int h, int stride,
                            int block, int range, int metric, int bx, int by, int fast,
                            int16_t *mv_xy, uint32_t *cost_out, orc_mv *entry)
{
    const int x0 = bx * block, y0 = by * block;
    const uint8_t *c = cur + (size_t)y0 * stride + x0;
    uint64_t best = UINT64_MAX;
    for (int dy = -range; dy <= range; dy++) {
        int py = y0 + dy;
        if (py < 0 || py + block > h) continue;
        for (int dx = -range; dx <= range; dx++) {
            int px = x0 + dx;
            if (px < 0 || px + block > w) continue;
            const uint8_t *p = prev + (size_t)py * stride + px;
            uint32_t cost = fast ? block_cost_fast(c, p, stride, block, metric)
                                 : block_cost_plain(c, p, stride, block, metric);
            uint64_t key = ((uint64_t)cost << 27) | ((uint64_t)(dx * dx + dy * dy) << 14) |
                           ((uint64_t)(dy + range) << 7) | (uint64_t)(dx + range);
            if (key < best) best = key;
        }
    }
    int dx = (int)(best & 127) - range, dy = (int)((best >> 7) & 127) - range;
    if (mv_xy) { mv_xy[0] = (int16_t)dx; mv_xy[1] = (int16_t)dy; }
    if (cost_out) *cost_out = (uint32_t)(best >> 27);
    if (entry) {
        /* av-decoder/src/lib.rs:404-419: frame_norm = (1/W, 1/H); pos = src * frame_norm;
         * motion = (motion/scale) * -frame_norm, with dst = block centre, src = dst + (dx,dy),
         * motion_scale = 1 */
        float nx = 1.0f / (float)w, ny = 1.0f / (float)h;
        int src_x = x0 + block / 2 + dx, src_y = y0 + block / 2 + dy;
        entry->px = (float)src_x * nx;
        entry->py = (float)src_y * ny;
        entry->mx = ((float)dx / 1.0f) * -nx;
        entry->my = ((float)dy / 1.0f) * -ny;
    }
}

static long block_match_impl(const uint8_t *prev, const uint8_t *cur, int w, int h, int stride,
                             int block, int range, int metric, int16_t *mv_xy, uint32_t *cost,
                             orc_mv *entries, int threads, int fast)
{
    if (!prev || !cur || w <= 0 || h <= 0 || stride < w || block <= 0 || range < 0 || range > 63 ||
        (metric != 0 && metric != 1) || block > 64)
        return -1;
    const int nbx = w / block, nby = h / block;
    const long nblocks = (long)nbx * nby;
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads > 1 ? threads : 1)
#endif
    for (long i = 0; i < nblocks; i++) {
        int bx = (int)(i % nbx), by = (int)(i / nbx);
        match_one_block(prev, cur, w, h, stride, block, range, metric, bx, by, fast,
                        mv_xy ? mv_xy + 2 * i : NULL, cost ? cost + i : NULL, entries ? entries + i : NULL);
    }
    return nblocks;
}

long orc_block_match(const uint8_t *prev, const uint8_t *cur, int w, int h, int stride,
                     int block, int range, int metric,
                     int16_t *mv_xy, uint32_t *cost, orc_mv *entries, int threads)
{
    return block_match_impl(prev, cur, w, h, stride, block, range, metric, mv_xy, cost, entries, threads, 0);
}

long orc_block_match_fast(const uint8_t *prev, const uint8_t *cur, int w, int h, int stride,
                          int block, int range, int metric,
                          int16_t *mv_xy, uint32_t *cost, orc_mv *entries, int threads)
{
    return block_match_impl(prev, cur, w, h, stride, block, range, metric, mv_xy, cost, entries, threads, 1);
}
