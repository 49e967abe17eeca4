// K5 / K6: AlmeidaEstimator — iterated 3-parameter least-squares camera rotation fit
// (almeida-estimator/src/lib.rs:123-200, solve_ypr_given) and its RANSAC wrapper (:202-251),
// over StandardCamera::delta / point_angle (ofps/src/camera.rs:45-117, 120-161).
//
// Arithmetic contract.  Everything that happens PER ENTRY (un-project, rotate, project, the
// divide by NDC z, the four deltas, the twelve products) is evaluated in f32 in the reference's
// operation order with no FMA contraction (the library is built with --fmad=false), i.e. the
// per-entry terms are the ones the CPU restatement produces.  What differs is the REDUCTION:
// the reference folds N terms sequentially in f32 (almeida:126-130); here each thread
// accumulates its terms in f64, blocks combine in a fixed order, and the 3x3 system is rounded
// to f32 before the reference's f32 LU.  That is deterministic (independent of block scheduling)
// and closer to the exact sum than the sequential fold, so the quaternion is compared with a
// tolerance (1e-4, see tests) instead of bit-for-bit.
//
// Work saving that does not change results: the three prototype fields (roll/pitch/yaw deltas,
// almeida:30-47, 151-153) do not depend on the iteration, so the 3x3 normal matrix A is
// accumulated once, and later iterations only accumulate the right-hand side b.
//
// Camera matrices and the prototype rotations are built on the host in f32 (tanf/sinf/cosf of
// the host libm, as the CPU path does) and passed by value.
#include "common.cuh"

#include <cooperative_groups.h>

#include <cmath>

namespace ofpsb {

namespace {

constexpr int LSQ_NT = 256;
constexpr int LSQ_ITERS = 30;          // ceil(15 / ALPHA), almeida:132
constexpr float ALPHA = 0.5f;          // almeida:18

struct AlmeidaConst {
    float unproj[16];   // view^T * inv_proj, row-major (camera.rs:54 with rotate()'s view, :91-96)
    float proj[4];      // m00, m11, m22, m23 of Perspective3
    float roll[9], pitch[9], yaw[9];   // 3x3 parts of the prototype rotations (almeida:33,37,41)
    float fx, fy;       // intrinsics focal lengths (camera.rs:120-129)
    float eps_r;        // EPS = 0.001 deg in radians (almeida:17)
};

struct AlmeidaState {
    float rotation[4];   // running estimate (w,i,j,k)
    float a[9];          // normal matrix, accumulated at iteration 0
    unsigned int ticket;
    unsigned int epoch;   // persistent kernel: iterations finished (grid barrier between the 30 dependent iterations)
};

// ---- camera (f32, reference operation order; see oracle/camera_almeida.inc)
__device__ __forceinline__ void unproject_world(const AlmeidaConst& c, float x, float y, float w[3])
{
    const float p0 = x * 2.0f - 1.0f, p1 = y * 2.0f - 1.0f, p2 = 1.0f;
    const float* m = c.unproj;
    float t[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        float acc = m[i * 4 + 0] * p0;
        acc = acc + m[i * 4 + 1] * p1;
        acc = acc + m[i * 4 + 2] * p2;
        t[i] = acc + m[i * 4 + 3];
    }
    float n = (m[12] * p0 + m[13] * p1) + m[14] * p2;
    n = n + m[15];
    if (n != 0.0f) {
        w[0] = t[0] / n; w[1] = t[1] / n; w[2] = t[2] / n;
    } else {
        w[0] = t[0]; w[1] = t[1]; w[2] = t[2];
    }
}

// rotate the world point by a pure rotation (homogeneous last row 0,0,0,1 -> divide by 1), apply
// rotate()'s fixed view, project, divide by NDC z, back to [0,1], subtract the start (camera.rs:72-117)
__device__ __forceinline__ void delta_from_world(const AlmeidaConst& c, const float w[3], const float* r, float x,
                                                 float y, float out[2])
{
    float rw[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        float acc = r[i * 3 + 0] * w[0];
        acc = acc + r[i * 3 + 1] * w[1];
        acc = acc + r[i * 3 + 2] * w[2];
        rw[i] = acc + 0.0f;
    }
    const float v0 = -rw[0], v1 = rw[2], v2 = rw[1];     // view = rows (-1,0,0),(0,0,1),(0,1,0)
    const float inv_denom = -1.0f / v2;
    const float sx = c.proj[0] * v0 * inv_denom;
    const float sy = c.proj[1] * v1 * inv_denom;
    const float sz = (c.proj[2] * v2 + c.proj[3]) * inv_denom;
    const float qx = sx / sz, qy = sy / sz;
    out[0] = (qx + 1.0f) * 0.5f - x;
    out[1] = (qy + 1.0f) * 0.5f - y;
}

__device__ __forceinline__ void quat_to_mat3(const float q[4], float m[9])
{
    const float w = q[0], i = q[1], j = q[2], k = q[3];
    const float ww = w * w, ii = i * i, jj = j * j, kk = k * k;
    const float ij = i * j * 2.0f, wk = w * k * 2.0f, wj = w * j * 2.0f;
    const float ik = i * k * 2.0f, jk = j * k * 2.0f, wi = w * i * 2.0f;
    m[0] = ww + ii - jj - kk; m[1] = ij - wk;           m[2] = wj + ik;
    m[3] = wk + ij;           m[4] = ww - ii + jj - kk; m[5] = jk - wi;
    m[6] = ik - wj;           m[7] = wi + jk;           m[8] = ww - ii - jj + kk;
}

__device__ __forceinline__ void quat_from_euler(float roll, float pitch, float yaw, float q[4])
{
    const float sr = sinf(roll * 0.5f), cr = cosf(roll * 0.5f);
    const float sp = sinf(pitch * 0.5f), cp = cosf(pitch * 0.5f);
    const float sy = sinf(yaw * 0.5f), cy = cosf(yaw * 0.5f);
    q[0] = cr * cp * cy + sr * sp * sy;
    q[1] = sr * cp * cy - cr * sp * sy;
    q[2] = cr * sp * cy + sr * cp * sy;
    q[3] = cr * cp * sy - sr * sp * cy;
}

// from_euler_angles on half-angle sines / cosines (same expression tree as quat_from_euler: literal 0 / 1 arguments fold the
// same way)
__device__ __forceinline__ void quat_from_sincos(float sr, float cr, float sp, float cp, float sy, float cy, float q[4])
{
    q[0] = cr * cp * cy + sr * sp * sy;
    q[1] = sr * cp * cy - cr * sp * sy;
    q[2] = cr * sp * cy + sr * cp * sy;
    q[3] = cr * cp * sy - sr * sp * cy;
}

__device__ __forceinline__ void quat_mul(const float a[4], const float b[4], float o[4])
{
    const float w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    const float i = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    const float j = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    const float k = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
    o[0] = w; o[1] = i; o[2] = j; o[3] = k;
}

// Matrix3::lu() with partial pivoting + solve; false when a pivot is zero (almeida:181-183).  Every index is a
// compile-time constant after unrolling (row swaps are conditional swaps of fixed rows, the right-hand side is swapped with
// the rows instead of being permuted at the end), so the matrix stays in registers: the solver step is one thread's
// dependent chain between two barriers of every iteration, and the local-memory version cost ~1 us of it.
__device__ __forceinline__ bool lu3_solve(const float a_in[9], const float b_in[3], float x[3])
{
    float a[3][3], y[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        y[r] = b_in[r];
#pragma unroll
        for (int c = 0; c < 3; c++) a[r][c] = a_in[r * 3 + c];
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
        int piv = i;
        float best = fabsf(a[i][i]), pv = a[i][i];
#pragma unroll
        for (int r = i + 1; r < 3; r++) {
            const float v = fabsf(a[r][i]);
            if (v > best) { best = v; piv = r; pv = a[r][i]; }
        }
        if (pv != 0.0f) {
#pragma unroll
            for (int r = i + 1; r < 3; r++)
                if (piv == r) {
#pragma unroll
                    for (int c = 0; c < 3; c++) { const float t = a[i][c]; a[i][c] = a[r][c]; a[r][c] = t; }
                    const float t = y[i]; y[i] = y[r]; y[r] = t;
                }
            const float inv_diag = 1.0f / a[i][i];
#pragma unroll
            for (int r = i + 1; r < 3; r++) a[r][i] = a[r][i] * inv_diag;
#pragma unroll
            for (int c = i + 1; c < 3; c++) {
                const float pr = a[i][c];
#pragma unroll
                for (int r = i + 1; r < 3; r++) a[r][c] = a[r][c] + (-pr) * a[r][i];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const float coeff = y[i];
#pragma unroll
        for (int r = i + 1; r < 3; r++) y[r] = y[r] + (-coeff) * a[r][i];
    }
#pragma unroll
    for (int i = 2; i >= 0; i--) {
        const float d = a[i][i];
        if (d == 0.0f) return false;
        y[i] = y[i] / d;
#pragma unroll
        for (int r = 0; r < i; r++) y[r] = y[r] + (-y[i]) * a[r][i];
    }
    x[0] = y[0]; x[1] = y[1]; x[2] = y[2];
    return true;
}

// one solver step from the reduced system (almeida:181-195)
__device__ void lsq_step(const float a[9], const float b[3], float eps_r, int it, float rotation[4])
{
    const float alpha = (it == LSQ_ITERS - 1) ? 1.0f : ALPHA;   // almeida:138
    float model[3];
    if (!lu3_solve(a, b, model)) model[0] = model[1] = model[2] = 0.0f;
    for (int k = 0; k < 3; k++) model[k] = model[k] * eps_r * alpha;
    float roll[4], pitch[4], yaw[4], pr[4], rot[4], nr[4];
    quat_from_euler(0.0f, model[0], 0.0f, roll);
    quat_from_euler(model[1], 0.0f, 0.0f, pitch);
    quat_from_euler(0.0f, 0.0f, -model[2], yaw);
    quat_mul(pitch, roll, pr);
    quat_mul(pr, yaw, rot);
    quat_mul(rotation, rot, nr);
    for (int k = 0; k < 4; k++) rotation[k] = nr[k];
}

// The same step taken by a whole warp (all 32 lanes call it converged; a, b, rotation are lane 0's): lane 0 solves the 3x3
// system, lanes 0-2 take the sine and cosine of one angle each — six libm calls in a row are most of the scalar step —
// and lane 0 composes the rotation.  Same operations on the same values as lsq_step.
__device__ __forceinline__ void lsq_step_warp(const float a[9], const float b[3], float eps_r, int it, float rotation[4], int lane)
{
    const float alpha = (it == LSQ_ITERS - 1) ? 1.0f : ALPHA;   // almeida:138
    float model[3] = {0.0f, 0.0f, 0.0f};
    if (lane == 0) {
        if (!lu3_solve(a, b, model)) model[0] = model[1] = model[2] = 0.0f;
        for (int k = 0; k < 3; k++) model[k] = model[k] * eps_r * alpha;
    }
    const float m0 = __shfl_sync(0xffffffffu, model[0], 0), m1 = __shfl_sync(0xffffffffu, model[1], 0);
    const float m2 = __shfl_sync(0xffffffffu, model[2], 0);
    const float ang = lane == 0 ? m0 : lane == 1 ? m1 : -m2;
    const float sn = sinf(ang * 0.5f), cs = cosf(ang * 0.5f);
    const float s0 = __shfl_sync(0xffffffffu, sn, 0), c0 = __shfl_sync(0xffffffffu, cs, 0);
    const float s1 = __shfl_sync(0xffffffffu, sn, 1), c1 = __shfl_sync(0xffffffffu, cs, 1);
    const float s2 = __shfl_sync(0xffffffffu, sn, 2), c2 = __shfl_sync(0xffffffffu, cs, 2);
    if (lane == 0) {
        float roll[4], pitch[4], yaw[4], pr[4], rot[4], nr[4];
        quat_from_sincos(0.0f, 1.0f, s0, c0, 0.0f, 1.0f, roll);    // quat_from_euler(0, model[0], 0)
        quat_from_sincos(s1, c1, 0.0f, 1.0f, 0.0f, 1.0f, pitch);   // quat_from_euler(model[1], 0, 0)
        quat_from_sincos(0.0f, 1.0f, 0.0f, 1.0f, s2, c2, yaw);     // quat_from_euler(0, 0, -model[2])
        quat_mul(pitch, roll, pr);
        quat_mul(pr, yaw, rot);
        quat_mul(rotation, rot, nr);
        for (int k = 0; k < 4; k++) rotation[k] = nr[k];
    }
}

// the four vectors of one entry (almeida:142-157)
__device__ __forceinline__ void entry_vectors(const AlmeidaConst& c, const float4 e, const float rotm[9], bool protos,
                                              float v[4][2])
{
    float w[3], d[2];
    unproject_world(c, e.x, e.y, w);
    delta_from_world(c, w, rotm, e.x, e.y, d);
    v[0][0] = e.z - d[0];
    v[0][1] = e.w - d[1];
    delta_from_world(c, w, c.roll, e.x, e.y, v[1]);
    delta_from_world(c, w, c.pitch, e.x, e.y, v[2]);
    delta_from_world(c, w, c.yaw, e.x, e.y, v[3]);
    (void)protos;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// Sum of the per-block partials of components k0..11 -> red[0][k].  16 lanes per component walk the blocks with stride
// 16 in ascending order, then a 4-step xor tree: a fixed order (same bits in every CTA and in every launch mode), and a
// few dozen dependent adds instead of gridDim.x.
__device__ __forceinline__ void combine_partials(const double* __restrict__ buf, unsigned grid, int k0, double (*red)[12], int tid)
{
    const int k = tid >> 4, l = tid & 15;
    if (k >= 12) return;                       // warps 6, 7 (LSQ_NT = 256): no component
    double s = 0.0;
    if (k >= k0)
        for (unsigned b = (unsigned)l; b < grid; b += 16) s += __ldcg(&buf[(size_t)b * 12 + k]);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (l == 0 && k >= k0) red[0][k] = s;
}

// ------------------------------------------------------------------ least-squares kernel
// MODE 0 (SINGLE): one CTA runs all 30 iterations (small inputs: RANSAC refits).
// MODE 2 (persistent): ONE launch of a co-resident grid runs all 30 iterations; per-block f64 partials are combined by
//   the last block to arrive, in block order, and the other CTAs wait for its `epoch` store (a grid barrier costs
//   ~2 us; 30 separate launches cost ~10 us each — the reference's own 12,600-vector fields took 0.35 ms of launches
//   for ~30 us of arithmetic).
// MODE 1: the same grid, one launch per iteration (fallback when a cooperative launch is not possible).
constexpr int LSQ_SINGLE = 0, LSQ_STEPWISE = 1, LSQ_PERSISTENT = 2;
template <int MODE>
__global__ void __launch_bounds__(LSQ_NT) almeida_lsq_kernel(const ofps_mv* __restrict__ entries,
                                                             const uint32_t* __restrict__ idx, size_t n_arg,
                                                             const uint32_t* __restrict__ n_ptr, const AlmeidaConst cst,
                                                             AlmeidaState* __restrict__ state,
                                                             double* __restrict__ partial, float* __restrict__ out_quat,
                                                             int it_arg)
{
    __shared__ double red[LSQ_NT / 32][12];
    __shared__ float s_rot[4];
    __shared__ float s_a[9];
    __shared__ bool s_last;
    constexpr bool SINGLE = MODE == LSQ_SINGLE, PERSIST = MODE == LSQ_PERSISTENT;
    const int tid = threadIdx.x;
    const size_t n = n_ptr ? (size_t)*n_ptr : n_arg;
    const float4* e4 = reinterpret_cast<const float4*>(entries);

    if (n_ptr && n < 3) {   // solve_ypr_ransac: fewer than 3 inliers -> identity (almeida:246-250)
        if (blockIdx.x == 0 && tid == 0 && (SINGLE || PERSIST || it_arg == LSQ_ITERS - 1)) {
            out_quat[0] = 1.0f; out_quat[1] = 0.0f; out_quat[2] = 0.0f; out_quat[3] = 0.0f;
        }
        return;
    }
    if (tid < 4) s_rot[tid] = (SINGLE || PERSIST || it_arg == 0) ? (tid == 0 ? 1.0f : 0.0f) : state->rotation[tid];
    if (MODE == LSQ_STEPWISE && it_arg > 0 && tid < 9) s_a[tid] = state->a[tid];
    __syncthreads();

    const int it_begin = MODE == LSQ_STEPWISE ? it_arg : 0;
    const int it_end = MODE == LSQ_STEPWISE ? it_arg + 1 : LSQ_ITERS;
    for (int it = it_begin; it < it_end; it++) {
        float rot[4] = {s_rot[0], s_rot[1], s_rot[2], s_rot[3]};
        float rotm[9];
        quat_to_mat3(rot, rotm);
        const bool first = it == 0;
        double acc[12];
#pragma unroll
        for (int k = 0; k < 12; k++) acc[k] = 0.0;
        for (size_t i = (size_t)blockIdx.x * LSQ_NT + tid; i < n; i += (size_t)gridDim.x * LSQ_NT) {
            const float4 e = __ldg(e4 + (idx ? idx[i] : i));
            float v[4][2];
            entry_vectors(cst, e, rotm, first, v);
            if (first) {
#pragma unroll
                for (int r = 0; r < 3; r++)
#pragma unroll
                    for (int c = 0; c < 3; c++)
                        acc[r * 3 + c] += (double)(v[r + 1][0] * v[c + 1][0] + v[r + 1][1] * v[c + 1][1]);
            }
#pragma unroll
            for (int r = 0; r < 3; r++) acc[9 + r] += (double)(v[r + 1][0] * v[0][0] + v[r + 1][1] * v[0][1]);
        }
        // block reduction (fixed order: lanes by shuffle tree, warps sequentially)
        const int k0 = first ? 0 : 9;
        for (int k = k0; k < 12; k++) {
            const double s = warp_sum(acc[k]);
            if ((tid & 31) == 0) red[tid >> 5][k] = s;
        }
        __syncthreads();
        if (SINGLE) {
            if (tid < 32) {   // warp 0 takes the step (the values are lane 0's)
                float a[9] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f}, b[3] = {0.0f, 0.0f, 0.0f};
                if (tid == 0) {
                    for (int k = 0; k < 12; k++) {
                        if (k < k0) continue;
                        double s = 0.0;
                        for (int w = 0; w < LSQ_NT / 32; w++) s += red[w][k];
                        if (k < 9) s_a[k] = (float)s;
                        else b[k - 9] = (float)s;
                    }
                    for (int k = 0; k < 9; k++) a[k] = s_a[k];
                }
                float r4[4] = {s_rot[0], s_rot[1], s_rot[2], s_rot[3]};
                __syncwarp();   // every lane has read s_rot before lane 0 rewrites it
                lsq_step_warp(a, b, cst.eps_r, it, r4, tid);
                if (tid == 0)
                    for (int k = 0; k < 4; k++) s_rot[k] = r4[k];
            }
            __syncthreads();
        } else if (PERSIST) {
            // Every CTA publishes its partials, waits until all have (one monotonic counter, no reset), then combines ALL
            // partials itself — in block order, so every CTA computes the same bits — and takes the solver step itself.
            // One hop through L2 per iteration instead of three (ticket -> last block's result -> everybody reads it).
            // Partials are double-buffered by iteration parity: a CTA can be one iteration ahead of a slow reader.
            double* buf = partial + (size_t)(it & 1) * gridDim.x * 12;
            if (tid < 12 && tid >= k0) {
                double s = 0.0;
                for (int w = 0; w < LSQ_NT / 32; w++) s += red[w][tid];
                buf[(size_t)blockIdx.x * 12 + tid] = s;
            }
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                atomicAdd(&state->ticket, 1u);
                const unsigned target = (unsigned)(it + 1) * gridDim.x;
                while (*reinterpret_cast<volatile unsigned*>(&state->ticket) < target) __nanosleep(32);
                __threadfence();
            }
            __syncthreads();
            combine_partials(buf, gridDim.x, k0, red, tid);
            __syncthreads();
            if (tid < 32) {
                float a[9], b[3];
                for (int k = 0; k < 9; k++) a[k] = first ? (float)red[0][k] : s_a[k];
                for (int k = 0; k < 3; k++) b[k] = (float)red[0][9 + k];
                float r4[4] = {s_rot[0], s_rot[1], s_rot[2], s_rot[3]};
                __syncwarp();   // every lane has read s_a / s_rot before lane 0 rewrites them
                lsq_step_warp(a, b, cst.eps_r, it, r4, tid);
                if (tid == 0) {
                    for (int k = 0; k < 4; k++) s_rot[k] = r4[k];
                    if (first) for (int k = 0; k < 9; k++) s_a[k] = a[k];
                }
            }
            s_last = blockIdx.x == 0;
            __syncthreads();
        } else {
            if (tid < 12 && tid >= k0) {
                double s = 0.0;
                for (int w = 0; w < LSQ_NT / 32; w++) s += red[w][tid];
                partial[(size_t)blockIdx.x * 12 + tid] = s;
            }
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                const unsigned t = atomicAdd(&state->ticket, 1u);
                s_last = t == gridDim.x - 1;
            }
            __syncthreads();
            if (s_last) {
                __threadfence();
                __syncthreads();   // red[][] of the block reduction has been consumed
                combine_partials(partial, gridDim.x, k0, red, tid);
                __syncthreads();
                if (tid < 32) {
                    float a[9], b[3];
                    for (int k = 0; k < 9; k++) a[k] = first ? (float)red[0][k] : s_a[k];
                    for (int k = 0; k < 3; k++) b[k] = (float)red[0][9 + k];
                    float r4[4] = {s_rot[0], s_rot[1], s_rot[2], s_rot[3]};
                    __syncwarp();
                    lsq_step_warp(a, b, cst.eps_r, it, r4, tid);
                    if (tid == 0) {
                        for (int k = 0; k < 4; k++) { state->rotation[k] = r4[k]; s_rot[k] = r4[k]; }
                        if (first) for (int k = 0; k < 9; k++) state->a[k] = a[k];
                        state->ticket = 0;
                    }
                }
                __syncthreads();
            }
        }
    }
    // rotation.inverse() (almeida:199)
    if ((SINGLE || (s_last && (PERSIST || it_arg == LSQ_ITERS - 1))) && tid == 0) {
        out_quat[0] = s_rot[0]; out_quat[1] = -s_rot[1]; out_quat[2] = -s_rot[2]; out_quat[3] = -s_rot[3];
    }
}

// ------------------------------------------------------------------ least-squares kernel, one thread-block cluster
// The reference's own field sizes (a few thousand to 12,600 vectors) are latency-bound in the grid version above: ~6 us
// per iteration, of which ~0.3 us is arithmetic and the rest is the trip of the partial sums through L2 (publish, fence,
// ticket, poll, re-read).  Here ONE cluster of up to 16 CTAs x 1024 threads holds one entry per thread for all 30
// iterations: the un-projected point and the three prototype vectors of the entry — iteration-invariant, recomputed
// every iteration by the reference and by the kernels above — stay in registers, so an iteration is one delta() and
// three dot products per thread; the per-CTA sums go straight into every CTA's shared memory (DSMEM stores), the
// hardware cluster barrier replaces the L2 ticket, and every CTA takes the identical solver step.  Partials are
// double-buffered by iteration parity (a CTA can be one iteration ahead of a slow reader, never two: the next barrier
// needs everybody).  Sums are f64 in a fixed order (lanes by shuffle tree, warps, then CTAs ascending).
constexpr int CL_NT = 1024;
constexpr int CL_MAX_CTAS = 16;

__global__ void __launch_bounds__(CL_NT, 1) almeida_lsq_cluster_kernel(const ofps_mv* __restrict__ entries,
                                                                      const uint32_t* __restrict__ idx, size_t n_arg,
                                                                      const uint32_t* __restrict__ n_ptr,
                                                                      const AlmeidaConst cst, float* __restrict__ out_quat)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ double red[CL_NT / 32][12];
    __shared__ double part[2][CL_MAX_CTAS][12];   // [parity][CTA of the cluster][component], written by the owners
    __shared__ float s_rot[4];
    __shared__ float s_a[9];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned rank = cluster.block_rank(), ncta = cluster.num_blocks();
    const size_t n = n_ptr ? (size_t)*n_ptr : n_arg;
    if (n_ptr && n < 3) {   // solve_ypr_ransac: fewer than 3 inliers -> identity (almeida:246-250); cluster-uniform
        if (rank == 0 && tid == 0) { out_quat[0] = 1.0f; out_quat[1] = 0.0f; out_quat[2] = 0.0f; out_quat[3] = 0.0f; }
        return;
    }
    if (tid < 4) s_rot[tid] = tid == 0 ? 1.0f : 0.0f;
    // this thread's entry: world point and prototypes once (almeida:151-153 recomputes them every iteration)
    const size_t i = (size_t)rank * blockDim.x + tid;
    const bool have = i < n;
    float4 e = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    float w[3] = {0.0f, 0.0f, 1.0f}, p1[2] = {0.0f, 0.0f}, p2[2] = {0.0f, 0.0f}, p3[2] = {0.0f, 0.0f};
    if (have) {
        e = __ldg(reinterpret_cast<const float4*>(entries) + (idx ? idx[i] : i));
        unproject_world(cst, e.x, e.y, w);
        delta_from_world(cst, w, cst.roll, e.x, e.y, p1);
        delta_from_world(cst, w, cst.pitch, e.x, e.y, p2);
        delta_from_world(cst, w, cst.yaw, e.x, e.y, p3);
    }
    cluster.sync();   // every CTA of the cluster is running before anyone stores into its shared memory (also the CTA barrier)

    for (int it = 0; it < LSQ_ITERS; it++) {
        const bool first = it == 0;
        const int k0 = first ? 0 : 9;
        double acc[12];
#pragma unroll
        for (int k = 0; k < 12; k++) acc[k] = 0.0;
        if (have) {
            const float rot[4] = {s_rot[0], s_rot[1], s_rot[2], s_rot[3]};
            float rotm[9], d[2];
            quat_to_mat3(rot, rotm);
            delta_from_world(cst, w, rotm, e.x, e.y, d);
            const float v0[2] = {e.z - d[0], e.w - d[1]};
            const float* pv[3] = {p1, p2, p3};
            if (first) {
#pragma unroll
                for (int r = 0; r < 3; r++)
#pragma unroll
                    for (int c = 0; c < 3; c++) acc[r * 3 + c] = (double)(pv[r][0] * pv[c][0] + pv[r][1] * pv[c][1]);
            }
#pragma unroll
            for (int r = 0; r < 3; r++) acc[9 + r] = (double)(pv[r][0] * v0[0] + pv[r][1] * v0[1]);
        }
        for (int k = k0; k < 12; k++) {
            const double sum = warp_sum(acc[k]);
            if (lane == 0) red[warp][k] = sum;
        }
        __syncthreads();
        // CTA total of component k (thread k), stored into every CTA's part[parity][rank][k]
        if (tid < 12 && tid >= k0) {
            double sum = 0.0;
            for (int wv = 0; wv < (int)(blockDim.x >> 5); wv++) sum += red[wv][tid];
            for (unsigned r = 0; r < ncta; r++) *cluster.map_shared_rank(&part[it & 1][rank][tid], r) = sum;
        }
        cluster.sync();   // release / acquire across the cluster: every CTA's partials have landed everywhere
        if (tid < 32) {
            // lanes k0 .. 11 add up the CTA partials of one component each (CTAs ascending), lane 0 collects them
            double sum = 0.0;
            if (tid >= k0 && tid < 12)
                for (unsigned r = 0; r < ncta; r++) sum += part[it & 1][r][tid];
            float a[9], b[3];
            const float fs = (float)sum;
#pragma unroll
            for (int k = 0; k < 9; k++) {
                const float v = __shfl_sync(0xffffffffu, fs, k);
                a[k] = first ? v : s_a[k];
            }
#pragma unroll
            for (int k = 0; k < 3; k++) b[k] = __shfl_sync(0xffffffffu, fs, 9 + k);
            float r4[4] = {s_rot[0], s_rot[1], s_rot[2], s_rot[3]};
            __syncwarp();
            lsq_step_warp(a, b, cst.eps_r, it, r4, tid);
            if (tid == 0) {
                for (int k = 0; k < 4; k++) s_rot[k] = r4[k];
                if (first) for (int k = 0; k < 9; k++) s_a[k] = a[k];
            }
        }
        __syncthreads();
    }
    cluster.sync();   // no CTA exits while another may still store into its shared memory
    if (rank == 0 && tid == 0) {   // rotation.inverse() (almeida:199)
        out_quat[0] = s_rot[0]; out_quat[1] = -s_rot[1]; out_quat[2] = -s_rot[2]; out_quat[3] = -s_rot[3];
    }
}

// ------------------------------------------------------------------ seeded sampling (K6)
// Same keyed permutation as oracle/ofps_oracle.c (orc_perm_index): 4-round balanced Feistel with
// cycle walking, round function splitmix64.  Stands in for rand::thread_rng +
// SliceRandom::choose_multiple (almeida:212-222), which are not reproducible.
__host__ __device__ __forceinline__ unsigned long long splitmix64(unsigned long long x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__device__ unsigned long long perm_index(unsigned long long seed, unsigned long long iter, unsigned long long stream,
                                         unsigned long long j, unsigned long long n)
{
    if (n <= 1) return 0;
    unsigned bits = 0;
    while (((unsigned long long)1 << bits) < n) bits++;
    unsigned hb = (bits + 1) / 2;
    if (hb == 0) hb = 1;
    const unsigned long long mask = ((unsigned long long)1 << hb) - 1;
    const unsigned long long key = splitmix64(seed ^ splitmix64(iter * 2 + stream + 0x0F95B200ull));
    unsigned long long v = j;
    do {
        unsigned long long l = (v >> hb) & mask, r = v & mask;
        for (unsigned round = 0; round < 4; round++) {
            const unsigned long long f = splitmix64(key + ((unsigned long long)round << 32) + r) & mask;
            const unsigned long long nl = r, nr = l ^ f;
            l = nl; r = nr;
        }
        v = (l << hb) | r;
    } while (v >= n);
    return v;
}

// inlier test of one entry under hypothesis matrix `mat` (almeida:226-239)
__device__ __forceinline__ bool is_inlier(const AlmeidaConst& c, const float4 e, const float mat[9], float target_sq)
{
    float w[3], d[2];
    unproject_world(c, e.x, e.y, w);
    delta_from_world(c, w, mat, e.x, e.y, d);
    const float sx = e.x + d[0], sy = e.y + d[1];
    const float vx = e.z - d[0], vy = e.w - d[1];
    const float ax = atanf((sx - 0.5f) / c.fx), ay = atanf((sy - 0.5f) / c.fy);   // point_angle
    const float cx = vx * cosf(ax), cy = vy * cosf(ay);
    return cx * cx + cy * cy <= target_sq;
}

constexpr int RS_NT = 128;

// One CTA per RANSAC iteration: 3-sample fit by one thread in the reference's sequential order,
// then all threads score the iteration's sample subset.
__global__ void __launch_bounds__(RS_NT) ransac_hypothesis_kernel(const ofps_mv* __restrict__ entries, size_t n,
                                                                  const AlmeidaConst cst, unsigned long long seed,
                                                                  size_t k_samples, float target_sq,
                                                                  float* __restrict__ fits, uint32_t* __restrict__ counts)
{
    __shared__ float s_mat[9];
    __shared__ unsigned s_cnt;
    const int tid = threadIdx.x;
    const unsigned long long it = blockIdx.x;
    const float4* e4 = reinterpret_cast<const float4*>(entries);
    if (tid == 0) {
        s_cnt = 0;
        const int s3 = n < 3 ? (int)n : 3;
        float4 smp[3];
        for (int j = 0; j < s3; j++) smp[j] = __ldg(e4 + perm_index(seed, it, 0, j, n));
        float rotation[4] = {1.0f, 0.0f, 0.0f, 0.0f};
        float a[9];
        for (int iter = 0; iter < LSQ_ITERS; iter++) {
            float rotm[9], b[3] = {0.0f, 0.0f, 0.0f};
            quat_to_mat3(rotation, rotm);
            if (iter == 0) for (int k = 0; k < 9; k++) a[k] = 0.0f;
            for (int j = 0; j < s3; j++) {
                float v[4][2];
                entry_vectors(cst, smp[j], rotm, iter == 0, v);
                if (iter == 0)
                    for (int r = 0; r < 3; r++)
                        for (int c = 0; c < 3; c++)
                            a[r * 3 + c] = a[r * 3 + c] + (v[r + 1][0] * v[c + 1][0] + v[r + 1][1] * v[c + 1][1]);
                for (int r = 0; r < 3; r++) b[r] = b[r] + (v[r + 1][0] * v[0][0] + v[r + 1][1] * v[0][1]);
            }
            lsq_step(a, b, cst.eps_r, iter, rotation);
        }
        // fit = rotation.inverse(); mat = fit.inverse().to_homogeneous() (almeida:199, 224)
        const float fit[4] = {rotation[0], -rotation[1], -rotation[2], -rotation[3]};
        const float inv[4] = {fit[0], -fit[1], -fit[2], -fit[3]};
        float m[9];
        quat_to_mat3(inv, m);
        for (int k = 0; k < 9; k++) s_mat[k] = m[k];
        for (int k = 0; k < 4; k++) fits[it * 4 + k] = fit[k];
    }
    __syncthreads();
    float mat[9];
    for (int k = 0; k < 9; k++) mat[k] = s_mat[k];
    unsigned local = 0;
    for (size_t j = tid; j < k_samples; j += RS_NT) {
        const float4 e = __ldg(e4 + perm_index(seed, it, 1, j, n));
        local += is_inlier(cst, e, mat, target_sq) ? 1u : 0u;
    }
    for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
    if ((tid & 31) == 0) atomicAdd(&s_cnt, local);
    __syncthreads();
    if (tid == 0) counts[it] = s_cnt;
}

constexpr int SEL_NT = 1024;

// Pick the iteration with the most inliers (strictly greater wins -> earliest on ties,
// almeida:241-243) and rebuild its inlier list in sample order.
__global__ void __launch_bounds__(SEL_NT) ransac_select_kernel(const ofps_mv* __restrict__ entries, size_t n,
                                                               const AlmeidaConst cst, unsigned long long seed,
                                                               size_t k_samples, float target_sq, size_t num_iters,
                                                               const float* __restrict__ fits,
                                                               const uint32_t* __restrict__ counts,
                                                               uint32_t* __restrict__ inlier_idx,
                                                               uint32_t* __restrict__ result /* [count, iter] */)
{
    __shared__ unsigned long long s_best;
    __shared__ float s_mat[9];
    __shared__ unsigned s_warp[SEL_NT / 32];
    __shared__ unsigned s_base;
    const int tid = threadIdx.x;
    if (tid == 0) { s_best = 0ull; s_base = 0; }
    __syncthreads();
    for (size_t i = tid; i < num_iters; i += SEL_NT)
        atomicMax(&s_best, ((unsigned long long)counts[i] << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)i));
    __syncthreads();
    const uint32_t best_cnt = (uint32_t)(s_best >> 32);
    const uint32_t best_it = 0xFFFFFFFFu - (uint32_t)(s_best & 0xFFFFFFFFull);
    if (tid == 0) {
        result[0] = best_cnt;
        result[1] = best_cnt ? best_it : 0u;
        const float* fit = fits + (size_t)best_it * 4;
        const float inv[4] = {fit[0], -fit[1], -fit[2], -fit[3]};
        float m[9];
        quat_to_mat3(inv, m);
        for (int k = 0; k < 9; k++) s_mat[k] = m[k];
    }
    __syncthreads();
    if (best_cnt == 0) return;
    float mat[9];
    for (int k = 0; k < 9; k++) mat[k] = s_mat[k];
    const float4* e4 = reinterpret_cast<const float4*>(entries);
    for (size_t base = 0; base < k_samples; base += SEL_NT) {
        const size_t j = base + tid;
        uint32_t e_idx = 0;
        bool in = false;
        if (j < k_samples) {
            e_idx = (uint32_t)perm_index(seed, best_it, 1, j, n);
            in = is_inlier(cst, __ldg(e4 + e_idx), mat, target_sq);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        if ((tid & 31) == 0) s_warp[tid >> 5] = __popc(bal);
        __syncthreads();
        unsigned off = s_base;
        for (int w = 0; w < (tid >> 5); w++) off += s_warp[w];
        if (in) inlier_idx[off + __popc(bal & ((1u << (tid & 31)) - 1u))] = e_idx;
        __syncthreads();
        if (tid == 0) {
            unsigned tot = 0;
            for (int w = 0; w < SEL_NT / 32; w++) tot += s_warp[w];
            s_base += tot;
        }
        __syncthreads();
    }
}

void mat3_from_euler_host(float roll, float pitch, float yaw, float m[9])
{
    // Rotation3::from_euler_angles(roll, pitch, yaw) = Rz(yaw) Ry(pitch) Rx(roll)
    const float sr = sinf(roll), cr = cosf(roll);
    const float sp = sinf(pitch), cp = cosf(pitch);
    const float sy = sinf(yaw), cy = cosf(yaw);
    m[0] = cy * cp; m[1] = cy * sp * sr - sy * cr; m[2] = cy * sp * cr + sy * sr;
    m[3] = sy * cp; m[4] = sy * sp * sr + cy * cr; m[5] = sy * sp * cr - cy * sr;
    m[6] = -sp;     m[7] = cp * sr;                m[8] = cp * cr;
}

void make_const(float aspect, float fov_y_deg, AlmeidaConst& c)
{
    // StandardCamera::new (camera.rs:26-35): Perspective3::new(aspect, fovy, 0.1, 10) and its inverse
    const float DEG2RAD = 0.017453292519943295f;
    const float znear = 0.1f, zfar = 10.0f;
    const float fovy = fov_y_deg * DEG2RAD;
    const float m11 = 1.0f / tanf(fovy / 2.0f);
    const float m00 = m11 / aspect;
    const float m22 = (zfar + znear) / (znear - zfar);
    const float m23 = zfar * znear * 2.0f / (znear - zfar);
    c.proj[0] = m00; c.proj[1] = m11; c.proj[2] = m22; c.proj[3] = m23;
    float ip[16] = {0};
    const float m32 = -1.0f;
    ip[0] = 1.0f / m00;
    ip[5] = 1.0f / m11;
    ip[10] = 0.0f;
    ip[11] = 1.0f / m32;
    ip[14] = 1.0f / m23;
    ip[15] = -m22 / (m23 * m32);
    static const float view_t[16] = {-1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1};
    for (int j = 0; j < 4; j++)       // nalgebra gemm order: first column scaled, the rest axpy'd
        for (int i = 0; i < 4; i++) {
            float acc = view_t[i * 4 + 0] * ip[0 * 4 + j];
            for (int k = 1; k < 4; k++) acc = acc + view_t[i * 4 + k] * ip[k * 4 + j];
            c.unproj[i * 4 + j] = acc;
        }
    const float PI_F = 3.14159265358979323846f;
    c.eps_r = 0.001f * PI_F / 180.0f;                       // almeida:17
    mat3_from_euler_host(0.0f, c.eps_r, 0.0f, c.roll);      // almeida:33
    mat3_from_euler_host(c.eps_r, 0.0f, 0.0f, c.pitch);     // almeida:37
    mat3_from_euler_host(0.0f, 0.0f, -c.eps_r, c.yaw);      // almeida:41
    c.fy = 0.5f / tanf((fov_y_deg * DEG2RAD) / 2.0f);       // camera.rs:120-129
    c.fx = c.fy / aspect;
}

// Entries handled by the one-CTA solver; above, a persistent multi-CTA grid with ONE entry per thread where the device
// has room: every iteration is ~370 dependent warp instructions per entry, so the iteration time is set by the entries a
// thread walks, not by the grid barrier (ncu r2: 12,600 entries at 4 per thread = 10 us per iteration; 2,000 entries in
// one CTA = 7.5 us per iteration).
constexpr size_t SINGLE_MAX = 512;

int run_lsq(const ofps_mv* d_entries, const uint32_t* d_idx, size_t n, const uint32_t* d_n_ptr, const AlmeidaConst& cst,
            float* d_quat, AlmeidaScratch& s, int sm_count, cudaStream_t stream, uint64_t* launches)
{
    if (int rc = s.state.reserve(sizeof(AlmeidaState))) return rc;
    AlmeidaState* st = s.state.as<AlmeidaState>();
    if (n <= SINGLE_MAX) {
        almeida_lsq_kernel<LSQ_SINGLE><<<1, LSQ_NT, 0, stream>>>(d_entries, d_idx, n, d_n_ptr, cst, st, nullptr, d_quat, 0);
        OFPSB_CUDA_TRY(cudaGetLastError());
        if (launches) ++*launches;
        return OFPSB_OK;
    }
    int dev = 0;
    OFPSB_CUDA_TRY(cudaGetDevice(&dev));
    // up to 16 x 1024 entries: one thread-block cluster, one entry per thread (sizes 1, 2, 4, 8 are portable, 16 needs the
    // opt-in and a GPC with 16 free SMs: checked once per device)
    if (n <= (size_t)CL_MAX_CTAS * CL_NT && !s.no_cooperative && !s.no_cluster) {
        static int cluster_ok[64] = {};   // 0 unknown, 1 yes, -1 no
        // as many CTAs as the cluster may have once there are 64 entries for each: an iteration is bound by what ONE SM has to
        // do (1024 threads: ~1,000 cycles of shuffles for the f64 warp sums alone), not by the exchange
        unsigned ncta = 1;
        while (ncta < (unsigned)CL_MAX_CTAS && (size_t)ncta * 64 < n) ncta *= 2;
        const unsigned nt = (unsigned)(((n + ncta - 1) / ncta + 31) / 32 * 32);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(ncta);
        cfg.blockDim = dim3(nt < 32 ? 32 : nt);
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = ncta;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int& ok = cluster_ok[dev & 63];
        if (ok == 0) {
            ok = -1;
            if (cudaFuncSetAttribute(almeida_lsq_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
                cudaLaunchConfig_t probe = cfg;
                probe.gridDim = dim3(CL_MAX_CTAS);
                probe.blockDim = dim3(CL_NT);
                cudaLaunchAttribute pa[1] = {attr[0]};
                pa[0].val.clusterDim.x = CL_MAX_CTAS;
                probe.attrs = pa;
                int nclusters = 0;
                if (cudaOccupancyMaxActiveClusters(&nclusters, almeida_lsq_cluster_kernel, &probe) == cudaSuccess && nclusters >= 1) ok = 1;
            }
            cudaGetLastError();
        }
        if (ok == 1 &&
            cudaLaunchKernelEx(&cfg, almeida_lsq_cluster_kernel, d_entries, d_idx, n, d_n_ptr, cst, d_quat) == cudaSuccess) {
            if (launches) ++*launches;
            return OFPSB_OK;
        }
        cudaGetLastError();
    }
    // a co-resident grid: the device can hold `per_sm` CTAs per SM at once
    static int per_sm_cached[64] = {};
    int per_sm = dev >= 0 && dev < 64 ? per_sm_cached[dev] : 0;
    if (per_sm == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, almeida_lsq_kernel<LSQ_PERSISTENT>, LSQ_NT, 0) != cudaSuccess ||
            per_sm < 1) {
            cudaGetLastError();
            per_sm = 1;
        }
        if (dev >= 0 && dev < 64) per_sm_cached[dev] = per_sm;
    }
    const size_t want = (n + (size_t)LSQ_NT - 1) / (size_t)LSQ_NT;
    const size_t cap = (size_t)(sm_count > 0 ? sm_count : 148) * (size_t)(per_sm < 4 ? per_sm : 4);
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    if (int rc = s.partial.reserve((size_t)grid * 2 * 12 * sizeof(double))) return rc;
    OFPSB_CUDA_TRY(cudaMemsetAsync(st, 0, sizeof(AlmeidaState), stream));
    double* partial = s.partial.as<double>();
    int it0 = 0;
    void* args[] = {(void*)&d_entries, (void*)&d_idx, (void*)&n, (void*)&d_n_ptr, (void*)&cst, (void*)&st, (void*)&partial,
                    (void*)&d_quat, (void*)&it0};
    if (!s.no_cooperative &&
        cudaLaunchCooperativeKernel((const void*)almeida_lsq_kernel<LSQ_PERSISTENT>, dim3(grid), dim3(LSQ_NT), args, 0, stream) ==
            cudaSuccess) {
        if (launches) ++*launches;
        return OFPSB_OK;
    }
    cudaGetLastError();
    for (int it = 0; it < LSQ_ITERS; it++)
        almeida_lsq_kernel<LSQ_STEPWISE><<<grid, LSQ_NT, 0, stream>>>(d_entries, d_idx, n, d_n_ptr, cst, st, partial, d_quat, it);
    OFPSB_CUDA_TRY(cudaGetLastError());
    if (launches) *launches += LSQ_ITERS;
    return OFPSB_OK;
}

}  // namespace

int launch_almeida(const ofps_mv* d_entries, size_t n, float aspect, float fov_y_deg, int use_ransac, size_t num_iters,
                   float inlier_angle_deg, size_t ransac_samples, uint64_t seed, float* d_quat, AlmeidaScratch& s,
                   int sm_count, cudaStream_t stream, uint64_t* launches)
{
    if (n > 0xFFFFFFF0ull) {
        set_error("almeida: too many entries (%zu)", n);
        return OFPSB_E_INVALID;
    }
    AlmeidaConst cst;
    make_const(aspect, fov_y_deg, cst);
    if (!use_ransac) return run_lsq(d_entries, nullptr, n, nullptr, cst, d_quat, s, sm_count, stream, launches);

    if (num_iters == 0 || num_iters > 1000000) {
        set_error("almeida: ransac iterations %zu out of range", num_iters);
        return OFPSB_E_INVALID;
    }
    const size_t k = ransac_samples < n ? ransac_samples : n;
    const float target = inlier_angle_deg * 0.017453292519943295f;   // almeida:210
    const float target_sq = target * target;
    if (int rc = s.hyp.reserve(num_iters * 4 * sizeof(float) + num_iters * sizeof(uint32_t))) return rc;
    if (int rc = s.inlier_idx.reserve((k ? k : 1) * sizeof(uint32_t))) return rc;
    if (int rc = s.flags.reserve(2 * sizeof(uint32_t))) return rc;
    float* fits = s.hyp.as<float>();
    uint32_t* counts = reinterpret_cast<uint32_t*>(fits + num_iters * 4);
    uint32_t* result = s.flags.as<uint32_t>();
    ransac_hypothesis_kernel<<<(unsigned)num_iters, RS_NT, 0, stream>>>(d_entries, n, cst, seed, k, target_sq, fits, counts);
    ransac_select_kernel<<<1, SEL_NT, 0, stream>>>(d_entries, n, cst, seed, k, target_sq, num_iters, fits, counts,
                                                   s.inlier_idx.as<uint32_t>(), result);
    OFPSB_CUDA_TRY(cudaGetLastError());
    if (launches) *launches += 2;
    // refit on the best inlier set; the count lives on the device (result[0])
    return run_lsq(d_entries, s.inlier_idx.as<uint32_t>(), k, result, cst, d_quat, s, sm_count, stream, launches);
}

}  // namespace ofpsb
