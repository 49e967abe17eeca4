// Internal declarations shared by the kernels and the C ABI (api.cu).
#pragma once

#ifdef OFPSB_EMU
#include "cuda_emu.h"   // tests/emu: CPU stand-in used by the kernel-logic tests only (never in the product build)
#else
#include <cuda_runtime.h>
// Kernel launch on `stream` without dynamic shared memory; the emulation build redefines it.
#define OFPSB_LAUNCH(kernel, grid, block, stream, ...) kernel<<<grid, block, 0, stream>>>(__VA_ARGS__)
#define OFPSB_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...) kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
// the kernel's dynamic shared memory as a byte array
#define OFPSB_DYN_SMEM(name) extern __shared__ __align__(16) uint8_t name[]
// keeps the compiler from sinking already-issued loads under a later predicate
#define OFPSB_KEEP_LOADED(a, b) asm volatile("" : "+f"(a), "+f"(b))
#endif
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstddef>
#include <vector>

#include "../../include/ofps_b200.h"

namespace ofpsb {

void set_error(const char* fmt, ...);

#define OFPSB_CUDA_TRY(expr)                                                                      \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            ::ofpsb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                               __LINE__);                                                         \
            return OFPSB_E_CUDA;                                                                  \
        }                                                                                         \
    } while (0)

// Growable device scratch buffer owned by a context.
struct DevBuf {
    void* ptr = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);
    void release();
    template <typename T> T* as() const { return reinterpret_cast<T*>(ptr); }
};

// Growable pinned host buffer.
struct PinBuf {
    void* ptr = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);
    void release();
    template <typename T> T* as() const { return reinterpret_cast<T*>(ptr); }
};

// ---------------------------------------------------------------------------- block matcher
struct BlockMatchParams {
    const uint8_t* prev;   // row 0 = the row of the previous frame aligned with cur row 0
    const uint8_t* cur;    // row 0 of the strip / frame
    int w;                 // frame width in pixels
    int strip_h;           // rows of cur handled by this launch
    int stride;            // bytes per row (both planes)
    long long pair_stride; // bytes between consecutive pairs (batched launches)
    int n_pairs;
    int halo_top, halo_bottom;  // valid prev rows before row 0 / after row strip_h-1
    int y_offset, full_h;       // position of the strip in the whole frame
    int block, range, metric;
    int nbx, nby;               // full blocks in this launch
    int16_t* mv_xy;             // [n_pairs][nby*nbx][2] or null
    uint32_t* cost;             // [n_pairs][nby*nbx] or null
    ofps_mv* entries;           // [n_pairs][nby*nbx] or null
};

// Launches the block matcher (tuned template instance if one exists for (block, range),
// otherwise the generic kernel).  Returns 0 or OFPSB_E_*; *launches += kernels launched.
int launch_block_match(const BlockMatchParams& p, cudaStream_t stream, uint64_t* launches, int variant = 0);
// Scratch of the pruned search (block_match_prune.cu): window-sum planes and the work list.
struct BlockMatchScratch {
    DevBuf sums, worklist;
    bool collect_stats = false;
    bool profile = false;     // record events around the SEA / work-list kernels (ofpsb_block_match_kernel_ms)
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    // content feedback: the work-list count of every SEA launch is read back asynchronously; when most blocks of the last
    // launch ended on the work list (heavy noise: the bounds prune nothing) the next launches go straight to the exhaustive
    // kernel, with a SEA launch every 16th call to notice when the content changes.  Results are identical either way.
    uint32_t* h_listed = nullptr;       // pinned: [0] listed blocks of the last SEA launch
    cudaEvent_t ev_listed = nullptr;
    long long listed_total = 0;         // blocks of that launch
    int skip_calls = 0;                 // SEA launches still to skip
    int over_streak = 0;                // consecutive SEA launches that left most blocks undecided
    int adaptive = 1;
    int tile_h = 0;           // SEA kernel tile height: 0 = by launch size, 32 / 64 = forced (tests, A-B)
    int prefetch_tiles = -1;  // SEA kernel: L2 prefetch distance in tiles (-1 = three CTAs per SM, 0 = off)
    int pruner = 0;           // 0 = fused SEA kernel where it applies (default), 1 = round-1 window-sum pipeline (tests / A-B)
    int chunk_pairs = 0;      // pairs per pruned chunk (0 = whole batch; smaller chunks stay L2-resident but measured slower)
    size_t l2_bytes = 0;
};
// Peer-halo mode of the SEA kernel (spatial tiling): p.prev = first OWN row, own_rows rows in this rank's memory; the
// halo rows above / below come from the neighbours' tensors (null = frame border).
struct SeaPeer {
    int own_rows;
    const uint8_t* up;      // first own row of the upper neighbour's strip (peer-mapped), up_rows rows
    int up_rows, up_stride;
    long long up_pair_stride;
    const uint8_t* down;
    int down_rows, down_stride;
    long long down_pair_stride;
};
// Fused SEA kernel + exhaustive work list (block_match_sea.cu).  Returns 0 when launched, 1 when it does not apply.
int launch_block_match_sea(const BlockMatchParams& p, BlockMatchScratch& scratch, int sm_count, cudaStream_t stream,
                           uint64_t* launches, const SeaPeer* peer = nullptr, cudaEvent_t before_list = nullptr);
// Exact pruned SAD search (window-sum bounds + exhaustive search of the undecided blocks only).
// Returns 0 when launched, 1 when the path does not apply, < 0 on error.
int launch_block_match_pruned(const BlockMatchParams& p, BlockMatchScratch& scratch, int sm_count, cudaStream_t stream,
                              uint64_t* launches);
// TMA-staged instances (block_match_tma.cu).  Returns 0 when launched, 1 when no instance applies
// (geometry, or base / strides not 16-byte aligned), < 0 on error.
int launch_block_match_tma(const BlockMatchParams& p, cudaStream_t stream, int variant);
// LDG-staged tile instances only (first-generation kernel), then generic.
int launch_block_match_ldg(const BlockMatchParams& p, cudaStream_t stream, uint64_t* launches);
// Forces the generic (untuned) kernel; used by tests to cross-check the tuned instances.
int launch_block_match_generic(const BlockMatchParams& p, cudaStream_t stream, uint64_t* launches);

// ---------------------------------------------------------------------------- densify / detect
struct DensifyScratch {
    DevBuf keys_a, keys_b, vals_a, vals_b, hist, cell_start, sums;
};

// entries (device) -> field (device, gw*gh*2) [+ counts]; bit-exact reference order.
// force_path: 0 = choose by size, 1 = scan path, 2 = sort path (tests cross-check both).
// raw != 0: d_field receives the un-divided sums (the densifier state) instead of the means.
int launch_densify(const ofps_mv* d_entries, size_t n, size_t gw, size_t gh, float* d_field, float* d_counts,
                   DensifyScratch& scratch, cudaStream_t stream, uint64_t* launches, int force_path = 0, int raw = 0);

struct DetectResult {   // written by the detector kernel (device), copied to host
    unsigned long long best_key;
    unsigned int area;
    unsigned int seed_cell;
    int has_motion;
    int pad;
};

// mean field (device, dim*dim*2) -> island field (device) + result record.
int launch_detect(const float* d_mean_field, size_t dim, float target_motion, float min_size,
                  float* d_out_field, DetectResult* d_result, DevBuf& scratch, cudaStream_t stream,
                  uint64_t* launches, int force_union_find = 0);

// MotionFieldDensifier::interpolate_empty_cells on the host (hole_fill.cu): sequential by definition.
// sums / counts: 2*w*h floats each (the densifier state), updated in place.
void interpolate_empty_cells_host(float* sums, float* counts, size_t w, size_t h);

// ---------------------------------------------------------------------------- almeida
struct AlmeidaScratch {
    DevBuf state, partial, hyp, inlier_idx, flags;
    bool no_cooperative = false;   // tests: force the one-launch-per-iteration fallback of the multi-CTA solver
    bool no_cluster = false;       // tests: skip the one-cluster solver (option "almeida_cluster" = 0)
};

int launch_almeida(const ofps_mv* d_entries, size_t n, float aspect, float fov_y_deg, int use_ransac,
                   size_t num_iters, float inlier_angle_deg, size_t ransac_samples, uint64_t seed,
                   float* d_quat, AlmeidaScratch& scratch, int sm_count, cudaStream_t stream, uint64_t* launches);

// ---------------------------------------------------------------------------- cv-decoder front end (cv_front.cu)
struct FlowScratch {
    DevBuf bounds, cells, tiles;
};
// BGR(A)/RGB(A) u8 -> gray (OpenCV BGR2GRAY) and / or RGBA; either output may be null.
int launch_frame_convert(const uint8_t* d_src, int w, int h, int stride, int channels, int rgb_order, uint8_t* d_gray,
                         int gray_stride, uint8_t* d_rgba, cudaStream_t stream, uint64_t* launches);
// resize(INTER_LINEAR), 8-bit interleaved, reductions only (OpenCV fixed-point bilinear).
int launch_frame_resize(const uint8_t* d_src, int sw, int sh, int stride, int channels, uint8_t* d_dst, int dw, int dh,
                        int dst_stride, cudaStream_t stream, uint64_t* launches);
// gray -> contrast mask (0 / 255 bytes): Sobel(1,1,k5) > 20, dilated by the 11x11 ellipse, REFLECT_101 borders.
int launch_contrast_mask(const uint8_t* d_gray, int w, int h, int stride, uint8_t* d_mask, int mask_stride,
                         cudaStream_t stream, uint64_t* launches);
// dense flow (+ mask) -> MotionEntry list (device), count -> *d_count.  gw == gh == 0: one entry per kept pixel.
int launch_flow_entries(const float* d_flow, size_t flow_stride, const uint8_t* d_mask, size_t mask_stride, int w, int h,
                        size_t gw, size_t gh, ofps_mv* d_entries, size_t cap, unsigned long long* d_count,
                        FlowScratch& scratch, cudaStream_t stream, uint64_t* launches);

}  // namespace ofpsb

struct ofpsb_ctx {
    int device = 0;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    size_t l2_bytes = 0, mem_bytes = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // host->device staging of the batched host entry points
    cudaStream_t d2h_stream = nullptr;    // device->host return of their results
    cudaStream_t stream = nullptr;        // current (own or borrowed)
    uint64_t launches = 0;
    // options (ofpsb_set_option)
    int opt_densify_path = 0;
    int opt_block_match_kernel = 0;
    int opt_batch_chunk_pairs = 0;
    int opt_block_match_prune = 1;
    int opt_detect_union_find = 0;   // tests: force the union-find detector kernel on small grids
    ofpsb::BlockMatchScratch bm_scratch;
    // scratch
    ofpsb::DevBuf d_frames, d_mv, d_cost, d_entries, d_field, d_field2, d_counts, d_misc, d_detect_scratch;
    ofpsb::PinBuf h_misc;
    ofpsb::DensifyScratch densify;
    ofpsb::AlmeidaScratch almeida;
    ofpsb::FlowScratch flow;
    ofpsb::DevBuf d_cv_src, d_cv_small, d_cv_gray, d_cv_rgba, d_cv_mask, d_cv_flow, d_cv_count;
    std::vector<cudaEvent_t> events;      // pool for the batch pipeline (no timing)
};

#ifndef OFPSB_EMU
namespace ofpsb {
int launch_block_match_ctx(ofpsb_ctx* ctx, const BlockMatchParams& p, BlockMatchScratch& scratch, cudaStream_t stream);

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

#define OFPSB_ENTER(ctx)                                           \
    if (!(ctx)) {                                                  \
        ::ofpsb::set_error("null context");                        \
        return OFPSB_E_INVALID;                                    \
    }                                                              \
    ::ofpsb::DeviceGuard _guard((ctx)->device);                    \
    if (!_guard.ok) {                                              \
        ::ofpsb::set_error("cudaSetDevice(%d) failed", (ctx)->device); \
        return OFPSB_E_CUDA;                                       \
    }
}  // namespace ofpsb
#endif
