// C ABI of libofps_b200.so (include/ofps_b200.h): context, host/device entry points, the
// pipelined batch path and the .mvec / .flo interchange files.  No CPU fallback anywhere: every
// compute entry point needs a context, and a context needs an sm_100 device.
#include "common.cuh"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>

namespace ofpsb {

static thread_local char g_err[768] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int DevBuf::reserve(size_t bytes)
{
    if (bytes <= cap && ptr) return OFPSB_OK;
    release();
    size_t want = (bytes + 255) & ~(size_t)255;
    if (want == 0) want = 256;
    cudaError_t e = cudaMalloc(&ptr, want);
    if (e != cudaSuccess) {
        ptr = nullptr;
        set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        return OFPSB_E_NOMEM;
    }
    cap = want;
    return OFPSB_OK;
}

void DevBuf::release()
{
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
}

int PinBuf::reserve(size_t bytes)
{
    if (bytes <= cap && ptr) return OFPSB_OK;
    release();
    size_t want = (bytes + 4095) & ~(size_t)4095;
    if (want == 0) want = 4096;
    cudaError_t e = cudaHostAlloc(&ptr, want, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        ptr = nullptr;
        set_error("cudaHostAlloc(%zu) failed: %s", want, cudaGetErrorString(e));
        return OFPSB_E_NOMEM;
    }
    cap = want;
    return OFPSB_OK;
}

void PinBuf::release()
{
    if (ptr) cudaFreeHost(ptr);
    ptr = nullptr;
    cap = 0;
}

// The block-matching dispatch of a context on a given scratch / stream (the streaming decoder alternates two of each so
// that consecutive pairs overlap on the device).
int launch_block_match_ctx(ofpsb_ctx* ctx, const BlockMatchParams& p, BlockMatchScratch& scratch, cudaStream_t stream)
{
    // grid.z carries the pair index: split very large batches
    const size_t nb = (size_t)p.nbx * p.nby;
    for (int first = 0; first < p.n_pairs; first += 32768) {
        BlockMatchParams q = p;
        q.n_pairs = p.n_pairs - first < 32768 ? p.n_pairs - first : 32768;
        q.prev = p.prev + (long long)first * p.pair_stride;
        q.cur = p.cur + (long long)first * p.pair_stride;
        if (p.mv_xy) q.mv_xy = p.mv_xy + 2 * nb * first;
        if (p.cost) q.cost = p.cost + nb * first;
        if (p.entries) q.entries = p.entries + nb * first;
        int rc = 1;
        if (ctx->opt_block_match_prune && ctx->opt_block_match_kernel == 0)
            rc = launch_block_match_pruned(q, scratch, ctx->sm_count, stream, &ctx->launches);
        if (rc != 1) {
            if (rc) return rc;
            continue;
        }
        switch (ctx->opt_block_match_kernel) {
            case 1: rc = launch_block_match_generic(q, stream, &ctx->launches); break;
            case 2: rc = launch_block_match_ldg(q, stream, &ctx->launches); break;
            case 3: rc = launch_block_match(q, stream, &ctx->launches, 1); break;
            default: rc = launch_block_match(q, stream, &ctx->launches, 0); break;
        }
        if (rc) return rc;
    }
    return OFPSB_OK;
}

namespace {

int get_event(ofpsb_ctx* ctx, size_t idx, cudaEvent_t* out)
{
    while (ctx->events.size() <= idx) {
        cudaEvent_t e;
        OFPSB_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->events.push_back(e);
    }
    *out = ctx->events[idx];
    return OFPSB_OK;
}

int fill_params(BlockMatchParams& p, const uint8_t* d_prev, const uint8_t* d_cur, int w, int h, int stride,
                size_t pair_stride, int n_pairs, int block, int range, int metric, int16_t* mv, uint32_t* cost,
                ofps_mv* entries)
{
    if (block <= 0) {
        set_error("block_match: block must be positive");
        return OFPSB_E_INVALID;
    }
    p.prev = d_prev;
    p.cur = d_cur;
    p.w = w;
    p.strip_h = h;
    p.stride = stride;
    p.pair_stride = (long long)pair_stride;
    p.n_pairs = n_pairs;
    p.halo_top = 0;
    p.halo_bottom = 0;
    p.y_offset = 0;
    p.full_h = h;
    p.block = block;
    p.range = range;
    p.metric = metric;
    p.nbx = w / block;
    p.nby = h / block;
    p.mv_xy = mv;
    p.cost = cost;
    p.entries = entries;
    return OFPSB_OK;
}

int launch_bm(ofpsb_ctx* ctx, const BlockMatchParams& p) { return launch_block_match_ctx(ctx, p, ctx->bm_scratch, ctx->stream); }

size_t block_dim_host(float min_size, size_t subdivide)
{
    // block-motion-detector/src/lib.rs:53-54
    const float block_width = sqrtf(min_size) / (float)subdivide;
    const float v = ceilf(1.0f / block_width);
    if (!(v > 0.0f)) return 0;                       // NaN / negative -> `as usize` = 0
    if (v >= 18446744073709551616.0f) return SIZE_MAX;
    return (size_t)v;
}

// entries on the device -> densify -> detect -> host results
int detect_from_device_entries(ofpsb_ctx* ctx, const ofps_mv* d_entries, size_t n, float min_size, size_t subdivide,
                               float target_motion, int* has_motion, size_t* area, size_t* dim_out, float* field_xy,
                               size_t field_cap_cells)
{
    const size_t dim = block_dim_host(min_size, subdivide);
    if (dim_out) *dim_out = dim;
    if (has_motion) *has_motion = 0;
    if (area) *area = 0;
    if (dim == 0 || dim > 16384) {
        set_error("detect_block_motion: block_dim %zu out of range (min_size=%g subdivide=%zu)", dim, (double)min_size,
                  subdivide);
        return OFPSB_E_INVALID;
    }
    const size_t cells = dim * dim;
    if (field_xy && field_cap_cells < cells) {
        set_error("detect_block_motion: field buffer holds %zu cells, %zu needed", field_cap_cells, cells);
        return OFPSB_E_CAPACITY;
    }
    // island field and result record live back to back so that one D2H copy (into pinned memory) returns both
    const size_t rec = (sizeof(DetectResult) + 15) & ~(size_t)15;
    if (int rc = ctx->d_field.reserve(cells * 8)) return rc;
    if (int rc = ctx->d_field2.reserve(rec + cells * 8)) return rc;
    if (int rc = ctx->h_misc.reserve(rec + cells * 8)) return rc;
    DetectResult* d_res = ctx->d_field2.as<DetectResult>();
    float* d_island = reinterpret_cast<float*>(ctx->d_field2.as<char>() + rec);
    if (int rc = launch_densify(d_entries, n, dim, dim, ctx->d_field.as<float>(), nullptr, ctx->densify, ctx->stream,
                                &ctx->launches, ctx->opt_densify_path))
        return rc;
    if (int rc = launch_detect(ctx->d_field.as<float>(), dim, target_motion, min_size, d_island, d_res,
                               ctx->d_detect_scratch, ctx->stream, &ctx->launches, ctx->opt_detect_union_find))
        return rc;
    OFPSB_CUDA_TRY(cudaMemcpyAsync(ctx->h_misc.ptr, d_res, field_xy ? rec + cells * 8 : rec, cudaMemcpyDeviceToHost,
                                   ctx->stream));
    OFPSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (field_xy) memcpy(field_xy, ctx->h_misc.as<char>() + rec, cells * 8);
    const DetectResult* r = ctx->h_misc.as<DetectResult>();
    if (has_motion) *has_motion = r->has_motion;
    if (area) *area = r->area;
    return OFPSB_OK;
}

}  // namespace
}  // namespace ofpsb

using namespace ofpsb;

#pragma GCC visibility push(default)
extern "C" {

// ------------------------------------------------------------------------------------ context
int ofpsb_create(int device, ofpsb_ctx** out)
{
    if (!out) {
        set_error("ofpsb_create: null output pointer");
        return OFPSB_E_INVALID;
    }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("no CUDA device available (%s); libofps_b200 has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return OFPSB_E_NODEVICE;
    }
    if (device < 0 || device >= count) {
        set_error("device %d out of range (0..%d)", device, count - 1);
        return OFPSB_E_INVALID;
    }
    cudaDeviceProp prop;
    OFPSB_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d (%s, cc %d.%d) is not an sm_100-class part; libofps_b200 carries sm_100a code only",
                  device, prop.name, prop.major, prop.minor);
        return OFPSB_E_NODEVICE;
    }
    DeviceGuard guard(device);
    if (!guard.ok) {
        set_error("cudaSetDevice(%d) failed", device);
        return OFPSB_E_CUDA;
    }
    ofpsb_ctx* ctx = new (std::nothrow) ofpsb_ctx();
    if (!ctx) {
        set_error("out of host memory");
        return OFPSB_E_NOMEM;
    }
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->cc_major = prop.major;
    ctx->cc_minor = prop.minor;
    ctx->l2_bytes = (size_t)prop.l2CacheSize;
    ctx->bm_scratch.l2_bytes = ctx->l2_bytes;
    ctx->mem_bytes = prop.totalGlobalMem;
    cudaError_t e1 = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    cudaError_t e2 = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    cudaError_t e3 = cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
        set_error("cudaStreamCreate failed");
        ofpsb_destroy(ctx);
        return OFPSB_E_CUDA;
    }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return OFPSB_OK;
}

void ofpsb_destroy(ofpsb_ctx* ctx)
{
    if (!ctx) return;
    DeviceGuard guard(ctx->device);
    cudaDeviceSynchronize();
    DevBuf* bufs[] = {&ctx->d_frames, &ctx->d_mv, &ctx->d_cost, &ctx->d_entries, &ctx->d_field, &ctx->d_field2,
                      &ctx->d_counts, &ctx->d_misc, &ctx->d_detect_scratch, &ctx->densify.keys_a, &ctx->densify.keys_b,
                      &ctx->densify.vals_a, &ctx->densify.vals_b, &ctx->densify.hist, &ctx->densify.cell_start,
                      &ctx->densify.sums, &ctx->bm_scratch.sums, &ctx->bm_scratch.worklist, &ctx->almeida.state, &ctx->almeida.partial, &ctx->almeida.hyp,
                      &ctx->almeida.inlier_idx, &ctx->almeida.flags, &ctx->flow.bounds, &ctx->flow.cells, &ctx->flow.tiles,
                      &ctx->d_cv_src, &ctx->d_cv_small, &ctx->d_cv_gray, &ctx->d_cv_rgba, &ctx->d_cv_mask, &ctx->d_cv_flow, &ctx->d_cv_count};
    for (DevBuf* b : bufs) b->release();
    ctx->h_misc.release();
    for (cudaEvent_t ev : ctx->events) cudaEventDestroy(ev);
    for (cudaEvent_t ev : ctx->bm_scratch.ev)
        if (ev) cudaEventDestroy(ev);
    if (ctx->bm_scratch.ev_listed) cudaEventDestroy(ctx->bm_scratch.ev_listed);
    if (ctx->bm_scratch.h_listed) cudaFreeHost(ctx->bm_scratch.h_listed);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
    delete ctx;
}

const char* ofpsb_last_error(void) { return g_err; }

const char* ofpsb_version(void) { return "ofps_b200 0.1.0 (sm_100a)"; }

int ofpsb_set_stream(ofpsb_ctx* ctx, void* cuda_stream)
{
    OFPSB_ENTER(ctx);
    ctx->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return OFPSB_OK;
}

void* ofpsb_get_stream(ofpsb_ctx* ctx) { return ctx ? reinterpret_cast<void*>(ctx->stream) : nullptr; }

int ofpsb_sync(ofpsb_ctx* ctx)
{
    OFPSB_ENTER(ctx);
    OFPSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return OFPSB_OK;
}

int ofpsb_device_info(ofpsb_ctx* ctx, int* sm_count, size_t* l2_bytes, size_t* mem_bytes, int* cc_major, int* cc_minor)
{
    if (!ctx) {
        set_error("null context");
        return OFPSB_E_INVALID;
    }
    if (sm_count) *sm_count = ctx->sm_count;
    if (l2_bytes) *l2_bytes = ctx->l2_bytes;
    if (mem_bytes) *mem_bytes = ctx->mem_bytes;
    if (cc_major) *cc_major = ctx->cc_major;
    if (cc_minor) *cc_minor = ctx->cc_minor;
    return OFPSB_OK;
}

uint64_t ofpsb_launch_count(ofpsb_ctx* ctx) { return ctx ? ctx->launches : 0; }

int ofpsb_block_match_stats(ofpsb_ctx* ctx, uint64_t out[4])
{
    OFPSB_ENTER(ctx);
    if (!out) {
        set_error("block_match_stats: null output");
        return OFPSB_E_INVALID;
    }
    out[0] = out[1] = out[2] = out[3] = 0;
    if (!ctx->bm_scratch.worklist.ptr) return OFPSB_OK;
    OFPSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    unsigned char raw[64];
    OFPSB_CUDA_TRY(cudaMemcpy(raw, ctx->bm_scratch.worklist.ptr, 64, cudaMemcpyDeviceToHost));
    uint32_t cnt;
    memcpy(&cnt, raw, 4);
    memcpy(out, raw + 8, 24);
    (void)cnt;
    out[3] = out[0] - out[1];
    return OFPSB_OK;
}

int ofpsb_block_match_kernel_ms(ofpsb_ctx* ctx, float out[2])
{
    OFPSB_ENTER(ctx);
    if (!out) {
        set_error("block_match_kernel_ms: null output");
        return OFPSB_E_INVALID;
    }
    out[0] = out[1] = 0.0f;
    BlockMatchScratch& sc = ctx->bm_scratch;
    if (!sc.profile || !sc.ev[0]) {
        set_error("block_match_kernel_ms: option block_match_profile is off");
        return OFPSB_E_INVALID;
    }
    OFPSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (cudaEventElapsedTime(&out[0], sc.ev[0], sc.ev[1]) != cudaSuccess ||
        cudaEventElapsedTime(&out[1], sc.ev[1], sc.ev[2]) != cudaSuccess) {
        cudaGetLastError();
        set_error("block_match_kernel_ms: no default-path launch recorded yet");
        return OFPSB_E_INVALID;
    }
    return OFPSB_OK;
}

int ofpsb_set_option(ofpsb_ctx* ctx, const char* key, long long value)
{
    if (!ctx || !key) {
        set_error("set_option: null argument");
        return OFPSB_E_INVALID;
    }
    if (!strcmp(key, "densify_path") && value >= 0 && value <= 2) ctx->opt_densify_path = (int)value;
    else if (!strcmp(key, "block_match_kernel") && value >= 0 && value <= 3) ctx->opt_block_match_kernel = (int)value;
    else if (!strcmp(key, "batch_chunk_pairs") && value >= 0 && value <= 65535) ctx->opt_batch_chunk_pairs = (int)value;
    else if (!strcmp(key, "block_match_prune") && value >= 0 && value <= 1) ctx->opt_block_match_prune = (int)value;
    else if (!strcmp(key, "block_match_stats") && value >= 0 && value <= 1) ctx->bm_scratch.collect_stats = value != 0;
    else if (!strcmp(key, "block_match_adaptive") && value >= 0 && value <= 1) {
        ctx->bm_scratch.adaptive = (int)value;
        ctx->bm_scratch.skip_calls = 0;
        ctx->bm_scratch.over_streak = 0;
        ctx->bm_scratch.listed_total = 0;
    }
    else if (!strcmp(key, "block_match_tile_h") && (value == 0 || value == 32 || value == 64)) ctx->bm_scratch.tile_h = (int)value;
    else if (!strcmp(key, "block_match_prefetch_tiles") && value >= -1 && value <= 1000000) ctx->bm_scratch.prefetch_tiles = (int)value;
    else if (!strcmp(key, "block_match_profile") && value >= 0 && value <= 1) {
        OFPSB_ENTER(ctx);
        for (int i = 0; i < 3 && value; i++)
            if (!ctx->bm_scratch.ev[i]) OFPSB_CUDA_TRY(cudaEventCreate(&ctx->bm_scratch.ev[i]));
        ctx->bm_scratch.profile = value != 0;
    }
    else if (!strcmp(key, "detect_union_find") && value >= 0 && value <= 1) ctx->opt_detect_union_find = (int)value;
    else if (!strcmp(key, "almeida_stepwise") && value >= 0 && value <= 1) ctx->almeida.no_cooperative = value != 0;
    else if (!strcmp(key, "almeida_cluster") && value >= 0 && value <= 1) ctx->almeida.no_cluster = value == 0;
    else if (!strcmp(key, "block_match_pruner") && value >= 0 && value <= 1) ctx->bm_scratch.pruner = (int)value;
    else if (!strcmp(key, "block_match_chunk_pairs") && value >= 0 && value <= 32768) ctx->bm_scratch.chunk_pairs = (int)value;
    else {
        set_error("set_option: unknown key or value out of range: %s = %lld", key, value);
        return OFPSB_E_INVALID;
    }
    return OFPSB_OK;
}

int ofpsb_host_alloc(void** out, size_t bytes)
{
    if (!out) {
        set_error("host_alloc: null output pointer");
        return OFPSB_E_INVALID;
    }
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        *out = nullptr;
        set_error("cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        return OFPSB_E_NOMEM;
    }
    return OFPSB_OK;
}

void ofpsb_host_free(void* p)
{
    if (p) cudaFreeHost(p);
}

int ofpsb_dev_alloc(ofpsb_ctx* ctx, void** out, size_t bytes)
{
    OFPSB_ENTER(ctx);
    if (!out) {
        set_error("dev_alloc: null output pointer");
        return OFPSB_E_INVALID;
    }
    cudaError_t e = cudaMalloc(out, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        *out = nullptr;
        set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        return OFPSB_E_NOMEM;
    }
    return OFPSB_OK;
}

void ofpsb_dev_free(ofpsb_ctx* ctx, void* p)
{
    if (!ctx || !p) return;
    DeviceGuard guard(ctx->device);
    cudaFree(p);
}

int ofpsb_copy_to_device(ofpsb_ctx* ctx, void* d_dst, const void* h_src, size_t bytes)
{
    OFPSB_ENTER(ctx);
    OFPSB_CUDA_TRY(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return OFPSB_OK;
}

int ofpsb_copy_to_host(ofpsb_ctx* ctx, void* h_dst, const void* d_src, size_t bytes)
{
    OFPSB_ENTER(ctx);
    OFPSB_CUDA_TRY(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return OFPSB_OK;
}

// ------------------------------------------------------------------------------ block matcher
int ofpsb_block_match_dev(ofpsb_ctx* ctx, const uint8_t* d_prev, const uint8_t* d_cur, int w, int h, int stride,
                          size_t pair_stride, int n_pairs, int block, int range, int metric, int16_t* d_mv_xy,
                          uint32_t* d_cost, ofps_mv* d_entries)
{
    OFPSB_ENTER(ctx);
    BlockMatchParams p;
    if (int rc = fill_params(p, d_prev, d_cur, w, h, stride, pair_stride, n_pairs, block, range, metric, d_mv_xy, d_cost,
                             d_entries))
        return rc;
    return launch_bm(ctx, p);
}

int ofpsb_block_match_strip_batch_dev(ofpsb_ctx* ctx, const uint8_t* d_prev, const uint8_t* d_cur, int w, int strip_h,
                                      int stride, size_t pair_stride, int n_pairs, int halo_top, int halo_bottom,
                                      int y_offset, int full_h, int block, int range, int metric, int16_t* d_mv_xy,
                                      uint32_t* d_cost, ofps_mv* d_entries)
{
    OFPSB_ENTER(ctx);
    BlockMatchParams p;
    if (int rc = fill_params(p, d_prev, d_cur, w, strip_h, stride, pair_stride, n_pairs, block, range, metric, d_mv_xy,
                             d_cost, d_entries))
        return rc;
    if (halo_top < 0 || halo_bottom < 0 || y_offset < 0 || full_h < y_offset + strip_h) {
        set_error("block_match_strip: bad strip placement (y_offset=%d strip_h=%d full_h=%d halos=%d/%d)", y_offset,
                  strip_h, full_h, halo_top, halo_bottom);
        return OFPSB_E_INVALID;
    }
    p.halo_top = halo_top;
    p.halo_bottom = halo_bottom;
    p.y_offset = y_offset;
    p.full_h = full_h;
    return launch_bm(ctx, p);
}

int ofpsb_block_match_strip_dev(ofpsb_ctx* ctx, const uint8_t* d_prev, const uint8_t* d_cur, int w, int strip_h,
                                int stride, int halo_top, int halo_bottom, int y_offset, int full_h, int block,
                                int range, int metric, int16_t* d_mv_xy, uint32_t* d_cost, ofps_mv* d_entries)
{
    return ofpsb_block_match_strip_batch_dev(ctx, d_prev, d_cur, w, strip_h, stride, 0, 1, halo_top, halo_bottom, y_offset,
                                             full_h, block, range, metric, d_mv_xy, d_cost, d_entries);
}

int ofpsb_block_match_batch(ofpsb_ctx* ctx, const uint8_t* prev, const uint8_t* cur, int w, int h, int stride,
                            size_t pair_stride, int n_pairs, int block, int range, int metric, int16_t* mv_xy,
                            uint32_t* cost, ofps_mv* entries, size_t* n_blocks)
{
    OFPSB_ENTER(ctx);
    if (!prev || !cur || w <= 0 || h <= 0 || stride < w || n_pairs <= 0 || block <= 0) {
        set_error("block_match: invalid arguments (w=%d h=%d stride=%d block=%d pairs=%d)", w, h, stride, block, n_pairs);
        return OFPSB_E_INVALID;
    }
    const size_t frame_bytes = (size_t)stride * h;
    if (n_pairs > 1 && pair_stride < frame_bytes) {
        set_error("block_match_batch: pair_stride %zu smaller than one frame (%zu bytes)", pair_stride, frame_bytes);
        return OFPSB_E_INVALID;
    }
    const size_t nb = (size_t)(w / block) * (size_t)(h / block);
    if (n_blocks) *n_blocks = nb;
    // Consecutive frames of one stream (pair i = frames i, i+1): each frame is uploaded once.
    const bool stream_mode = n_pairs > 1 && cur == prev + pair_stride;
    // Bound the device staging: split very large batches.
    const size_t max_stage = (size_t)8 << 30;
    const size_t per_pair = stream_mode ? frame_bytes : 2 * frame_bytes;
    if ((size_t)n_pairs * per_pair > max_stage && n_pairs > 1) {
        const int half = n_pairs / 2;
        int rc = ofpsb_block_match_batch(ctx, prev, cur, w, h, stride, pair_stride, half, block, range, metric, mv_xy,
                                         cost, entries, nullptr);
        if (rc) return rc;
        return ofpsb_block_match_batch(ctx, prev + (size_t)half * pair_stride, cur + (size_t)half * pair_stride, w, h,
                                       stride, pair_stride, n_pairs - half, block, range, metric,
                                       mv_xy ? mv_xy + 2 * nb * half : nullptr, cost ? cost + nb * half : nullptr,
                                       entries ? entries + nb * half : nullptr, nullptr);
    }
    const size_t n_frames_dev = stream_mode ? (size_t)n_pairs + 1 : 2 * (size_t)n_pairs;
    if (int rc = ctx->d_frames.reserve(n_frames_dev * frame_bytes)) return rc;
    if (mv_xy) if (int rc = ctx->d_mv.reserve(nb * n_pairs * 4)) return rc;
    if (cost) if (int rc = ctx->d_cost.reserve(nb * n_pairs * 4)) return rc;
    if (entries) if (int rc = ctx->d_entries.reserve(nb * n_pairs * sizeof(ofps_mv))) return rc;
    uint8_t* d_base = ctx->d_frames.as<uint8_t>();
    uint8_t* d_prev = d_base;
    uint8_t* d_cur = stream_mode ? d_base + frame_bytes : d_base + (size_t)n_pairs * frame_bytes;

    int chunk = ctx->opt_batch_chunk_pairs;
    if (chunk <= 0) {
        chunk = (int)(((size_t)24 << 20) / per_pair);   // ~24 MB of frames per chunk
        if (chunk < 1) chunk = 1;
    }
    if (chunk > n_pairs) chunk = n_pairs;
    const int n_chunks = (n_pairs + chunk - 1) / chunk;
    for (int k = 0; k < n_chunks; k++) {
        const int first = k * chunk;
        const int cnt = n_pairs - first < chunk ? n_pairs - first : chunk;
        cudaEvent_t ev_in, ev_k;
        if (int rc = get_event(ctx, 2 * (size_t)k, &ev_in)) return rc;
        if (int rc = get_event(ctx, 2 * (size_t)k + 1, &ev_k)) return rc;
        if (stream_mode) {
            const int f0 = k == 0 ? 0 : first + 1;            // frames first+1 .. first+cnt (plus frame 0 once)
            const int fn = first + cnt + 1 - f0;
            OFPSB_CUDA_TRY(cudaMemcpy2DAsync(d_base + (size_t)f0 * frame_bytes, frame_bytes, prev + (size_t)f0 * pair_stride,
                                             pair_stride, frame_bytes, (size_t)fn, cudaMemcpyHostToDevice,
                                             ctx->copy_stream));
        } else if (n_pairs == 1) {
            OFPSB_CUDA_TRY(cudaMemcpyAsync(d_prev, prev, frame_bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
            OFPSB_CUDA_TRY(cudaMemcpyAsync(d_cur, cur, frame_bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
        } else {
            OFPSB_CUDA_TRY(cudaMemcpy2DAsync(d_prev + (size_t)first * frame_bytes, frame_bytes,
                                             prev + (size_t)first * pair_stride, pair_stride, frame_bytes, (size_t)cnt,
                                             cudaMemcpyHostToDevice, ctx->copy_stream));
            OFPSB_CUDA_TRY(cudaMemcpy2DAsync(d_cur + (size_t)first * frame_bytes, frame_bytes,
                                             cur + (size_t)first * pair_stride, pair_stride, frame_bytes, (size_t)cnt,
                                             cudaMemcpyHostToDevice, ctx->copy_stream));
        }
        OFPSB_CUDA_TRY(cudaEventRecord(ev_in, ctx->copy_stream));
        OFPSB_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ev_in, 0));
        BlockMatchParams p;
        if (int rc = fill_params(p, d_prev + (size_t)first * frame_bytes, d_cur + (size_t)first * frame_bytes, w, h, stride,
                                 frame_bytes, cnt, block, range, metric,
                                 mv_xy ? ctx->d_mv.as<int16_t>() + 2 * nb * first : nullptr,
                                 cost ? ctx->d_cost.as<uint32_t>() + nb * first : nullptr,
                                 entries ? ctx->d_entries.as<ofps_mv>() + nb * first : nullptr))
            return rc;
        if (int rc = launch_bm(ctx, p)) {
            cudaStreamSynchronize(ctx->copy_stream);
            return rc;
        }
        OFPSB_CUDA_TRY(cudaEventRecord(ev_k, ctx->stream));
        OFPSB_CUDA_TRY(cudaStreamWaitEvent(ctx->d2h_stream, ev_k, 0));
        if (mv_xy)
            OFPSB_CUDA_TRY(cudaMemcpyAsync(mv_xy + 2 * nb * first, ctx->d_mv.as<int16_t>() + 2 * nb * first, nb * cnt * 4,
                                           cudaMemcpyDeviceToHost, ctx->d2h_stream));
        if (cost)
            OFPSB_CUDA_TRY(cudaMemcpyAsync(cost + nb * first, ctx->d_cost.as<uint32_t>() + nb * first, nb * cnt * 4,
                                           cudaMemcpyDeviceToHost, ctx->d2h_stream));
        if (entries)
            OFPSB_CUDA_TRY(cudaMemcpyAsync(entries + nb * first, ctx->d_entries.as<ofps_mv>() + nb * first,
                                           nb * cnt * sizeof(ofps_mv), cudaMemcpyDeviceToHost, ctx->d2h_stream));
    }
    OFPSB_CUDA_TRY(cudaStreamSynchronize(ctx->d2h_stream));
    OFPSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return OFPSB_OK;
}

int ofpsb_block_match(ofpsb_ctx* ctx, const uint8_t* prev, const uint8_t* cur, int w, int h, int stride, int block,
                      int range, int metric, int16_t* mv_xy, uint32_t* cost, ofps_mv* entries, size_t* n_blocks)
{
    return ofpsb_block_match_batch(ctx, prev, cur, w, h, stride, 0, 1, block, range, metric, mv_xy, cost, entries,
                                   n_blocks);
}

// ----------------------------------------------------------------------------------- densifier
int ofpsb_densify_dev(ofpsb_ctx* ctx, const ofps_mv* d_entries, size_t n, size_t gw, size_t gh, float* d_field_xy,
                      float* d_counts)
{
    OFPSB_ENTER(ctx);
    if (!d_field_xy || (n && !d_entries)) {
        set_error("densify: null pointer");
        return OFPSB_E_INVALID;
    }
    return launch_densify(d_entries, n, gw, gh, d_field_xy, d_counts, ctx->densify, ctx->stream, &ctx->launches,
                          ctx->opt_densify_path);
}

int ofpsb_densify(ofpsb_ctx* ctx, const ofps_mv* entries, size_t n, size_t gw, size_t gh, float* field_xy, float* counts)
{
    OFPSB_ENTER(ctx);
    if (!field_xy || (n && !entries)) {
        set_error("densify: null pointer");
        return OFPSB_E_INVALID;
    }
    if (gw == 0 || gh == 0 || gw > (1u << 24) || gh > (1u << 24)) {
        set_error("densify: invalid grid %zux%zu", gw, gh);
        return OFPSB_E_INVALID;
    }
    const size_t cells = gw * gh;
    if (int rc = ctx->d_entries.reserve(n * sizeof(ofps_mv))) return rc;
    if (int rc = ctx->d_field.reserve(cells * 8)) return rc;
    if (counts) if (int rc = ctx->d_counts.reserve(cells * 8)) return rc;
    if (n) OFPSB_CUDA_TRY(cudaMemcpyAsync(ctx->d_entries.ptr, entries, n * sizeof(ofps_mv), cudaMemcpyHostToDevice, ctx->stream));
    if (int rc = launch_densify(ctx->d_entries.as<ofps_mv>(), n, gw, gh, ctx->d_field.as<float>(),
                                counts ? ctx->d_counts.as<float>() : nullptr, ctx->densify, ctx->stream, &ctx->launches,
                                ctx->opt_densify_path))
        return rc;
    OFPSB_CUDA_TRY(cudaMemcpyAsync(field_xy, ctx->d_field.ptr, cells * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (counts) OFPSB_CUDA_TRY(cudaMemcpyAsync(counts, ctx->d_counts.ptr, cells * 8, cudaMemcpyDeviceToHost, ctx->stream));
    OFPSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return OFPSB_OK;
}

int ofpsb_interpolate_empty_cells(float* sums_xy, float* counts_xy, size_t w, size_t h)
{
    if (!sums_xy || !counts_xy || w == 0 || h == 0 || w > (1u << 24) || h > (1u << 24)) {
        set_error("interpolate_empty_cells: invalid arguments");
        return OFPSB_E_INVALID;
    }
    try {
        interpolate_empty_cells_host(sums_xy, counts_xy, w, h);
    } catch (const std::bad_alloc&) {   // no exception may cross the C ABI
        set_error("interpolate_empty_cells: out of host memory (%zux%zu cells)", w, h);
        return OFPSB_E_NOMEM;
    }
    return OFPSB_OK;
}

int ofpsb_flow_field(ofpsb_ctx* ctx, const ofps_mv* entries, size_t n, size_t w, size_t h, float* field_xy)
{
    OFPSB_ENTER(ctx);
    if (!field_xy || (n && !entries) || w == 0 || h == 0 || w > (1u << 24) || h > (1u << 24)) {
        set_error("flow_field: invalid arguments (%zux%zu)", w, h);
        return OFPSB_E_INVALID;
    }
    const size_t cells = w * h;
    if (int rc = ctx->d_entries.reserve(n * sizeof(ofps_mv))) return rc;
    if (int rc = ctx->d_field.reserve(cells * 8)) return rc;
    if (int rc = ctx->d_counts.reserve(cells * 8)) return rc;
    if (n) OFPSB_CUDA_TRY(cudaMemcpyAsync(ctx->d_entries.ptr, entries, n * sizeof(ofps_mv), cudaMemcpyHostToDevice, ctx->stream));
    if (int rc = launch_densify(ctx->d_entries.as<ofps_mv>(), n, w, h, ctx->d_field.as<float>(), ctx->d_counts.as<float>(),
                                ctx->densify, ctx->stream, &ctx->launches, ctx->opt_densify_path, /*raw=*/1))
        return rc;
    try {
        std::vector<float> counts(2 * cells);
        OFPSB_CUDA_TRY(cudaMemcpyAsync(field_xy, ctx->d_field.ptr, cells * 8, cudaMemcpyDeviceToHost, ctx->stream));
        OFPSB_CUDA_TRY(cudaMemcpyAsync(counts.data(), ctx->d_counts.ptr, cells * 8, cudaMemcpyDeviceToHost, ctx->stream));
        OFPSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        interpolate_empty_cells_host(field_xy, counts.data(), w, h);
        for (size_t i = 0; i < 2 * cells; i++) field_xy[i] = field_xy[i] / counts[i];   // MotionField::from (:297-308)
    } catch (const std::bad_alloc&) {   // no exception may cross the C ABI
        set_error("flow_field: out of host memory (%zux%zu cells)", w, h);
        return OFPSB_E_NOMEM;
    }
    return OFPSB_OK;
}

// ------------------------------------------------------------------------------------ detector
int ofpsb_block_dim(float min_size, size_t subdivide, size_t* dim)
{
    if (!dim) {
        set_error("block_dim: null output pointer");
        return OFPSB_E_INVALID;
    }
    *dim = block_dim_host(min_size, subdivide);
    return OFPSB_OK;
}

int ofpsb_detect_block_motion_dev(ofpsb_ctx* ctx, const ofps_mv* d_entries, size_t n, float min_size, size_t subdivide,
                                  float target_motion, int* has_motion, size_t* area, size_t* dim, float* field_xy,
                                  size_t field_cap_cells)
{
    OFPSB_ENTER(ctx);
    if (n && !d_entries) {
        set_error("detect_block_motion: null entries");
        return OFPSB_E_INVALID;
    }
    return detect_from_device_entries(ctx, d_entries, n, min_size, subdivide, target_motion, has_motion, area, dim,
                                      field_xy, field_cap_cells);
}

int ofpsb_detect_block_motion(ofpsb_ctx* ctx, const ofps_mv* entries, size_t n, float min_size, size_t subdivide,
                              float target_motion, int* has_motion, size_t* area, size_t* dim, float* field_xy,
                              size_t field_cap_cells)
{
    OFPSB_ENTER(ctx);
    if (n && !entries) {
        set_error("detect_block_motion: null entries");
        return OFPSB_E_INVALID;
    }
    if (int rc = ctx->d_entries.reserve(n * sizeof(ofps_mv))) return rc;
    if (n) OFPSB_CUDA_TRY(cudaMemcpyAsync(ctx->d_entries.ptr, entries, n * sizeof(ofps_mv), cudaMemcpyHostToDevice, ctx->stream));
    return detect_from_device_entries(ctx, ctx->d_entries.as<ofps_mv>(), n, min_size, subdivide, target_motion,
                                      has_motion, area, dim, field_xy, field_cap_cells);
}

// ----------------------------------------------------------------------------------- estimator
int ofpsb_almeida_dev(ofpsb_ctx* ctx, const ofps_mv* d_entries, size_t n, float aspect, float fov_y_deg, int use_ransac,
                      size_t num_iters, float inlier_angle_deg, size_t ransac_samples, uint64_t seed, float quat_wijk[4])
{
    OFPSB_ENTER(ctx);
    if (!quat_wijk || (n && !d_entries)) {
        set_error("almeida: null pointer");
        return OFPSB_E_INVALID;
    }
    if (int rc = ctx->d_misc.reserve(256)) return rc;
    if (int rc = ctx->h_misc.reserve(256)) return rc;
    float* d_quat = ctx->d_misc.as<float>();
    if (int rc = launch_almeida(d_entries, n, aspect, fov_y_deg, use_ransac, num_iters, inlier_angle_deg, ransac_samples,
                                seed, d_quat, ctx->almeida, ctx->sm_count, ctx->stream, &ctx->launches))
        return rc;
    OFPSB_CUDA_TRY(cudaMemcpyAsync(ctx->h_misc.ptr, d_quat, 16, cudaMemcpyDeviceToHost, ctx->stream));
    OFPSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    memcpy(quat_wijk, ctx->h_misc.ptr, 16);
    return OFPSB_OK;
}

int ofpsb_almeida(ofpsb_ctx* ctx, const ofps_mv* entries, size_t n, float aspect, float fov_y_deg, int use_ransac,
                  size_t num_iters, float inlier_angle_deg, size_t ransac_samples, uint64_t seed, float quat_wijk[4])
{
    OFPSB_ENTER(ctx);
    if (!quat_wijk || (n && !entries)) {
        set_error("almeida: null pointer");
        return OFPSB_E_INVALID;
    }
    if (int rc = ctx->d_entries.reserve(n * sizeof(ofps_mv))) return rc;
    if (n) OFPSB_CUDA_TRY(cudaMemcpyAsync(ctx->d_entries.ptr, entries, n * sizeof(ofps_mv), cudaMemcpyHostToDevice, ctx->stream));
    return ofpsb_almeida_dev(ctx, ctx->d_entries.as<ofps_mv>(), n, aspect, fov_y_deg, use_ransac, num_iters,
                             inlier_angle_deg, ransac_samples, seed, quat_wijk);
}

// ------------------------------------------------------------------------- fused per-frame path
int ofpsb_frame_detect(ofpsb_ctx* ctx, const uint8_t* prev, const uint8_t* cur, int w, int h, int stride, int block,
                       int range, int metric, float min_size, size_t subdivide, float target_motion, ofps_mv* entries,
                       size_t* n_blocks, int* has_motion, size_t* area, size_t* dim, float* field_xy,
                       size_t field_cap_cells)
{
    OFPSB_ENTER(ctx);
    if (!prev || !cur || w <= 0 || h <= 0 || stride < w || block <= 0) {
        set_error("frame_detect: invalid arguments (w=%d h=%d stride=%d block=%d)", w, h, stride, block);
        return OFPSB_E_INVALID;
    }
    const size_t frame_bytes = (size_t)stride * h;
    const size_t nb = (size_t)(w / block) * (size_t)(h / block);
    if (n_blocks) *n_blocks = nb;
    if (int rc = ctx->d_frames.reserve(2 * frame_bytes)) return rc;
    if (int rc = ctx->d_entries.reserve(nb * sizeof(ofps_mv))) return rc;
    uint8_t* d_prev = ctx->d_frames.as<uint8_t>();
    uint8_t* d_cur = d_prev + frame_bytes;
    OFPSB_CUDA_TRY(cudaMemcpyAsync(d_prev, prev, frame_bytes, cudaMemcpyHostToDevice, ctx->stream));
    OFPSB_CUDA_TRY(cudaMemcpyAsync(d_cur, cur, frame_bytes, cudaMemcpyHostToDevice, ctx->stream));
    BlockMatchParams p;
    if (int rc = fill_params(p, d_prev, d_cur, w, h, stride, frame_bytes, 1, block, range, metric, nullptr, nullptr,
                             ctx->d_entries.as<ofps_mv>()))
        return rc;
    if (int rc = launch_bm(ctx, p)) return rc;
    if (entries && nb)
        OFPSB_CUDA_TRY(cudaMemcpyAsync(entries, ctx->d_entries.ptr, nb * sizeof(ofps_mv), cudaMemcpyDeviceToHost, ctx->stream));
    return detect_from_device_entries(ctx, ctx->d_entries.as<ofps_mv>(), nb, min_size, subdivide, target_motion,
                                      has_motion, area, dim, field_xy, field_cap_cells);
}

// ------------------------------------------------------------- cv-decoder dense-flow front end
int ofpsb_mfield_size(size_t frame_w, size_t frame_h, size_t ar_x, size_t ar_y, size_t max_w, size_t max_h, size_t* dx,
                      size_t* dy)
{
    if (!dx || !dy || frame_w == 0 || frame_h == 0 || ar_x == 0 || ar_y == 0) {
        set_error("mfield_size: invalid arguments");
        return OFPSB_E_INVALID;
    }
    // cv-decoder/src/lib.rs:90-118 (usize arithmetic, truncating division)
    const size_t r0 = frame_w * ar_x, r1 = frame_h * ar_y;
    const size_t w = max_w < frame_w ? max_w : frame_w, h = max_h < frame_h ? max_h : frame_h;
    const size_t wb1 = w * r1 / r0, hb0 = h * r0 / r1;
    if (w < hb0) {
        *dx = w;
        *dy = wb1;
    } else {
        *dx = hb0;
        *dy = h;
    }
    return OFPSB_OK;
}

int ofpsb_frame_convert_dev(ofpsb_ctx* ctx, const uint8_t* d_src, int w, int h, int stride, int channels, int rgb_order,
                            uint8_t* d_gray, int gray_stride, uint8_t* d_rgba)
{
    OFPSB_ENTER(ctx);
    if (!d_src) {
        set_error("frame_convert: null source");
        return OFPSB_E_INVALID;
    }
    return launch_frame_convert(d_src, w, h, stride, channels, rgb_order, d_gray, gray_stride, d_rgba, ctx->stream,
                                &ctx->launches);
}

int ofpsb_frame_convert(ofpsb_ctx* ctx, const uint8_t* src, int w, int h, int stride, int channels, int rgb_order,
                        uint8_t* gray, uint8_t* rgba)
{
    OFPSB_ENTER(ctx);
    if (!src || w <= 0 || h <= 0 || (channels != 3 && channels != 4) || stride < w * channels) {
        set_error("frame_convert: invalid arguments (w=%d h=%d stride=%d channels=%d)", w, h, stride, channels);
        return OFPSB_E_INVALID;
    }
    const size_t npix = (size_t)w * h;
    const int gstride = (w + 15) & ~15;               // 16-byte aligned device pitches: the kernel's vector path
    const int sstride = (w * channels + 15) & ~15;
    if (int rc = ctx->d_cv_src.reserve((size_t)sstride * h)) return rc;
    if (gray) if (int rc = ctx->d_cv_gray.reserve((size_t)gstride * h)) return rc;
    if (rgba) if (int rc = ctx->d_cv_rgba.reserve(npix * 4)) return rc;
    OFPSB_CUDA_TRY(cudaMemcpy2DAsync(ctx->d_cv_src.ptr, (size_t)sstride, src, (size_t)stride, (size_t)w * channels, (size_t)h,
                                     cudaMemcpyHostToDevice, ctx->stream));
    if (int rc = launch_frame_convert(ctx->d_cv_src.as<uint8_t>(), w, h, sstride, channels, rgb_order,
                                      gray ? ctx->d_cv_gray.as<uint8_t>() : nullptr, gstride,
                                      rgba ? ctx->d_cv_rgba.as<uint8_t>() : nullptr, ctx->stream, &ctx->launches))
        return rc;
    if (gray)
        OFPSB_CUDA_TRY(cudaMemcpy2DAsync(gray, (size_t)w, ctx->d_cv_gray.ptr, (size_t)gstride, (size_t)w, (size_t)h,
                                         cudaMemcpyDeviceToHost, ctx->stream));
    if (rgba) OFPSB_CUDA_TRY(cudaMemcpyAsync(rgba, ctx->d_cv_rgba.ptr, npix * 4, cudaMemcpyDeviceToHost, ctx->stream));
    OFPSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return OFPSB_OK;
}

int ofpsb_frame_resize_dev(ofpsb_ctx* ctx, const uint8_t* d_src, int sw, int sh, int stride, int channels, uint8_t* d_dst,
                           int dw, int dh, int dst_stride)
{
    OFPSB_ENTER(ctx);
    if (!d_src || !d_dst) {
        set_error("frame_resize: null pointer");
        return OFPSB_E_INVALID;
    }
    return launch_frame_resize(d_src, sw, sh, stride, channels, d_dst, dw, dh, dst_stride, ctx->stream, &ctx->launches);
}

int ofpsb_frame_resize(ofpsb_ctx* ctx, const uint8_t* src, int sw, int sh, int stride, int channels, uint8_t* dst, int dw,
                       int dh)
{
    OFPSB_ENTER(ctx);
    if (!src || !dst || sw <= 0 || sh <= 0 || dw <= 0 || dh <= 0 || (channels != 3 && channels != 4) ||
        stride < sw * channels) {
        set_error("frame_resize: invalid arguments (%dx%d -> %dx%d, channels=%d)", sw, sh, dw, dh, channels);
        return OFPSB_E_INVALID;
    }
    const size_t out_bytes = (size_t)dw * dh * channels;
    if (int rc = ctx->d_cv_src.reserve((size_t)stride * sh)) return rc;
    if (int rc = ctx->d_cv_small.reserve(out_bytes)) return rc;
    OFPSB_CUDA_TRY(cudaMemcpyAsync(ctx->d_cv_src.ptr, src, (size_t)stride * sh, cudaMemcpyHostToDevice, ctx->stream));
    if (int rc = launch_frame_resize(ctx->d_cv_src.as<uint8_t>(), sw, sh, stride, channels, ctx->d_cv_small.as<uint8_t>(), dw,
                                     dh, dw * channels, ctx->stream, &ctx->launches))
        return rc;
    OFPSB_CUDA_TRY(cudaMemcpyAsync(dst, ctx->d_cv_small.ptr, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    OFPSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return OFPSB_OK;
}

int ofpsb_contrast_mask_dev(ofpsb_ctx* ctx, const uint8_t* d_gray, int w, int h, int stride, uint8_t* d_mask,
                            int mask_stride)
{
    OFPSB_ENTER(ctx);
    if (!d_gray || !d_mask) {
        set_error("contrast_mask: null pointer");
        return OFPSB_E_INVALID;
    }
    return launch_contrast_mask(d_gray, w, h, stride, d_mask, mask_stride, ctx->stream, &ctx->launches);
}

namespace {
// gray (host, `stride` bytes per row) -> device plane with 16-byte aligned rows; returns its pitch
int upload_gray(ofpsb_ctx* ctx, const uint8_t* gray, int w, int h, int stride, int* pitch)
{
    const int gp = (w + 15) & ~15;
    if (int rc = ctx->d_cv_gray.reserve((size_t)gp * h)) return rc;
    OFPSB_CUDA_TRY(cudaMemcpy2DAsync(ctx->d_cv_gray.ptr, (size_t)gp, gray, (size_t)stride, (size_t)w, (size_t)h,
                                     cudaMemcpyHostToDevice, ctx->stream));
    *pitch = gp;
    return OFPSB_OK;
}

// count -> host, then the first min(count, cap) entries
int fetch_entries(ofpsb_ctx* ctx, const ofps_mv* d_entries, ofps_mv* entries, size_t cap, size_t* n, const char* what)
{
    if (int rc = ctx->h_misc.reserve(256)) return rc;
    OFPSB_CUDA_TRY(cudaMemcpyAsync(ctx->h_misc.ptr, ctx->d_cv_count.ptr, 8, cudaMemcpyDeviceToHost, ctx->stream));
    OFPSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    const size_t cnt = (size_t)*ctx->h_misc.as<unsigned long long>();
    if (n) *n = cnt;
    const size_t take = cnt < cap ? cnt : cap;
    if (entries && take) {
        OFPSB_CUDA_TRY(cudaMemcpyAsync(entries, d_entries, take * sizeof(ofps_mv), cudaMemcpyDeviceToHost, ctx->stream));
        OFPSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    if (cnt > cap) {
        set_error("%s: %zu entries, buffer holds %zu", what, cnt, cap);
        return OFPSB_E_CAPACITY;
    }
    return OFPSB_OK;
}
}  // namespace

int ofpsb_contrast_mask(ofpsb_ctx* ctx, const uint8_t* gray, int w, int h, int stride, uint8_t* mask)
{
    OFPSB_ENTER(ctx);
    if (!gray || !mask || w <= 0 || h <= 0 || stride < w) {
        set_error("contrast_mask: invalid arguments (w=%d h=%d stride=%d)", w, h, stride);
        return OFPSB_E_INVALID;
    }
    int gp = 0;
    if (int rc = upload_gray(ctx, gray, w, h, stride, &gp)) return rc;
    if (int rc = ctx->d_cv_mask.reserve((size_t)gp * h)) return rc;
    if (int rc = launch_contrast_mask(ctx->d_cv_gray.as<uint8_t>(), w, h, gp, ctx->d_cv_mask.as<uint8_t>(), gp, ctx->stream,
                                      &ctx->launches))
        return rc;
    OFPSB_CUDA_TRY(cudaMemcpy2DAsync(mask, (size_t)w, ctx->d_cv_mask.ptr, (size_t)gp, (size_t)w, (size_t)h,
                                     cudaMemcpyDeviceToHost, ctx->stream));
    OFPSB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return OFPSB_OK;
}

int ofpsb_flow_entries_dev(ofpsb_ctx* ctx, const float* d_flow_xy, size_t flow_stride, const uint8_t* d_mask,
                           size_t mask_stride, int w, int h, size_t gw, size_t gh, ofps_mv* d_entries, size_t cap, size_t* n)
{
    OFPSB_ENTER(ctx);
    if (!d_flow_xy || (!d_entries && cap)) {
        set_error("flow_entries: null pointer");
        return OFPSB_E_INVALID;
    }
    if (int rc = ctx->d_cv_count.reserve(256)) return rc;
    if (int rc = launch_flow_entries(d_flow_xy, flow_stride, d_mask, mask_stride, w, h, gw, gh, d_entries, cap,
                                     ctx->d_cv_count.as<unsigned long long>(), ctx->flow, ctx->stream, &ctx->launches))
        return rc;
    return fetch_entries(ctx, nullptr, nullptr, cap, n, "flow_entries");
}

namespace {
int flow_entries_host(ofpsb_ctx* ctx, const float* flow_xy, const uint8_t* d_mask, size_t mask_stride, int w, int h,
                      size_t gw, size_t gh, ofps_mv* entries, size_t cap, size_t* n, const char* what)
{
    const size_t npix = (size_t)w * h;
    const size_t max_out = gw ? gw * gh : npix;
    const size_t dev_cap = cap < max_out ? cap : max_out;
    if (int rc = ctx->d_cv_flow.reserve(npix * 8)) return rc;
    if (int rc = ctx->d_entries.reserve(dev_cap * sizeof(ofps_mv))) return rc;
    if (int rc = ctx->d_cv_count.reserve(256)) return rc;
    OFPSB_CUDA_TRY(cudaMemcpyAsync(ctx->d_cv_flow.ptr, flow_xy, npix * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (int rc = launch_flow_entries(ctx->d_cv_flow.as<float>(), 2 * (size_t)w, d_mask, mask_stride, w, h, gw, gh,
                                     ctx->d_entries.as<ofps_mv>(), dev_cap, ctx->d_cv_count.as<unsigned long long>(),
                                     ctx->flow, ctx->stream, &ctx->launches))
        return rc;
    return fetch_entries(ctx, ctx->d_entries.as<ofps_mv>(), entries, cap, n, what);
}
}  // namespace

int ofpsb_flow_entries(ofpsb_ctx* ctx, const float* flow_xy, const uint8_t* mask, int w, int h, size_t gw, size_t gh,
                       ofps_mv* entries, size_t cap, size_t* n)
{
    OFPSB_ENTER(ctx);
    if (!flow_xy || (!entries && cap) || w <= 0 || h <= 0) {
        set_error("flow_entries: invalid arguments (w=%d h=%d)", w, h);
        return OFPSB_E_INVALID;
    }
    const uint8_t* d_mask = nullptr;
    if (mask) {
        if (int rc = ctx->d_cv_mask.reserve((size_t)w * h)) return rc;
        OFPSB_CUDA_TRY(cudaMemcpyAsync(ctx->d_cv_mask.ptr, mask, (size_t)w * h, cudaMemcpyHostToDevice, ctx->stream));
        d_mask = ctx->d_cv_mask.as<uint8_t>();
    }
    return flow_entries_host(ctx, flow_xy, d_mask, (size_t)w, w, h, gw, gh, entries, cap, n, "flow_entries");
}

int ofpsb_cv_flow_frame(ofpsb_ctx* ctx, const uint8_t* gray, int gray_stride, const float* flow_xy, int w, int h,
                        int use_mask, size_t gw, size_t gh, ofps_mv* entries, size_t cap, size_t* n)
{
    OFPSB_ENTER(ctx);
    if (!flow_xy || (!entries && cap) || w <= 0 || h <= 0 || (use_mask && (!gray || gray_stride < w))) {
        set_error("cv_flow_frame: invalid arguments (w=%d h=%d gray_stride=%d)", w, h, gray_stride);
        return OFPSB_E_INVALID;
    }
    const uint8_t* d_mask = nullptr;
    int gp = 0;
    if (use_mask) {
        if (int rc = upload_gray(ctx, gray, w, h, gray_stride, &gp)) return rc;
        if (int rc = ctx->d_cv_mask.reserve((size_t)gp * h)) return rc;
        if (int rc = launch_contrast_mask(ctx->d_cv_gray.as<uint8_t>(), w, h, gp, ctx->d_cv_mask.as<uint8_t>(), gp,
                                          ctx->stream, &ctx->launches))
            return rc;
        d_mask = ctx->d_cv_mask.as<uint8_t>();
    }
    return flow_entries_host(ctx, flow_xy, d_mask, (size_t)gp, w, h, gw, gh, entries, cap, n, "cv_flow_frame");
}

// --------------------------------------------------------------------------- interchange files
int ofpsb_mvec_append(const char* path, const ofps_mv* entries, size_t n, int truncate)
{
    if (!path || (n && !entries) || n > 0xFFFFFFFFull) {
        set_error("mvec_append: invalid arguments");
        return OFPSB_E_INVALID;
    }
    FILE* f = fopen(path, truncate ? "wb" : "ab");
    if (!f) {
        set_error("mvec_append: cannot open %s", path);
        return OFPSB_E_IO;
    }
    // motion-extract/src/main.rs:24-32: u32 LE count, then count x (px, py, mx, my) f32 LE
    const uint32_t cnt = (uint32_t)n;
    const unsigned char hdr[4] = {(unsigned char)(cnt & 255), (unsigned char)((cnt >> 8) & 255),
                                  (unsigned char)((cnt >> 16) & 255), (unsigned char)((cnt >> 24) & 255)};
    bool ok = fwrite(hdr, 1, 4, f) == 4;
    if (ok && n) ok = fwrite(entries, sizeof(ofps_mv), n, f) == n;   // little-endian host (x86-64 / aarch64)
    ok = (fclose(f) == 0) && ok;
    if (!ok) {
        set_error("mvec_append: write to %s failed", path);
        return OFPSB_E_IO;
    }
    return OFPSB_OK;
}

int ofpsb_mvec_read(const char* path, size_t frame_index, ofps_mv* entries, size_t cap, size_t* n)
{
    if (!path || !n) {
        set_error("mvec_read: invalid arguments");
        return OFPSB_E_INVALID;
    }
    FILE* f = fopen(path, "rb");
    if (!f) {
        set_error("mvec_read: cannot open %s", path);
        return OFPSB_E_IO;
    }
    int rc = OFPSB_OK;
    for (size_t frame = 0;; frame++) {
        unsigned char hdr[4];
        if (fread(hdr, 1, 4, f) != 4) {
            set_error("mvec_read: %s has no frame %zu", path, frame_index);
            rc = OFPSB_E_IO;
            break;
        }
        const size_t cnt = (size_t)hdr[0] | ((size_t)hdr[1] << 8) | ((size_t)hdr[2] << 16) | ((size_t)hdr[3] << 24);
        if (frame == frame_index) {
            *n = cnt;
            if (entries) {
                if (cap < cnt) {
                    set_error("mvec_read: frame holds %zu entries, buffer %zu", cnt, cap);
                    rc = OFPSB_E_CAPACITY;
                } else if (cnt && fread(entries, sizeof(ofps_mv), cnt, f) != cnt) {
                    set_error("mvec_read: truncated frame %zu in %s", frame, path);
                    rc = OFPSB_E_IO;
                }
            }
            break;
        }
        if (fseek(f, (long)(cnt * sizeof(ofps_mv)), SEEK_CUR) != 0) {
            set_error("mvec_read: seek failed in %s", path);
            rc = OFPSB_E_IO;
            break;
        }
    }
    fclose(f);
    return rc;
}

int ofpsb_flo_write(const char* path, const float* field_xy, size_t w, size_t h)
{
    if (!path || !field_xy || w == 0 || h == 0 || w > 0x7FFFFFFF || h > 0x7FFFFFFF) {
        set_error("flo_write: invalid arguments");
        return OFPSB_E_INVALID;
    }
    FILE* f = fopen(path, "wb");
    if (!f) {
        set_error("flo_write: cannot open %s", path);
        return OFPSB_E_IO;
    }
    const float tag = 202021.25f;   // "PIEH"
    const int32_t wi = (int32_t)w, hi = (int32_t)h;
    bool ok = fwrite(&tag, 4, 1, f) == 1 && fwrite(&wi, 4, 1, f) == 1 && fwrite(&hi, 4, 1, f) == 1 &&
              fwrite(field_xy, 8, w * h, f) == w * h;
    ok = (fclose(f) == 0) && ok;
    if (!ok) {
        set_error("flo_write: write to %s failed", path);
        return OFPSB_E_IO;
    }
    return OFPSB_OK;
}

}  // extern "C"
#pragma GCC visibility pop
