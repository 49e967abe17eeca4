// Streaming block-matching decoder behind the C ABI (include/ofps_b200.h, ofpsb_stream_*): the device side of
// `Decoder::process_frame` (ofps/src/decoder.rs:45-73) for a decoder that produces its motion vectors by block matching.
//
// the drop-in decoder called the single-pair entry point on pageable memory and re-uploaded BOTH frames on every
// call (0.42 ms per 1080p pair, 4.9 Gpix/s, 10x below the batch path).  Here
//   * every frame is uploaded ONCE: the previous frame of a pair is still in HBM (ring of device frames);
//   * the caller's buffer may be pageable: it is copied into a pinned ring owned by the library by a few helper threads
//     (a single-thread memcpy of a 2 MB frame costs more than its PCIe transfer), or handed to the copy engine directly
//     when it is already page-locked;
//   * upload, kernels and the read-back of the entry list run on three streams chained by events; with
//     ofpsb_stream_submit / ofpsb_stream_collect the upload of frame i+1 overlaps the kernels and read-back of pair i
//     (ofpsb_stream_push = submit + collect, the synchronous drop-in form).
#include "common.cuh"

#include <atomic>
#include <condition_variable>
#include <cstring>
#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#define OFPSB_NT_COPY 1
#endif
#include <mutex>
#include <new>
#include <thread>
#include <vector>

using namespace ofpsb;

namespace {

// A few helper threads that split one strided host copy; they spin briefly after a job (the next frame usually follows
// within microseconds) and then sleep on a condition variable.
class CopyPool {
public:
    explicit CopyPool(int helpers)
    {
        for (int i = 0; i < helpers; i++) th_.emplace_back([this, i] { worker(i + 1); });
    }
    ~CopyPool()
    {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
            gen_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    void copy(uint8_t* dst, size_t dst_stride, const uint8_t* src, size_t src_stride, size_t row_bytes, int rows)
    {
        const int parts = (int)th_.size() + 1;
        if (parts == 1 || (size_t)rows * row_bytes < (256u << 10)) {
            part(dst, dst_stride, src, src_stride, row_bytes, 0, rows);
            return;
        }
        dst_ = dst; dst_stride_ = dst_stride; src_ = src; src_stride_ = src_stride; row_bytes_ = row_bytes; rows_ = rows;
        pending_.store(parts - 1, std::memory_order_relaxed);
        {
            std::lock_guard<std::mutex> lk(m_);
            gen_.fetch_add(1, std::memory_order_release);
        }
        if (sleepers_.load(std::memory_order_acquire) > 0) cv_.notify_all();
        part(dst, dst_stride, src, src_stride, row_bytes, 0, rows / parts);
        while (pending_.load(std::memory_order_acquire) > 0) std::this_thread::yield();
    }

private:
    // The staging frame is read next by the copy engine, never by this core: non-temporal stores keep it out of the
    // caches and skip the read-for-ownership of the destination lines (a third of the memory traffic of a plain copy;
    // measured on the pool's 16-core hosts, pageable 1080p frames: 83 -> 66 us per frame, 24.9 -> 31.4 Gpix/s).
    static void copy_stream(uint8_t* dst, const uint8_t* src, size_t n)
    {
#ifndef OFPSB_NT_COPY
        memcpy(dst, src, n);
#else
        if (n < 4096) {
            memcpy(dst, src, n);
            return;
        }
        const size_t head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
        memcpy(dst, src, head);
        dst += head; src += head; n -= head;
        size_t i = 0;
        for (; i + 64 <= n; i += 64) {
            const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
            const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 16));
            const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 32));
            const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 48));
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), a);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 16), b);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 32), c);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 48), d);
        }
        memcpy(dst + i, src + i, n - i);
        _mm_sfence();
#endif
    }
    static void part(uint8_t* dst, size_t ds, const uint8_t* src, size_t ss, size_t rb, int r0, int r1)
    {
        if (ds == rb && ss == rb) {
            copy_stream(dst + (size_t)r0 * rb, src + (size_t)r0 * rb, (size_t)(r1 - r0) * rb);
            return;
        }
        for (int r = r0; r < r1; r++) copy_stream(dst + (size_t)r * ds, src + (size_t)r * ss, rb);
    }
    void worker(int idx)
    {
        uint64_t seen = 0;
        for (;;) {
            int spins = 0;
            while (gen_.load(std::memory_order_acquire) == seen) {
                if (++spins < 20000) {
                    std::this_thread::yield();
                    continue;
                }
                std::unique_lock<std::mutex> lk(m_);
                sleepers_.fetch_add(1, std::memory_order_release);
                cv_.wait(lk, [&] { return gen_.load(std::memory_order_acquire) != seen; });
                sleepers_.fetch_sub(1, std::memory_order_release);
            }
            seen = gen_.load(std::memory_order_acquire);
            if (stop_) return;
            const int parts = (int)th_.size() + 1;
            const int r0 = (int)((long long)rows_ * idx / parts), r1 = (int)((long long)rows_ * (idx + 1) / parts);
            part(dst_, dst_stride_, src_, src_stride_, row_bytes_, r0, r1);
            pending_.fetch_sub(1, std::memory_order_release);
        }
    }
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_;
    std::atomic<uint64_t> gen_{0};
    std::atomic<int> pending_{0}, sleepers_{0};
    bool stop_ = false;
    uint8_t* dst_ = nullptr;
    const uint8_t* src_ = nullptr;
    size_t dst_stride_ = 0, src_stride_ = 0, row_bytes_ = 0;
    int rows_ = 0;
};

}  // namespace

struct ofpsb_stream {
    ofpsb_ctx* ctx = nullptr;
    int w = 0, h = 0, block = 0, range = 0, metric = 0, depth = 0, stride = 0;
    size_t frame_bytes = 0, nb = 0;
    uint8_t* d_frames = nullptr;       // depth device frames (ring)
    ofps_mv* d_entries = nullptr;      // depth entry lists (ring)
    uint8_t* h_frames = nullptr;       // depth pinned staging frames
    ofps_mv* h_entries = nullptr;      // depth pinned entry lists
    std::vector<cudaEvent_t> h2d_done, comp_done, d2h_done;
    // consecutive pairs are independent and one 1080p pair fills half the machine: they alternate between two compute
    // streams (each with its own work-list scratch) so that pair k+1 runs beside pair k
    cudaStream_t comp[2] = {nullptr, nullptr};
    BlockMatchScratch scratch[2];
    long long submitted = 0;           // frames submitted
    long long collected = 0;           // pairs collected (pair j = frames j, j+1)
    CopyPool* pool = nullptr;
};

#pragma GCC visibility push(default)
extern "C" {

void ofpsb_stream_close(ofpsb_stream* s)
{
    if (!s) return;
    DeviceGuard guard(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    if (s->comp[1]) {
        cudaStreamSynchronize(s->comp[1]);
        cudaStreamDestroy(s->comp[1]);
    }
    for (auto& sc : s->scratch) {
        sc.sums.release();
        sc.worklist.release();
        if (sc.ev_listed) cudaEventDestroy(sc.ev_listed);
        if (sc.h_listed) cudaFreeHost(sc.h_listed);
    }
    cudaStreamSynchronize(s->ctx->copy_stream);
    cudaStreamSynchronize(s->ctx->d2h_stream);
    delete s->pool;
    for (auto* v : {&s->h2d_done, &s->comp_done, &s->d2h_done})
        for (cudaEvent_t e : *v)
            if (e) cudaEventDestroy(e);
    if (s->d_frames) cudaFree(s->d_frames);
    if (s->d_entries) cudaFree(s->d_entries);
    if (s->h_frames) cudaFreeHost(s->h_frames);
    if (s->h_entries) cudaFreeHost(s->h_entries);
    delete s;
}

int ofpsb_stream_open(ofpsb_ctx* ctx, int w, int h, int block, int range, int metric, int depth, ofpsb_stream** out)
{
    OFPSB_ENTER(ctx);
    if (!out || w <= 0 || h <= 0 || block <= 0 || (block & 3) || range < 0 || range > 63 || depth < 0 || depth > 64 ||
        (metric != OFPSB_METRIC_SAD && metric != OFPSB_METRIC_SSD)) {
        set_error("stream_open: invalid arguments (%dx%d block %d range %d metric %d depth %d)", w, h, block, range, metric, depth);
        return OFPSB_E_INVALID;
    }
    *out = nullptr;
    ofpsb_stream* s = new (std::nothrow) ofpsb_stream();
    if (!s) return OFPSB_E_NOMEM;
    s->ctx = ctx;
    s->w = w; s->h = h; s->block = block; s->range = range; s->metric = metric;
    s->depth = depth < 3 ? 3 : depth;   // a device frame is prev of one pair and cur of the one before
    s->stride = (w + 15) & ~15;
    s->frame_bytes = (((size_t)s->stride * h) + 255) & ~(size_t)255;
    s->nb = (size_t)(w / block) * (size_t)(h / block);
    const size_t ent_bytes = ((s->nb ? s->nb : 1) * sizeof(ofps_mv) + 255) & ~(size_t)255;
    cudaError_t e = cudaMalloc(&s->d_frames, s->frame_bytes * s->depth);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_entries, ent_bytes * s->depth);
    if (e == cudaSuccess) e = cudaHostAlloc(&s->h_frames, s->frame_bytes * s->depth, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc(&s->h_entries, ent_bytes * s->depth, cudaHostAllocDefault);
    s->comp[0] = ctx->stream;
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->comp[1], cudaStreamNonBlocking);
    for (auto* v : {&s->h2d_done, &s->comp_done, &s->d2h_done})
        for (int i = 0; i < s->depth && e == cudaSuccess; i++) {
            cudaEvent_t ev = nullptr;
            e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
            v->push_back(ev);
        }
    if (e != cudaSuccess) {
        set_error("stream_open: %s", cudaGetErrorString(e));
        ofpsb_stream_close(s);
        return e == cudaErrorMemoryAllocation ? OFPSB_E_NOMEM : OFPSB_E_CUDA;
    }
    unsigned hc = std::thread::hardware_concurrency();
    // a single core copies ~7 GB/s on the pool's hosts: a 2 MB frame needs several to stay below its PCIe time (measured:
    // 5 threads 61 us per 1080p frame against 40 us of H2D) — half the cores, at most 8 copy threads
    int helpers = s->frame_bytes >= (1u << 20) ? (hc >= 4 ? (int)(hc / 2 - 1 > 7 ? 7 : hc / 2 - 1) : 0) : 0;
    if (const char* e = getenv("OFPSB_COPY_THREADS")) {   // tuning knob: copy threads including the caller's (1 .. 32)
        const int n = atoi(e);
        if (n >= 1 && n <= 32) helpers = n - 1;
    }
    s->pool = new (std::nothrow) CopyPool(helpers);
    *out = s;
    return OFPSB_OK;
}

size_t ofpsb_stream_blocks(ofpsb_stream* s) { return s ? s->nb : 0; }

int ofpsb_stream_submit(ofpsb_stream* s, const uint8_t* frame, size_t stride)
{
    if (!s || !frame || stride < (size_t)s->w) {
        set_error("stream_submit: invalid arguments");
        return OFPSB_E_INVALID;
    }
    OFPSB_ENTER(s->ctx);
    ofpsb_ctx* ctx = s->ctx;
    const long long k = s->submitted;
    if (k - 1 - s->collected >= s->depth - 2) {
        set_error("stream_submit: %lld results outstanding — collect before submitting more (depth %d)", k - 1 - s->collected,
                  s->depth);
        return OFPSB_E_CAPACITY;
    }
    const int slot = (int)(k % s->depth);
    uint8_t* d_frame = s->d_frames + s->frame_bytes * slot;
    // device frame `slot` held frame k-depth: prev of pair k-depth (= frames k-depth, k-depth+1), computed after frame
    // k-depth+1 arrived — event comp_done[(k-depth+1) % depth]
    if (k >= s->depth) {
        // pair k-depth (frame k-depth as cur) and pair k-depth+1 (as prev) ran on different compute streams
        OFPSB_CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, s->comp_done[(size_t)slot], 0));
        OFPSB_CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, s->comp_done[(size_t)((k - s->depth + 1) % s->depth)], 0));
    }
    cudaPointerAttributes attr;
    const bool pinned = cudaPointerGetAttributes(&attr, frame) == cudaSuccess &&
                        (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged);
    cudaGetLastError();
    if (pinned) {
        OFPSB_CUDA_TRY(cudaMemcpy2DAsync(d_frame, s->stride, frame, stride, s->w, s->h, cudaMemcpyHostToDevice, ctx->copy_stream));
    } else {
        uint8_t* stage = s->h_frames + s->frame_bytes * slot;
        if (k >= s->depth) OFPSB_CUDA_TRY(cudaEventSynchronize(s->h2d_done[(size_t)slot]));   // its last upload has left
        // Nothing in flight (the caller collects every result before the next frame: the synchronous Decoder shape): the
        // frame goes in four row bands, each handed to the copy engine as soon as it is staged, so the upload of band i
        // runs behind the host copy of band i+1.  With results outstanding the next frame's staging already overlaps this
        // frame's upload, and one copy call per frame is cheaper.
        const int bands = (k - 1 - s->collected <= 0 && s->pool && s->frame_bytes >= (1u << 20)) ? 4 : 1;
        for (int bnd = 0; bnd < bands; bnd++) {
            const int r0 = (int)((long long)s->h * bnd / bands), r1 = (int)((long long)s->h * (bnd + 1) / bands);
            uint8_t* st = stage + (size_t)r0 * s->stride;
            const uint8_t* src = frame + (size_t)r0 * stride;
            if (s->pool) s->pool->copy(st, s->stride, src, stride, s->w, r1 - r0);
            else
                for (int r = 0; r < r1 - r0; r++) memcpy(st + (size_t)r * s->stride, src + (size_t)r * stride, s->w);
            OFPSB_CUDA_TRY(cudaMemcpyAsync(d_frame + (size_t)r0 * s->stride, st, (size_t)s->stride * (r1 - r0), cudaMemcpyHostToDevice,
                                           ctx->copy_stream));
        }
    }
    OFPSB_CUDA_TRY(cudaEventRecord(s->h2d_done[(size_t)slot], ctx->copy_stream));
    cudaStream_t cs = s->comp[k & 1];
    OFPSB_CUDA_TRY(cudaStreamWaitEvent(cs, s->h2d_done[(size_t)slot], 0));
    if (k > 0 && s->nb > 0) {
        const int pslot = (int)((k - 1) % s->depth);
        const size_t ent_bytes = ((s->nb * sizeof(ofps_mv)) + 255) & ~(size_t)255;
        ofps_mv* d_ent = reinterpret_cast<ofps_mv*>(reinterpret_cast<uint8_t*>(s->d_entries) + ent_bytes * slot);
        ofps_mv* h_ent = reinterpret_cast<ofps_mv*>(reinterpret_cast<uint8_t*>(s->h_entries) + ent_bytes * slot);
        // the previous frame was uploaded for the other compute stream; entry list `slot` was read back for pair
        // k-1-depth: its D2H must have left before the kernels overwrite it
        OFPSB_CUDA_TRY(cudaStreamWaitEvent(cs, s->h2d_done[(size_t)pslot], 0));
        if (k >= s->depth) OFPSB_CUDA_TRY(cudaStreamWaitEvent(cs, s->d2h_done[(size_t)slot], 0));
        BlockMatchParams p{};
        p.prev = s->d_frames + s->frame_bytes * pslot;
        p.cur = d_frame;
        p.w = s->w;
        p.strip_h = s->h;
        p.stride = s->stride;
        p.pair_stride = 0;
        p.n_pairs = 1;
        p.full_h = s->h;
        p.block = s->block;
        p.range = s->range;
        p.metric = s->metric;
        p.nbx = s->w / s->block;
        p.nby = s->h / s->block;
        p.entries = d_ent;
        // statistics / profiling options address the context's own scratch: those launches stay on its stream
        BlockMatchScratch& sc = (ctx->bm_scratch.collect_stats || ctx->bm_scratch.profile) ? ctx->bm_scratch : s->scratch[k & 1];
        if (&sc == &ctx->bm_scratch) {
            OFPSB_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, s->h2d_done[(size_t)slot], 0));
            OFPSB_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, s->h2d_done[(size_t)pslot], 0));
            if (k >= s->depth) OFPSB_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, s->d2h_done[(size_t)slot], 0));
            cs = ctx->stream;
        } else {
            sc.pruner = ctx->bm_scratch.pruner;
            sc.tile_h = ctx->bm_scratch.tile_h;
            sc.adaptive = ctx->bm_scratch.adaptive;
        }
        if (int rc = launch_block_match_ctx(ctx, p, sc, cs)) return rc;
        OFPSB_CUDA_TRY(cudaEventRecord(s->comp_done[(size_t)slot], cs));
        OFPSB_CUDA_TRY(cudaStreamWaitEvent(ctx->d2h_stream, s->comp_done[(size_t)slot], 0));
        OFPSB_CUDA_TRY(cudaMemcpyAsync(h_ent, d_ent, s->nb * sizeof(ofps_mv), cudaMemcpyDeviceToHost, ctx->d2h_stream));
        OFPSB_CUDA_TRY(cudaEventRecord(s->d2h_done[(size_t)slot], ctx->d2h_stream));
    } else {
        OFPSB_CUDA_TRY(cudaEventRecord(s->comp_done[(size_t)slot], cs));
    }
    s->submitted = k + 1;
    return OFPSB_OK;
}

int ofpsb_stream_collect(ofpsb_stream* s, ofps_mv* entries, size_t* n_entries)
{
    if (!s) {
        set_error("stream_collect: null stream");
        return OFPSB_E_INVALID;
    }
    if (n_entries) *n_entries = 0;
    if (s->collected >= s->submitted - 1) return OFPSB_OK;   // nothing outstanding (first frame: no pair yet)
    OFPSB_ENTER(s->ctx);
    const long long j = s->collected;                         // pair j = frames j, j+1 -> slot of frame j+1
    const int slot = (int)((j + 1) % s->depth);
    if (s->nb > 0) {
        OFPSB_CUDA_TRY(cudaEventSynchronize(s->d2h_done[(size_t)slot]));
        const ofps_mv* h_ent = reinterpret_cast<const ofps_mv*>(reinterpret_cast<const uint8_t*>(s->h_entries) +
                                                                (((s->nb * sizeof(ofps_mv)) + 255) & ~(size_t)255) * slot);
        if (entries) memcpy(entries, h_ent, s->nb * sizeof(ofps_mv));
    }
    if (n_entries) *n_entries = s->nb;
    s->collected = j + 1;
    return OFPSB_OK;
}

int ofpsb_stream_push(ofpsb_stream* s, const uint8_t* frame, size_t stride, ofps_mv* entries, size_t* n_entries)
{
    if (n_entries) *n_entries = 0;
    if (int rc = ofpsb_stream_submit(s, frame, stride)) return rc;
    return ofpsb_stream_collect(s, entries, n_entries);
}

}  // extern "C"
#pragma GCC visibility pop
