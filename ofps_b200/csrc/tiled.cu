// Spatial tiling of ONE large frame pair over the GPUs of a node, behind the C ABI (include/ofps_b200.h, ofpsb_tiled_*).
//
// SURVEY.md §8e / north star: "spatial tiles of one large frame shard across the 8 GPUs, NVLink only for the halo rows at
// the seams".  One rank (process or thread) per GPU owns a horizontal strip of whole block rows of every frame, in ITS
// memory.  Blocks never straddle strips, so the current frame needs nothing from the neighbours; the search window of the
// previous frame reaches `range` rows into the strips above and below.  Round 1 moved those rows with a grouped NCCL
// send/recv per pair from Python: ~150 us of latency around ~18 us of work, negative scaling.  Here there is no exchange
// step at all:
//   * every rank exports its frame buffer once (cudaIpcGetMemHandle; or the raw pointer inside one process) and maps
//     the two neighbours' buffers;
//   * the SEA kernel of a strip loads the tiles at the two seams 8 rows at a time, each group from the tensor that owns
//     it — own HBM, or the neighbour's HBM over NVLink through a tensor map encoded on the peer-mapped pointer
//     (block_match_sea.cu, peer-halo mode): the transfer is part of the tile load, tile by tile;
//   * the few blocks the SEA kernel leaves to the exhaustive kernel need the halo rows next to the strip: two peer
//     copies on a side stream run WHILE the SEA kernel runs and are joined before the work-list kernel;
//   * the whole sequence (memset, SEA, two copies, join, work list) is captured once per (slots, outputs) in a CUDA graph.
// "Frame slot ready" is a 4-byte epoch flag each rank writes into its neighbours' memory (ofpsb_tiled_publish) and the
// matcher waits for on the device (no host round trip, no collective).
#include "common.cuh"

#include <unistd.h>

#include <cstring>
#include <new>
#include <vector>

using namespace ofpsb;

namespace ofpsb {
int launch_block_match_pruned(const BlockMatchParams& p, BlockMatchScratch& scratch, int sm_count, cudaStream_t stream,
                              uint64_t* launches);
}

namespace {

constexpr int TILED_MAGIC = 0x0F95711D;

struct TiledHandle {        // what a rank tells its neighbours (ofpsb_tiled_export): fits OFPSB_TILED_HANDLE_BYTES
    cudaIpcMemHandle_t mem;
    int magic, rank, own_rows, stride, n_slots, range, device, pid;
    unsigned long long slot_bytes, flags_off;
};
static_assert(sizeof(TiledHandle) <= OFPSB_TILED_HANDLE_BYTES, "handle blob too small");

struct Neighbour {
    uint8_t* base = nullptr;      // the neighbour's allocation, mapped into this process / device
    bool ipc = false;
    int own_rows = 0, stride = 0, range = 0;
    size_t slot_bytes = 0, flags_off = 0;
    const uint8_t* own(int slot) const { return base + slot_bytes * (size_t)slot + (size_t)range * stride; }
    uint32_t* flags() const { return reinterpret_cast<uint32_t*>(base + flags_off); }
};

struct GraphKey {
    int prev, cur, n;
    void *entries, *mv, *cost;
    bool wait;
    bool operator==(const GraphKey& o) const
    {
        return prev == o.prev && cur == o.cur && n == o.n && entries == o.entries && mv == o.mv && cost == o.cost && wait == o.wait;
    }
};

// expected[slot] lives in this rank's own memory so that the wait kernel has constant arguments (graph capture)
__global__ void tiled_publish_kernel(uint32_t* a, uint32_t* b, uint32_t* expected, uint32_t v)
{
    __threadfence_system();   // the frame rows written before this kernel are visible before the flag
    if (a) *reinterpret_cast<volatile uint32_t*>(a) = v;
    if (b) *reinterpret_cast<volatile uint32_t*>(b) = v;
    *expected = v;
    __threadfence_system();
}

// a / b: flags of the first slot written by the upper / lower neighbour (two words per slot), n consecutive slots
__global__ void tiled_wait_kernel(const uint32_t* a, const uint32_t* b, const uint32_t* expected, int n)
{
    // epochs only grow; (int) difference so that a wrap after 2^31 publishes still compares correctly
    for (int i = 0; i < n; i++) {
        const uint32_t v = expected[i];
        while (a && (int)(*reinterpret_cast<const volatile uint32_t*>(a + 2 * i) - v) < 0) __nanosleep(100);
        while (b && (int)(*reinterpret_cast<const volatile uint32_t*>(b + 2 * i) - v) < 0) __nanosleep(100);
    }
    __threadfence_system();
}

// halo rows of the neighbours -> next to the strip (for the exhaustive work-list kernel): plain 16-byte loads from the
// peer-mapped pointers (a kernel node costs less set-up than two 2-D copy-engine nodes for 2 x 123 KB).
// blockIdx.y = slot; *_slot16: distance between consecutive slots in 16-byte units (own / upper / lower buffer)
__global__ void tiled_halo_kernel(uint4* __restrict__ dst_up, const uint4* __restrict__ src_up, size_t n_up,
                                  uint4* __restrict__ dst_dn, const uint4* __restrict__ src_dn, size_t n_dn, size_t own_slot16,
                                  size_t up_slot16, size_t dn_slot16)
{
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x, z = blockIdx.y;
    for (size_t i = i0; i < n_up; i += step) dst_up[z * own_slot16 + i] = src_up[z * up_slot16 + i];
    for (size_t i = i0; i < n_dn; i += step) dst_dn[z * own_slot16 + i] = src_dn[z * dn_slot16 + i];
}

}  // namespace

struct ofpsb_tiled {
    ofpsb_ctx* ctx = nullptr;
    int rank = 0, world = 1, w = 0, h = 0, block = 0, range = 0, n_slots = 0;
    int by0 = 0, nby = 0, nbx = 0, y0 = 0, rows = 0, own_rows = 0, halo_top = 0, halo_bottom = 0;
    int stride = 0;
    size_t slot_bytes = 0, flags_off = 0, alloc_bytes = 0;
    uint8_t* buf = nullptr;
    Neighbour up, down;
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    std::vector<uint32_t> epoch;                 // publishes per slot (this rank's own count = what the neighbours write)
    std::vector<std::pair<GraphKey, cudaGraphExec_t>> graphs;
    std::vector<GraphKey> seen;
    bool use_graph = true;

    uint8_t* own(int slot) const { return buf + slot_bytes * (size_t)slot + (size_t)range * stride; }
    uint32_t* flags() const { return reinterpret_cast<uint32_t*>(buf + flags_off); }
    uint32_t* expected() const { return flags() + 2 * n_slots; }   // this rank's own publish count per slot
};

namespace {

// strips of whole block rows, the first strips take the extra rows (same rule as ofps_b200/dist.py::strip_plan)
void plan_strip(int h, int block, int range, int rank, int world, ofpsb_tiled* t)
{
    const int nby_total = h / block, base = nby_total / world, extra = nby_total % world;
    int by0 = 0;
    for (int r = 0; r < rank; r++) by0 += base + (r < extra ? 1 : 0);
    t->by0 = by0;
    t->nby = base + (rank < extra ? 1 : 0);
    t->y0 = by0 * block;
    t->rows = t->nby * block;
    const bool last = rank == world - 1;
    t->own_rows = last ? h - t->y0 : t->rows;
    t->halo_top = range < t->y0 ? range : t->y0;
    const int below = h - (t->y0 + t->rows);
    t->halo_bottom = range < below ? range : below;     // last rank: the frame's remainder rows (its own memory)
}

// n_pairs > 1: pairs (prev_slot + i, prev_slot + i + 1) — consecutive slots of one stream — in ONE launch sequence
int enqueue_match(ofpsb_tiled* t, int prev_slot, int cur_slot, int n_pairs, ofps_mv* d_entries, int16_t* d_mv, uint32_t* d_cost,
                  bool wait_neighbours)
{
    ofpsb_ctx* ctx = t->ctx;
    cudaStream_t S = ctx->stream;
    BlockMatchParams p{};
    p.prev = t->own(prev_slot);
    p.cur = t->own(cur_slot);
    p.w = t->w;
    p.strip_h = t->rows;
    p.stride = t->stride;
    p.pair_stride = (long long)t->slot_bytes;
    p.n_pairs = n_pairs;
    p.halo_top = t->halo_top;
    p.halo_bottom = t->halo_bottom;
    p.y_offset = t->y0;
    p.full_h = t->h;
    p.block = t->block;
    p.range = t->range;
    p.metric = OFPSB_METRIC_SAD;
    p.nbx = t->nbx;
    p.nby = t->nby;
    p.mv_xy = d_mv;
    p.cost = d_cost;
    p.entries = d_entries;
    if (p.nbx == 0 || p.nby == 0) return OFPSB_OK;

    const bool has_up = t->up.base && t->halo_top > 0, has_down = t->down.base && t->rank + 1 < t->world && t->halo_bottom > 0;
    if (wait_neighbours && (has_up || has_down)) {
        // lock-step protocol: every rank publishes a slot once per frame it puts there, so the neighbours' epoch of the
        // slot must have reached this rank's own (kept on the device: constant kernel arguments, graph-capturable)
        tiled_wait_kernel<<<1, 1, 0, S>>>(has_up ? t->flags() + 2 * prev_slot + 0 : nullptr,
                                          has_down ? t->flags() + 2 * prev_slot + 1 : nullptr, t->expected() + prev_slot, n_pairs);
        OFPSB_CUDA_TRY(cudaGetLastError());
        ctx->launches++;
    }
    // halo rows next to the strip (for the exhaustive work-list kernel), copied from the neighbours on the side stream
    OFPSB_CUDA_TRY(cudaEventRecord(t->ev_fork, S));
    OFPSB_CUDA_TRY(cudaStreamWaitEvent(t->side, t->ev_fork, 0));
    if (has_up || has_down) {
        // strides are equal on every rank (same w): the halo rows are contiguous on both sides
        const size_t n_up = has_up ? (size_t)t->halo_top * t->stride / 16 : 0, n_dn = has_down ? (size_t)t->halo_bottom * t->stride / 16 : 0;
        tiled_halo_kernel<<<dim3(n_pairs > 8 ? 16 : 64, n_pairs), 256, 0, t->side>>>(
            reinterpret_cast<uint4*>(t->own(prev_slot) - (size_t)t->halo_top * t->stride),
            has_up ? reinterpret_cast<const uint4*>(t->up.own(prev_slot) + (size_t)(t->up.own_rows - t->halo_top) * t->up.stride) : nullptr,
            n_up, reinterpret_cast<uint4*>(t->own(prev_slot) + (size_t)t->own_rows * t->stride),
            has_down ? reinterpret_cast<const uint4*>(t->down.own(prev_slot)) : nullptr, n_dn, t->slot_bytes / 16,
            has_up ? t->up.slot_bytes / 16 : 0, has_down ? t->down.slot_bytes / 16 : 0);
        OFPSB_CUDA_TRY(cudaGetLastError());
        ctx->launches++;
    }
    OFPSB_CUDA_TRY(cudaEventRecord(t->ev_join, t->side));

    SeaPeer peer{};
    peer.own_rows = t->own_rows;
    if (has_up) {
        peer.up = t->up.own(prev_slot);
        peer.up_rows = t->up.own_rows;
        peer.up_stride = t->up.stride;
        peer.up_pair_stride = (long long)t->up.slot_bytes;
    }
    if (has_down) {
        peer.down = t->down.own(prev_slot);
        peer.down_rows = t->down.own_rows;
        peer.down_stride = t->down.stride;
        peer.down_pair_stride = (long long)t->down.slot_bytes;
    }
    int rc = 1;
    if (ctx->opt_block_match_prune && ctx->opt_block_match_kernel == 0 && ctx->bm_scratch.pruner == 0)
        rc = launch_block_match_sea(p, ctx->bm_scratch, ctx->sm_count, S, &ctx->launches, &peer, t->ev_join);
    if (rc == 1) {
        // geometry without a SEA instance (+-32, SSD is not offered here): halo rows first, then the strip kernels
        OFPSB_CUDA_TRY(cudaStreamWaitEvent(S, t->ev_join, 0));
        rc = ctx->opt_block_match_prune ? launch_block_match_pruned(p, ctx->bm_scratch, ctx->sm_count, S, &ctx->launches) : 1;
        if (rc == 1) rc = launch_block_match(p, S, &ctx->launches, 0);
    }
    return rc;
}

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

int ofpsb_tiled_create(ofpsb_ctx* ctx, int rank, int world, int w, int h, int block, int range, int n_slots, ofpsb_tiled** out)
{
    OFPSB_ENTER(ctx);
    if (!out || world < 1 || rank < 0 || rank >= world || w <= 0 || h <= 0 || block <= 0 || (block & 3) || range < 0 ||
        range > 63 || n_slots < 2 || n_slots > 1024 || h / block < world) {
        set_error("tiled_create: invalid arguments (rank %d/%d, %dx%d, block %d, range %d, slots %d)", rank, world, w, h, block,
                  range, n_slots);
        return OFPSB_E_INVALID;
    }
    *out = nullptr;
    ofpsb_tiled* t = new (std::nothrow) ofpsb_tiled();
    if (!t) return OFPSB_E_NOMEM;
    t->ctx = ctx;
    t->rank = rank; t->world = world; t->w = w; t->h = h; t->block = block; t->range = range; t->n_slots = n_slots;
    t->nbx = w / block;
    plan_strip(h, block, range, rank, world, t);
    t->stride = (w + 15) & ~15;
    t->slot_bytes = (((size_t)t->stride * (size_t)(t->own_rows + 2 * range)) + 255) & ~(size_t)255;
    t->flags_off = t->slot_bytes * (size_t)n_slots;
    t->alloc_bytes = t->flags_off + (((size_t)n_slots * 3 * sizeof(uint32_t) + 255) & ~(size_t)255);
    t->epoch.assign((size_t)n_slots, 0u);
    cudaError_t e = cudaMalloc(&t->buf, t->alloc_bytes);
    if (e == cudaSuccess) e = cudaMemset(t->buf, 0, t->alloc_bytes);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&t->side, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t->ev_join, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        set_error("tiled_create: %s", cudaGetErrorString(e));
        ofpsb_tiled_destroy(t);
        return OFPSB_E_CUDA;
    }
    *out = t;
    return OFPSB_OK;
}

void ofpsb_tiled_destroy(ofpsb_tiled* t)
{
    if (!t) return;
    DeviceGuard guard(t->ctx->device);
    cudaStreamSynchronize(t->ctx->stream);
    if (t->side) cudaStreamSynchronize(t->side);
    for (auto& g : t->graphs) cudaGraphExecDestroy(g.second);
    if (t->up.ipc && t->up.base) cudaIpcCloseMemHandle(t->up.base);
    if (t->down.ipc && t->down.base) cudaIpcCloseMemHandle(t->down.base);
    if (t->ev_fork) cudaEventDestroy(t->ev_fork);
    if (t->ev_join) cudaEventDestroy(t->ev_join);
    if (t->side) cudaStreamDestroy(t->side);
    if (t->buf) cudaFree(t->buf);
    delete t;
}

int ofpsb_tiled_info(ofpsb_tiled* t, int* y0, int* rows, int* own_rows, int* nbx, int* nby, int* stride)
{
    if (!t) {
        set_error("tiled_info: null handle");
        return OFPSB_E_INVALID;
    }
    if (y0) *y0 = t->y0;
    if (rows) *rows = t->rows;
    if (own_rows) *own_rows = t->own_rows;
    if (nbx) *nbx = t->nbx;
    if (nby) *nby = t->nby;
    if (stride) *stride = t->stride;
    return OFPSB_OK;
}

int ofpsb_tiled_export(ofpsb_tiled* t, void* handle)
{
    if (!t || !handle) {
        set_error("tiled_export: null argument");
        return OFPSB_E_INVALID;
    }
    OFPSB_ENTER(t->ctx);
    TiledHandle hd;
    memset(&hd, 0, sizeof(hd));
    OFPSB_CUDA_TRY(cudaIpcGetMemHandle(&hd.mem, t->buf));
    hd.magic = TILED_MAGIC;
    hd.rank = t->rank;
    hd.own_rows = t->own_rows;
    hd.stride = t->stride;
    hd.n_slots = t->n_slots;
    hd.range = t->range;
    hd.device = t->ctx->device;
    hd.pid = (int)getpid();
    hd.slot_bytes = t->slot_bytes;
    hd.flags_off = t->flags_off;
    memset(handle, 0, OFPSB_TILED_HANDLE_BYTES);
    memcpy(handle, &hd, sizeof(hd));
    return OFPSB_OK;
}

static int open_neighbour(ofpsb_tiled* t, const void* blob, int expect_rank, Neighbour* nb)
{
    TiledHandle hd;
    memcpy(&hd, blob, sizeof(hd));
    if (hd.magic != TILED_MAGIC || hd.rank != expect_rank || hd.n_slots != t->n_slots || hd.range != t->range) {
        set_error("tiled_connect: handle of rank %d expected (got rank %d, %d slots, range %d)", expect_rank, hd.rank, hd.n_slots,
                  hd.range);
        return OFPSB_E_INVALID;
    }
    if (hd.pid == (int)getpid()) {
        set_error("tiled_connect: rank %d lives in this process — use ofpsb_tiled_connect_local", expect_rank);
        return OFPSB_E_INVALID;
    }
    void* ptr = nullptr;
    OFPSB_CUDA_TRY(cudaIpcOpenMemHandle(&ptr, hd.mem, cudaIpcMemLazyEnablePeerAccess));
    nb->base = static_cast<uint8_t*>(ptr);
    nb->ipc = true;
    nb->own_rows = hd.own_rows;
    nb->stride = hd.stride;
    nb->range = hd.range;
    nb->slot_bytes = (size_t)hd.slot_bytes;
    nb->flags_off = (size_t)hd.flags_off;
    return OFPSB_OK;
}

int ofpsb_tiled_connect(ofpsb_tiled* t, const void* up_handle, const void* down_handle)
{
    if (!t) {
        set_error("tiled_connect: null handle");
        return OFPSB_E_INVALID;
    }
    OFPSB_ENTER(t->ctx);
    if ((t->rank > 0) != (up_handle != nullptr) || (t->rank + 1 < t->world) != (down_handle != nullptr)) {
        set_error("tiled_connect: rank %d of %d needs %s upper and %s lower neighbour", t->rank, t->world,
                  t->rank > 0 ? "an" : "no", t->rank + 1 < t->world ? "a" : "no");
        return OFPSB_E_INVALID;
    }
    if (up_handle)
        if (int rc = open_neighbour(t, up_handle, t->rank - 1, &t->up)) return rc;
    if (down_handle)
        if (int rc = open_neighbour(t, down_handle, t->rank + 1, &t->down)) return rc;
    return OFPSB_OK;
}

int ofpsb_tiled_connect_local(ofpsb_tiled* t, ofpsb_tiled* up, ofpsb_tiled* down)
{
    if (!t || (t->rank > 0) != (up != nullptr) || (t->rank + 1 < t->world) != (down != nullptr)) {
        set_error("tiled_connect_local: wrong neighbours for this rank");
        return OFPSB_E_INVALID;
    }
    OFPSB_ENTER(t->ctx);
    ofpsb_tiled* nbs[2] = {up, down};
    Neighbour* dst[2] = {&t->up, &t->down};
    for (int i = 0; i < 2; i++) {
        ofpsb_tiled* o = nbs[i];
        if (!o) continue;
        if (o->rank != t->rank + (i ? 1 : -1) || o->n_slots != t->n_slots || o->range != t->range || o->w != t->w) {
            set_error("tiled_connect_local: neighbour geometry differs");
            return OFPSB_E_INVALID;
        }
        if (o->ctx->device != t->ctx->device) {
            cudaError_t e = cudaDeviceEnablePeerAccess(o->ctx->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                set_error("tiled_connect_local: no peer access from device %d to %d (%s)", t->ctx->device, o->ctx->device,
                          cudaGetErrorString(e));
                return OFPSB_E_CUDA;
            }
            cudaGetLastError();
        }
        dst[i]->base = o->buf;
        dst[i]->ipc = false;
        dst[i]->own_rows = o->own_rows;
        dst[i]->stride = o->stride;
        dst[i]->range = o->range;
        dst[i]->slot_bytes = o->slot_bytes;
        dst[i]->flags_off = o->flags_off;
    }
    return OFPSB_OK;
}

void* ofpsb_tiled_slot_ptr(ofpsb_tiled* t, int slot)
{
    if (!t || slot < 0 || slot >= t->n_slots) return nullptr;
    return t->own(slot);
}

int ofpsb_tiled_upload(ofpsb_tiled* t, int slot, const uint8_t* host_rows, size_t host_stride)
{
    if (!t || !host_rows || slot < 0 || slot >= t->n_slots || host_stride < (size_t)t->w) {
        set_error("tiled_upload: invalid arguments");
        return OFPSB_E_INVALID;
    }
    OFPSB_ENTER(t->ctx);
    OFPSB_CUDA_TRY(cudaMemcpy2DAsync(t->own(slot), t->stride, host_rows, host_stride, t->w, t->own_rows, cudaMemcpyHostToDevice,
                                     t->ctx->stream));
    return OFPSB_OK;
}

int ofpsb_tiled_publish(ofpsb_tiled* t, int slot)
{
    if (!t || slot < 0 || slot >= t->n_slots) {
        set_error("tiled_publish: invalid arguments");
        return OFPSB_E_INVALID;
    }
    OFPSB_ENTER(t->ctx);
    const uint32_t e = ++t->epoch[(size_t)slot];
    // I am the upper neighbour's DOWN side and the lower neighbour's UP side
    uint32_t* a = t->up.base ? t->up.flags() + 2 * slot + 1 : nullptr;
    uint32_t* b = t->down.base ? t->down.flags() + 2 * slot + 0 : nullptr;
    {
        tiled_publish_kernel<<<1, 1, 0, t->ctx->stream>>>(a, b, t->expected() + slot, e);
        OFPSB_CUDA_TRY(cudaGetLastError());
        t->ctx->launches++;
    }
    return OFPSB_OK;
}

static int tiled_match_impl(ofpsb_tiled* t, int prev_slot, int cur_slot, int n_pairs, ofps_mv* d_entries, int16_t* d_mv_xy,
                            uint32_t* d_cost, int wait_neighbours)
{
    OFPSB_ENTER(t->ctx);
    ofpsb_ctx* ctx = t->ctx;
    const GraphKey key{prev_slot, cur_slot, n_pairs, d_entries, d_mv_xy, d_cost, wait_neighbours != 0};
    if (t->use_graph && !ctx->bm_scratch.collect_stats && !ctx->bm_scratch.profile) {
        for (auto& g : t->graphs)
            if (g.first == key) {
                OFPSB_CUDA_TRY(cudaGraphLaunch(g.second, ctx->stream));
                ctx->launches += 2;
                return OFPSB_OK;
            }
        bool second = false;
        for (auto& k : t->seen) second |= k == key;
        if (second) {   // the first call ran eagerly (scratch sized, attributes set): capture this one
            cudaGraph_t graph = nullptr;
            OFPSB_CUDA_TRY(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
            const uint64_t l0 = ctx->launches;
            const int rc = enqueue_match(t, prev_slot, cur_slot, n_pairs, d_entries, d_mv_xy, d_cost, wait_neighbours != 0);
            const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
            ctx->launches = l0;
            if (rc == OFPSB_OK && ce == cudaSuccess && graph) {
                cudaGraphExec_t exec = nullptr;
                if (cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
                    cudaGraphDestroy(graph);
                    if (t->graphs.size() >= 64) {
                        cudaGraphExecDestroy(t->graphs.front().second);
                        t->graphs.erase(t->graphs.begin());
                    }
                    t->graphs.emplace_back(key, exec);
                    OFPSB_CUDA_TRY(cudaGraphLaunch(exec, ctx->stream));
                    ctx->launches += 2;
                    return OFPSB_OK;
                }
            }
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            t->use_graph = false;   // capture not possible here: stay on the eager path
        } else {
            t->seen.push_back(key);
        }
    }
    return enqueue_match(t, prev_slot, cur_slot, n_pairs, d_entries, d_mv_xy, d_cost, wait_neighbours != 0);
}

int ofpsb_tiled_match(ofpsb_tiled* t, int prev_slot, int cur_slot, ofps_mv* d_entries, int16_t* d_mv_xy, uint32_t* d_cost,
                      int wait_neighbours)
{
    if (!t || prev_slot < 0 || prev_slot >= t->n_slots || cur_slot < 0 || cur_slot >= t->n_slots) {
        set_error("tiled_match: invalid arguments");
        return OFPSB_E_INVALID;
    }
    return tiled_match_impl(t, prev_slot, cur_slot, 1, d_entries, d_mv_xy, d_cost, wait_neighbours);
}

int ofpsb_tiled_match_stream(ofpsb_tiled* t, int first_slot, int n_pairs, ofps_mv* d_entries, int16_t* d_mv_xy, uint32_t* d_cost,
                             int wait_neighbours)
{
    if (!t || n_pairs < 1 || first_slot < 0 || first_slot + n_pairs >= t->n_slots) {
        set_error("tiled_match_stream: slots %d .. %d out of range (%d slots)", first_slot, first_slot + n_pairs,
                  t ? t->n_slots : 0);
        return OFPSB_E_INVALID;
    }
    return tiled_match_impl(t, first_slot, first_slot + 1, n_pairs, d_entries, d_mv_xy, d_cost, wait_neighbours);
}

}  // extern "C"
#pragma GCC visibility pop
