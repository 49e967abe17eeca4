// TEST INFRASTRUCTURE — runs the fused SEA block-matching kernel of ofps_b200/csrc/block_match_sea.cu on the CPU
// stand-in of tests/emu/cuda_emu.h (the two TMA box loads are replaced by plain copies with zero fill, everything
// after them is the product code, unmodified), so that its logic — window sums, bounds, tie-breaks, predictor and
// work-list decisions — can be checked against the oracle on the GPU-less build container.  Not shipped.
#define OFPSB_EMU 1
#include "../../ofps_b200/csrc/block_match_sea.cu"

namespace ofpsb {
void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    fputc('\n', stderr);
    va_end(ap);
}
}  // namespace ofpsb

using namespace ofpsb;

extern "C" {
// One strip (or whole frame: halo_top = halo_bottom = y_offset = 0, full_h = strip_h) of n_pairs pairs.
// worklist: capacity nbx*nby*n_pairs; *wl_count receives the number of blocks the kernel left to the exhaustive search
// (their outputs are untouched).  stats: 4 x u64.  tile_h: 64 or 32 rows per tile.
int emu_block_match_sea(const uint8_t* prev, const uint8_t* cur, int w, int strip_h, int stride, long long pair_stride,
                        int n_pairs, int halo_top, int halo_bottom, int y_offset, int full_h, int block, int range,
                        int16_t* mv, uint32_t* cost, ofps_mv* entries, uint32_t* worklist, uint32_t* wl_count,
                        unsigned long long* stats, int tile_h)
{
    BlockMatchParams p{};
    p.prev = prev; p.cur = cur; p.w = w; p.strip_h = strip_h; p.stride = stride; p.pair_stride = pair_stride;
    p.n_pairs = n_pairs; p.halo_top = halo_top; p.halo_bottom = halo_bottom; p.y_offset = y_offset; p.full_h = full_h;
    p.block = block; p.range = range; p.metric = OFPSB_METRIC_SAD; p.nbx = w / block; p.nby = strip_h / block;
    p.mv_xy = mv; p.cost = cost; p.entries = entries;
    SeaOut out;
    out.worklist = worklist; out.wl_count = wl_count; out.stats = stats;
    out.nx = 1.0f / (float)w; out.ny = 1.0f / (float)full_h; out.prefetch_tiles = 0; out.debug_stop = 0;
    *wl_count = 0;
    const int th = (tile_h == 32 || range > 16) ? 32 : 64;
    const dim3 grid((p.nbx * block + SEA_TILE_W - 1) / SEA_TILE_W, (p.nby * block + th - 1) / th, n_pairs);
    SeaMaps maps{};
#define EMU_SEA(BB, RR)                                                                                   \
    do {                                                                                                  \
        if (th == 32) OFPSB_LAUNCH_SMEM((sea_kernel<BB, RR, 32>), grid, SEA_NT, 0, nullptr, maps, p, out); \
        else OFPSB_LAUNCH_SMEM((sea_kernel<BB, RR, 64>), grid, SEA_NT, 0, nullptr, maps, p, out);          \
    } while (0)
    if (block == 16 && range == 16) EMU_SEA(16, 16);
    else if (block == 16 && range == 8) EMU_SEA(16, 8);
    else if (block == 8 && range == 16) EMU_SEA(8, 16);
    else if (block == 8 && range == 8) EMU_SEA(8, 8);
    else if (block == 8 && range == 32) OFPSB_LAUNCH_SMEM((sea_kernel<8, 32, 32>), grid, SEA_NT, 0, nullptr, maps, p, out);
    else if (block == 16 && range == 32) OFPSB_LAUNCH_SMEM((sea_kernel<16, 32, 32>), grid, SEA_NT, 0, nullptr, maps, p, out);
#undef EMU_SEA
    else return 1;
    return 0;
}
}
