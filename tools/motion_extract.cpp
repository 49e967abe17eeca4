// C++ counterpart of the reference's `motion-extract` binary (motion-extract/src/main.rs) for the B200 block-matching
// decoder: reads luma frames, runs the decoder and appends one .mvec frame per decoded frame
// (u32 LE count + count x 4 f32 LE, motion-extract/src/main.rs:23-35).
//
//   motion_extract <input> <output.mvec> [block] [range] [sad|ssd]
//     input: "WIDTHxHEIGHT@FPS:frames.y" (raw luma) or "clip.y4m"
//
// Build: g++ -std=c++17 -O2 -Iinclude tools/motion_extract.cpp -o motion_extract -Lofps_b200 -lofps_b200 -Wl,-rpath,$PWD/ofps_b200
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ofps_b200.hpp"

using namespace ofps_b200;

int main(int argc, char** argv)
{
    if (argc < 3) {
        std::fprintf(stderr, "usage: %s <WxH@FPS:frames.y | clip.y4m> <out.mvec> [block=16] [range=16] [sad|ssd]\n", argv[0]);
        return 2;
    }
    try {
        LumaFileSource src(argv[1]);
        auto ctx = std::make_shared<Context>(0);
        BlockMatchDecoder dec(ctx, src.width(), src.height(), src.framerate(), [&](uint8_t* luma) { return src.next(luma); });
        if (argc > 3) dec.block = std::atoi(argv[3]);
        if (argc > 4) dec.range = std::atoi(argv[4]);
        if (argc > 5 && !std::strcmp(argv[5], "ssd")) dec.metric = OFPSB_METRIC_SSD;
        MotionVectors mv;
        size_t frames = 0, vectors = 0;
        bool first = true;
        for (;;) {
            mv.clear();   // callers clear, decoders append (motion-extract/src/main.rs:34)
            bool filled;
            try {
                filled = dec.process_frame(mv, nullptr, nullptr, 0);
            } catch (const Error& e) {
                if (!std::strcmp(e.what(), "end of stream")) break;
                throw;
            }
            if (!filled) continue;   // first frame: no predecessor (an I-frame in av-decoder's terms)
            check(ofpsb_mvec_append(argv[2], mv.data(), mv.size(), first ? 1 : 0));
            first = false;
            frames++;
            vectors += mv.size();
        }
        std::printf("%zu frames, %zu vectors -> %s (%dx%d, block %d, range %d)\n", frames, vectors, argv[2], src.width(), src.height(),
                    dec.block, dec.range);
        return frames ? 0 : 1;
    } catch (const Error& e) {
        std::fprintf(stderr, "motion_extract: %s (code %d)\n", e.what(), e.code);
        return e.code == OFPSB_E_NODEVICE ? 3 : 1;
    }
}
