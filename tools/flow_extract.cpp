// C++ counterpart of the reference's `flow-extract` binary (flow-extract/src/main.rs:50-128) on top of the C ABI:
// every frame of a .mvec file (motion-loader/src/lib.rs:46-65) -> dense WIDTH x HEIGHT field (densifier on the GPU,
// hole fill on the host: ofpsb_flow_field) -> Middlebury .flo files `<outdir>/NNNNNN.flo` (flow-extract/src/main.rs:122).
// Like the reference, a frame without vectors repeats the previous field (:69-71).
//
//   flow_extract <input.mvec> <outdir> <width> <height>
//
// Build: g++ -std=c++17 -O2 -Iinclude tools/flow_extract.cpp -o flow_extract -Lofps_b200 -lofps_b200 -Wl,-rpath,$PWD/ofps_b200
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include <sys/stat.h>

#include "ofps_b200.hpp"

using namespace ofps_b200;

int main(int argc, char** argv)
{
    if (argc < 5) {
        std::fprintf(stderr, "usage: %s <input.mvec> <outdir> <width> <height>\n", argv[0]);
        return 2;
    }
    const size_t w = (size_t)std::atoll(argv[3]), h = (size_t)std::atoll(argv[4]);
    if (w == 0 || h == 0) {
        std::fprintf(stderr, "flow_extract: bad field size\n");
        return 2;
    }
    try {
        Context ctx(0);
        mkdir(argv[2], 0777);   // create_dir_all (:52); an existing directory is fine
        std::vector<float> field(2 * w * h, 0.0f);
        MotionVectors mv;
        size_t frames = 0;
        for (;; frames++) {
            size_t n = 0;
            if (ofpsb_mvec_read(argv[1], frames, nullptr, 0, &n) != OFPSB_OK) break;   // past the last frame
            if (n) {
                mv.resize(n);
                check(ofpsb_mvec_read(argv[1], frames, mv.data(), n, &n));
                check(ofpsb_flow_field(ctx.get(), mv.data(), n, w, h, field.data()));
            }
            char name[32];
            std::snprintf(name, sizeof name, "/%06zu.flo", frames);
            check(ofpsb_flo_write((std::string(argv[2]) + name).c_str(), field.data(), w, h));
        }
        std::printf("%zu frames -> %s (%zux%zu)\n", frames, argv[2], w, h);
        return frames ? 0 : 1;
    } catch (const Error& e) {
        std::fprintf(stderr, "flow_extract: %s (code %d)\n", e.what(), e.code);
        return e.code == OFPSB_E_NODEVICE ? 3 : 1;
    }
}
