// Compile/link check of the C++ host mirror (include/ofps_b200.hpp) against libofps_b200.so, and —
// with a GPU — a run of Decoder -> Detector -> Estimator through it.  Built and run by
// tests/test_capi_cpu.py (compile, no-GPU behaviour) and tests/test_gpu_cpp_host.py (run).
#include <cstdio>
#include <cstring>

#include "ofps_b200.hpp"

using namespace ofps_b200;

int main(int argc, char** argv)
{
    std::shared_ptr<Context> ctx;
    try {
        ctx = std::make_shared<Context>(0);
    } catch (const Error& e) {
        std::printf("NO_DEVICE code=%d %s\n", e.code, e.what());
        return e.code == OFPSB_E_NODEVICE ? 3 : 1;
    }
    const int w = 640, h = 360;
    int frame_no = 0;
    auto source = [&](uint8_t* luma) {
        if (frame_no >= 3) return false;
        // textured frame; a 160x96 rectangle moves +5,+3 px per frame, the background stays still
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                int sx = x, sy = y;
                const int ox = 200 + 5 * frame_no, oy = 100 + 3 * frame_no;
                if (x >= ox && x < ox + 160 && y >= oy && y < oy + 96) { sx = x - 5 * frame_no + 1000; sy = y - 3 * frame_no; }
                unsigned v = (unsigned)(sx * 2654435761u) ^ (unsigned)(sy * 40503u * 65537u);
                v ^= v >> 13; v *= 0x5bd1e995u; v ^= v >> 15;
                luma[(size_t)y * w + x] = (uint8_t)(v >> 24);
            }
        frame_no++;
        return true;
    };
    BlockMatchDecoder dec(ctx, w, h, 30.0, source);
    dec.range = 8;
    BlockMotionDetection det(ctx);
    AlmeidaEstimator est(ctx);
    est.use_ransac = argc > 1 && !std::strcmp(argv[1], "ransac");
    StandardCamera cam(16.0f / 9.0f, 22.275f);
    if (det.props_mut().size() != 3 || est.props_mut().size() != 4 || std::strcmp(det.props_mut()[0].first, "Min size")) return 1;
    MotionVectors mv;
    std::vector<RGBA> rgba;
    size_t height = 0;
    int frames_with_mv = 0, detections = 0;
    try {
        for (;;) {
            mv.clear();                                   // callers clear, decoders append (detection.rs:97)
            if (!dec.process_frame(mv, &rgba, &height, 0)) continue;
            frames_with_mv++;
            auto d = det.detect_motion(mv);
            if (d) detections++;
            const Pose p = est.estimate(mv.data(), mv.size(), cam);
            std::printf("frame %d: %zu vectors, detector %s area %zu, q = (%.6f %.6f %.6f %.6f)\n", frame_no, mv.size(),
                        d ? "Some" : "None", d ? d->first : 0, p.rotation[0], p.rotation[1], p.rotation[2], p.rotation[3]);
        }
    } catch (const Error& e) {
        if (std::strcmp(e.what(), "end of stream")) { std::printf("error: %s\n", e.what()); return 1; }
    }
    std::printf("OK frames=%d detections=%d rgba=%zu height=%zu\n", frames_with_mv, detections, rgba.size(), height);
    if (!(frames_with_mv == 2 && detections == 2 && rgba.size() == (size_t)w * h && height == (size_t)h)) return 1;

    // ---- DenseFlowDecoder (the cv-decoder mirror): BGR frames + a host-provided flow -> MotionEntry list
    for (int fullres = 1; fullres >= 0; fullres--) {
        int fno = 0;
        auto frames = [&](std::vector<uint8_t>& bgr, int& fw, int& fh) {
            if (fno >= 3) return false;
            fw = w; fh = h;
            bgr.assign((size_t)w * h * 3, 90);
            for (int y = 100; y < 200; y++)   // a bright rectangle: its corners raise the contrast mask
                for (int x = 150 + 4 * fno; x < 330 + 4 * fno; x++)
                    for (int c = 0; c < 3; c++) bgr[((size_t)y * w + x) * 3 + c] = (uint8_t)(200 + 10 * c);
            fno++;
            return true;
        };
        auto flow = [&](const uint8_t*, const uint8_t*, const uint8_t*, const uint8_t*, int fw, int fh, bool, bool, float* f) {
            for (size_t i = 0; i < (size_t)fw * fh; i++) { f[2 * i] = 4.0f; f[2 * i + 1] = 0.0f; }
        };
        DenseFlowDecoder dd(ctx, frames, flow, 25.0);
        dd.process_fullres = fullres != 0;
        if (dd.props_mut().size() != 4 || std::strcmp(dd.props_mut()[3].first, "Process Fullres")) return 1;
        int got = 0;
        size_t last_n = 0;
        try {
            for (;;) {
                mv.clear();
                if (!dd.process_frame(mv, &rgba, &height, 0)) continue;
                got++;
                last_n = mv.size();
                for (const auto& e : mv)
                    if (!(e.px > 0 && e.px < 1 && e.py > 0 && e.py < 1 && e.my == 0.0f && e.mx > 0.0f)) return 1;
            }
        } catch (const Error& e) {
            if (std::strcmp(e.what(), "Failed to grab frame")) { std::printf("error: %s\n", e.what()); return 1; }
        }
        const auto asp = dd.get_aspect();
        std::printf("DENSE fullres=%d frames=%d entries=%zu aspect=%zux%zu rgba=%zu\n", fullres, got, last_n, asp->first,
                    asp->second, rgba.size());
        const size_t ew = fullres ? (size_t)w : 150, eh = fullres ? (size_t)h : 84;
        if (got != 2 || last_n == 0 || last_n >= 150 * 84 || asp->first != ew || asp->second != eh || rgba.size() != ew * eh ||
            height != eh)
            return 1;
    }
    std::printf("DENSE OK\n");
    return 0;
}
