// ofps_b200.hpp — C++17 host-side mirror of the reference's plugin interface for the hot path,
// layered over the C ABI (ofps_b200.h).  Header-only; link with -lofps_b200.
//
// The reference is Rust; its plugin traits are not FFI-safe, so a drop-in needs a
// thin Rust cdylib per plugin (sources under rust/, binding shown in INTEGRATION.md).  This header
// is the same surface for C++ hosts, with the reference's names, argument meaning, property names /
// bounds and error behaviour:
//
//   ofps::Decoder::process_frame / get_framerate / get_aspect      ofps/src/decoder.rs:45-73
//   ofps::Detector::detect_motion                                   ofps/src/detection.rs:6-12
//   ofps::Estimator::estimate / motion_step                         ofps/src/estimator.rs:8-54
//   ofps::Properties::props_mut                                     ofps/src/plugins/properties.rs:6-18
//   ofps::MotionField, MotionFieldDensifier (read side)             ofps/src/motion_field.rs:7-115
//   ofps::StandardCamera::new / aspect_ratio / fov                  ofps/src/camera.rs:26-35, 166-177
#ifndef OFPS_B200_HPP
#define OFPS_B200_HPP

#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <variant>
#include <vector>

#include "ofps_b200.h"

namespace ofps_b200 {

// anyhow::Error of the reference's Decoder / Estimator results
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

inline void check(int rc)
{
    if (rc != OFPSB_OK) throw Error(rc, ofpsb_last_error());
}

using MotionEntry = ofps_mv;                       // (Point2 pos, Vector2 motion), ofps/src/decoder.rs:40
using MotionVectors = std::vector<MotionEntry>;    // ofps/src/decoder.rs:42
struct RGBA { uint8_t r, g, b, a; };               // ofps/src/decoder.rs:12-26

// One context (device, stream, scratch) shared by the plugins of a host thread.  Send, not Sync.
class Context {
public:
    explicit Context(int device = 0) { check(ofpsb_create(device, &ctx_)); }
    ~Context() { ofpsb_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    ofpsb_ctx* get() const { return ctx_; }

private:
    ofpsb_ctx* ctx_ = nullptr;
};

// ---- Properties (ofps/src/plugins/properties.rs:6-18, 120-125)
struct FloatProp { float* val; float min, max; };
struct UsizeProp { size_t* val; size_t min, max; };
struct BoolProp { bool* val; };
using PropertyMut = std::variant<FloatProp, UsizeProp, BoolProp>;
using PropList = std::vector<std::pair<const char*, PropertyMut>>;

// ---- MotionField (ofps/src/motion_field.rs:7-115): cell-major [x0,y0,x1,y1,...], cell = y*w + x
class MotionField {
public:
    MotionField(size_t width, size_t height) : vf_(2 * width * height, 0.0f), width_(width) {}
    std::pair<size_t, size_t> dim() const { return {width_, width_ ? vf_.size() / 2 / width_ : 0}; }
    size_t size() const { return vf_.size() / 2; }
    const float* as_slice() const { return vf_.data(); }
    float* as_mut_slice() { return vf_.data(); }
    std::array<float, 2> get_motion(size_t x, size_t y) const
    {
        const size_t i = y * width_ + x;
        return {vf_[2 * i], vf_[2 * i + 1]};
    }
    void set_motion(size_t x, size_t y, std::array<float, 2> m)
    {
        const size_t i = y * width_ + x;
        vf_[2 * i] = m[0];
        vf_[2 * i + 1] = m[1];
    }
    // motion_iter(): positions are cell corners (x/w, y/h), motion_field.rs:104-112
    MotionVectors motion_iter() const
    {
        MotionVectors out;
        const auto [w, h] = dim();
        for (size_t y = 0; y < h; y++)
            for (size_t x = 0; x < w; x++) {
                const auto m = get_motion(x, y);
                out.push_back({(float)x / (float)w, (float)y / (float)h, m[0], m[1]});
            }
        return out;
    }

private:
    std::vector<float> vf_;
    size_t width_;
};

// MotionFieldDensifier::add_vector over a slice + MotionField::from (motion_field.rs:133-190, 297-308)
inline MotionField densify(Context& ctx, const MotionEntry* entries, size_t n, size_t width, size_t height)
{
    MotionField mf(width, height);
    check(ofpsb_densify(ctx.get(), entries, n, width, height, mf.as_mut_slice(), nullptr));
    return mf;
}

// ---- StandardCamera (ofps/src/camera.rs:9-35): only what the estimator boundary needs
class StandardCamera {
public:
    StandardCamera(float aspect_ratio, float fov_y_deg) : aspect_(aspect_ratio), fov_y_(fov_y_deg) {}
    float aspect_ratio() const { return aspect_; }
    // (horizontal, vertical) field of view in degrees (camera.rs:166-177)
    std::pair<float, float> fov() const
    {
        const float half = std::tan(fov_y_ * 0.017453292519943295f / 2.0f);
        return {std::atan(half * aspect_) * 2.0f / 0.017453292519943295f, fov_y_};
    }

private:
    float aspect_, fov_y_;
};

// ---- Detector: block-motion-detector/src/lib.rs:13-119
class BlockMotionDetection {
public:
    float min_size = 0.05f;
    size_t subdivide = 3;
    float target_motion = 0.003f;

    explicit BlockMotionDetection(std::shared_ptr<Context> ctx) : ctx_(std::move(ctx)) {}

    PropList props_mut()
    {
        return {{"Min size", FloatProp{&min_size, 0.01f, 1.0f}},
                {"Subdivisions", UsizeProp{&subdivide, 1, 16}},
                {"Target motion", FloatProp{&target_motion, 0.0001f, 0.1f}}};
    }

    // None = no motion.  The detector has no error channel in the reference; a device failure throws.
    std::optional<std::pair<size_t, MotionField>> detect_motion(const MotionEntry* motion, size_t n) const
    {
        size_t dim = 0;
        check(ofpsb_block_dim(min_size, subdivide, &dim));
        MotionField mf(dim, dim);
        int has = 0;
        size_t area = 0;
        check(ofpsb_detect_block_motion(ctx_->get(), motion, n, min_size, subdivide, target_motion, &has, &area, &dim,
                                        mf.as_mut_slice(), mf.size()));
        if (!has) return std::nullopt;
        return std::make_pair(area, std::move(mf));
    }
    std::optional<std::pair<size_t, MotionField>> detect_motion(const MotionVectors& motion) const
    {
        return detect_motion(motion.data(), motion.size());
    }

private:
    std::shared_ptr<Context> ctx_;
};

// ---- Estimator: almeida-estimator/src/lib.rs:57-121
struct Pose {
    std::array<float, 4> rotation;      // UnitQuaternion (w, i, j, k)
    std::array<float, 3> translation;   // always zero (almeida:120)
};

class AlmeidaEstimator {
public:
    bool use_ransac = true;
    size_t num_iters = 200;
    float inlier_angle = 0.05f;
    size_t ransac_samples = 1000;
    uint64_t seed = 0;   // the reference draws from thread_rng(); here the draw is seeded and advances per call

    explicit AlmeidaEstimator(std::shared_ptr<Context> ctx) : ctx_(std::move(ctx)) {}

    PropList props_mut()
    {
        return {{"Use ransac", BoolProp{&use_ransac}},
                {"Ransac iters", UsizeProp{&num_iters, 1, 500}},
                {"Inlier threshold", FloatProp{&inlier_angle, 0.01f, 1.0f}},
                {"Ransac samples", UsizeProp{&ransac_samples, 100, 16000}}};
    }

    // `move_magnitude` is ignored, as in the reference (almeida:105)
    Pose estimate(const MotionEntry* motion, size_t n, const StandardCamera& camera, std::optional<float> = std::nullopt)
    {
        Pose p{{1, 0, 0, 0}, {0, 0, 0}};
        check(ofpsb_almeida(ctx_->get(), motion, n, camera.aspect_ratio(), camera.fov().second, use_ransac ? 1 : 0,
                            num_iters, inlier_angle, ransac_samples, seed++, p.rotation.data()));
        return p;
    }

private:
    std::shared_ptr<Context> ctx_;
};

// ---- Decoder: a luma frame source + the block matcher, emitting what av-decoder emits
// (av-decoder/src/lib.rs:396-419): one MotionEntry per block, appended to the caller's vector.
class BlockMatchDecoder {
public:
    int block = 16, range = 16, metric = OFPSB_METRIC_SAD;
    bool emit_zero_motion = true;   // FFmpeg exports nothing for skipped blocks; default keeps every block

    // `next_frame(luma)` fills a width*height u8 luma plane and returns false at end of stream.
    template <typename F>
    BlockMatchDecoder(std::shared_ptr<Context> ctx, int width, int height, double framerate, F next_frame)
        : ctx_(std::move(ctx)), w_(width), h_(height), fps_(framerate), next_(std::move(next_frame)),
          cur_((size_t)width * height)
    {
    }
    BlockMatchDecoder(const BlockMatchDecoder&) = delete;
    BlockMatchDecoder& operator=(const BlockMatchDecoder&) = delete;
    ~BlockMatchDecoder()
    {
        if (stream_) ofpsb_stream_close(stream_);
    }

    PropList props_mut() { return {}; }
    std::optional<double> get_framerate() const { return fps_ > 0 ? std::optional<double>(fps_) : std::nullopt; }
    std::optional<std::pair<size_t, size_t>> get_aspect() const { return std::make_pair((size_t)w_, (size_t)h_); }

    // Ok(true): vectors appended; Ok(false): frame had none (first frame); throws Error at end of stream /
    // on failure.  `out_frame` (RGBA, cleared then filled) and `skip_frames` as in decoder.rs:54-59.
    // Every decoded frame goes through the streaming entry points (ofpsb_stream_*): it is uploaded once and the
    // previous frame stays in HBM; a skipped frame is pushed too (it is the next pair's previous frame), its
    // vectors are dropped.
    bool process_frame(MotionVectors& field, std::vector<RGBA>* out_frame, size_t* out_height, size_t skip_frames)
    {
        if (!stream_ || s_block_ != block || s_range_ != range || s_metric_ != metric) {
            if (stream_) ofpsb_stream_close(stream_);
            stream_ = nullptr;
            check(ofpsb_stream_open(ctx_->get(), w_, h_, block, range, metric, 4, &stream_));
            s_block_ = block; s_range_ = range; s_metric_ = metric;
            scratch_.resize(ofpsb_stream_blocks(stream_));
        }
        size_t nb = 0;
        for (size_t i = 0; i <= skip_frames; i++) {
            if (!next_(cur_.data())) throw Error(OFPSB_E_IO, "end of stream");
            check(ofpsb_stream_push(stream_, cur_.data(), (size_t)w_, scratch_.data(), &nb));
        }
        if (out_frame) {
            out_frame->clear();
            for (uint8_t v : cur_) out_frame->push_back({v, v, v, 255});
            if (out_height) *out_height = (size_t)h_;
        }
        size_t pushed = 0;
        for (size_t i = 0; i < nb; i++)
            if (emit_zero_motion || scratch_[i].mx != 0.0f || scratch_[i].my != 0.0f) {
                field.push_back(scratch_[i]);
                pushed++;
            }
        return pushed > 0;
    }

private:
    std::shared_ptr<Context> ctx_;
    int w_, h_;
    double fps_;
    std::function<bool(uint8_t*)> next_;
    std::vector<uint8_t> cur_;
    MotionVectors scratch_;
    ofpsb_stream* stream_ = nullptr;
    int s_block_ = 0, s_range_ = 0, s_metric_ = 0;
};

// ---- Frame source for BlockMatchDecoder: 8-bit luma from a raw file or a YUV4MPEG2 (.y4m) stream.
//   "WIDTHxHEIGHT@FPS:path"  raw luma planes back to back (the input string of the Rust `b200_block` shim,
//                            rust/b200-block-decoder/src/lib.rs — the reference passes the decoder's input
//                            string through unchanged, ofps/src/plugins/mod.rs:146-159)
//   "path.y4m"               YUV4MPEG2: W / H / F from the header, Y plane of every FRAME (chroma skipped;
//                            C420* / C422 / C444 / Cmono, 8 bit)
class LumaFileSource {
public:
    explicit LumaFileSource(const std::string& input)
    {
        const auto colon = input.find(':');
        const auto x = input.find('x');
        // a geometry prefix is digits, 'x', '@' and '.' only, up to the first ':'
        if (colon != std::string::npos && x != std::string::npos && x < colon &&
            input.find_first_not_of("0123456789x@.", 0) == colon) {
            const auto at = input.find('@');
            w_ = std::atoi(input.substr(0, x).c_str());
            h_ = std::atoi(input.substr(x + 1, (at != std::string::npos && at < colon ? at : colon) - x - 1).c_str());
            if (at != std::string::npos && at < colon) fps_ = std::atof(input.substr(at + 1, colon - at - 1).c_str());
            open(input.substr(colon + 1));
            chroma_ = 0;
        } else {
            open(input);
            char line[256];
            if (!std::fgets(line, sizeof line, f_) || std::strncmp(line, "YUV4MPEG2", 9) != 0) fail("not a YUV4MPEG2 stream: " + input);
            int cw = 2, ch = 2;   // chroma subsampling divisors; C420 is the format's default
            for (char* tok = std::strtok(line + 9, " \n"); tok; tok = std::strtok(nullptr, " \n")) {
                if (tok[0] == 'W') w_ = std::atoi(tok + 1);
                else if (tok[0] == 'H') h_ = std::atoi(tok + 1);
                else if (tok[0] == 'F') {
                    int n = 0, d = 1;
                    if (std::sscanf(tok + 1, "%d:%d", &n, &d) == 2 && d > 0) fps_ = (double)n / d;
                } else if (tok[0] == 'C') {
                    if (!std::strncmp(tok, "C420", 4)) { cw = 2; ch = 2; }
                    else if (!std::strcmp(tok, "C422")) { cw = 2; ch = 1; }
                    else if (!std::strcmp(tok, "C444")) { cw = 1; ch = 1; }
                    else if (!std::strcmp(tok, "Cmono")) { cw = 0; ch = 0; }
                    else fail(std::string("unsupported y4m colour space ") + tok);
                }
            }
            if (w_ > 0 && h_ > 0) chroma_ = cw ? 2 * (size_t)((w_ + cw - 1) / cw) * (size_t)((h_ + ch - 1) / ch) : 0;
            y4m_ = true;
        }
        if (w_ <= 0 || h_ <= 0) fail("bad frame size in " + input);
    }
    ~LumaFileSource() { if (f_) std::fclose(f_); }
    LumaFileSource(const LumaFileSource&) = delete;
    LumaFileSource& operator=(const LumaFileSource&) = delete;
    int width() const { return w_; }
    int height() const { return h_; }
    double framerate() const { return fps_; }
    // Fills width*height bytes; false at end of stream.  Usable directly as BlockMatchDecoder's frame source.
    bool next(uint8_t* luma)
    {
        if (y4m_) {
            char line[128];
            if (!std::fgets(line, sizeof line, f_) || std::strncmp(line, "FRAME", 5) != 0) return false;
        }
        const size_t n = (size_t)w_ * h_;
        if (std::fread(luma, 1, n, f_) != n) return false;
        if (chroma_ && std::fseek(f_, (long)chroma_, SEEK_CUR) != 0) return false;
        return true;
    }

private:
    void open(const std::string& path)
    {
        f_ = std::fopen(path.c_str(), "rb");
        if (!f_) fail("cannot open " + path);
    }
    [[noreturn]] void fail(const std::string& what) const { throw Error(OFPSB_E_IO, what); }
    std::FILE* f_ = nullptr;
    int w_ = 0, h_ = 0;
    double fps_ = 0.0;
    size_t chroma_ = 0;
    bool y4m_ = false;
};

// ---- Decoder: the reference's cv-decoder (cv-decoder/src/lib.rs:17-307) with its two third-party calls
// abstracted — the frame source (cv::VideoCapture::read) and the optical flow (cv::calcOpticalFlowFarneback /
// cv::optflow::calcOpticalFlowDenseRLOF) are callbacks of the host; everything the reference does AROUND them
// runs on the GPU: resize (:127-135), BGR2GRAY (:138), RGBA out_frame (:145-153), the Sobel / threshold /
// dilate contrast mask (:204-236) and the flow -> MotionEntry conversion through the down-sampling
// MotionFieldDensifier (:238-291).  Same property names and bounds as the reference (:34-52).
class DenseFlowDecoder {
public:
    std::pair<size_t, size_t> max_mfield_size{150, 150};   // "Width", "Height"
    bool use_rlof = false;                                 // "RLOF": no contrast mask
    bool process_fullres = true;                           // "Process Fullres"

    // Fills `bgr` (h rows of w interleaved B,G,R bytes) and the size; false at end of stream.
    using FrameSource = std::function<bool(std::vector<uint8_t>& bgr, int& w, int& h)>;
    // Dense flow from the previous to the current frame into flow_xy (w*h*2 f32, pixels).  `initial` is true
    // when flow_xy still holds the previous result (OPTFLOW_USE_INITIAL_FLOW, cv-decoder:160-164).
    using FlowSource = std::function<void(const uint8_t* old_gray, const uint8_t* gray, const uint8_t* old_bgr,
                                          const uint8_t* bgr, int w, int h, bool rlof, bool initial, float* flow_xy)>;

    DenseFlowDecoder(std::shared_ptr<Context> ctx, FrameSource frames, FlowSource flow, double framerate = 0.0,
                     std::pair<size_t, size_t> aspect_ratio_scale = {1, 1})
        : ctx_(std::move(ctx)), frames_(std::move(frames)), flow_fn_(std::move(flow)), fps_(framerate), ar_(aspect_ratio_scale)
    {
    }

    PropList props_mut()
    {
        return {{"Width", UsizeProp{&max_mfield_size.first, 1, 2000}},
                {"Height", UsizeProp{&max_mfield_size.second, 1, 2000}},
                {"RLOF", BoolProp{&use_rlof}},
                {"Process Fullres", BoolProp{&process_fullres}}};
    }
    std::optional<double> get_framerate() const { return fps_ > 0 ? std::optional<double>(fps_) : std::nullopt; }
    std::optional<std::pair<size_t, size_t>> get_aspect() const { return std::make_pair((size_t)gw_, (size_t)gh_); }

    bool process_frame(MotionVectors& field, std::vector<RGBA>* out_frame, size_t* out_height, size_t skip_frames)
    {
        size_t dx = 0, dy = 0;
        for (size_t cnt = 0; cnt <= skip_frames; cnt++) {
            int w = 0, h = 0;
            if (!frames_(tmp_, w, h)) throw Error(OFPSB_E_IO, "Failed to grab frame");
            old_frame_.swap(frame_);
            old_gray_.swap(gray_);
            ogw_ = gw_;
            ogh_ = gh_;
            check(ofpsb_mfield_size((size_t)w, (size_t)h, ar_.first, ar_.second, max_mfield_size.first, max_mfield_size.second,
                                    &dx, &dy));
            if (process_fullres) {
                frame_.swap(tmp_);
                gw_ = w;
                gh_ = h;
            } else {
                frame_.resize(dx * dy * 3);
                check(ofpsb_frame_resize(ctx_->get(), tmp_.data(), w, h, w * 3, 3, frame_.data(), (int)dx, (int)dy));
                gw_ = (int)dx;
                gh_ = (int)dy;
            }
            gray_.resize((size_t)gw_ * gh_);
            const bool want_rgba = out_frame && cnt == skip_frames;
            if (want_rgba) out_frame->resize((size_t)gw_ * gh_);
            check(ofpsb_frame_convert(ctx_->get(), frame_.data(), gw_, gh_, gw_ * 3, 3, 0, gray_.data(),
                                      want_rgba ? reinterpret_cast<uint8_t*>(out_frame->data()) : nullptr));
        }
        if (out_frame && out_height) *out_height = (size_t)gh_;
        if (gw_ != ogw_ || gh_ != ogh_) return false;   // first frame / size change (cv-decoder:155-157)
        const size_t npix = (size_t)gw_ * gh_;
        const bool initial = flow_.size() == 2 * npix;
        flow_.resize(2 * npix);
        flow_fn_(old_gray_.data(), gray_.data(), old_frame_.data(), frame_.data(), gw_, gh_, use_rlof, initial, flow_.data());
        const size_t cap = process_fullres ? dx * dy : npix;
        scratch_.resize(cap ? cap : 1);
        size_t n = 0;
        check(ofpsb_cv_flow_frame(ctx_->get(), gray_.data(), gw_, flow_.data(), gw_, gh_, use_rlof ? 0 : 1,
                                  process_fullres ? dx : 0, process_fullres ? dy : 0, scratch_.data(), cap, &n));
        field.insert(field.end(), scratch_.begin(), scratch_.begin() + (std::ptrdiff_t)n);
        return true;
    }

private:
    std::shared_ptr<Context> ctx_;
    FrameSource frames_;
    FlowSource flow_fn_;
    double fps_;
    std::pair<size_t, size_t> ar_;
    std::vector<uint8_t> tmp_, frame_, old_frame_, gray_, old_gray_;
    std::vector<float> flow_;
    MotionVectors scratch_;
    int gw_ = 0, gh_ = 0, ogw_ = 0, ogh_ = 0;
};

}  // namespace ofps_b200

#endif  // OFPS_B200_HPP
