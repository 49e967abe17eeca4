// TEST INFRASTRUCTURE — runs the kernels of ofps_b200/csrc/cv_front.cu on the CPU stand-in of tests/emu/cuda_emu.h
// so that their logic (indices, barriers, border handling, ordering) can be checked against the oracle on the
// GPU-less build container.  Built by tests/test_emu_cv_front.py into tests/emu/_build/; not shipped.
#define OFPSB_EMU 1
#include "../../ofps_b200/csrc/cv_front.cu"

namespace ofpsb {
void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    fputc('\n', stderr);
    va_end(ap);
}
int DevBuf::reserve(size_t bytes)
{
    if (bytes <= cap && ptr) return 0;
    release();
    ptr = malloc(bytes ? bytes : 1);
    cap = bytes;
    return ptr ? 0 : OFPSB_E_NOMEM;
}
void DevBuf::release()
{
    free(ptr);
    ptr = nullptr;
    cap = 0;
}
}  // namespace ofpsb

extern "C" {
int emu_frame_convert(const uint8_t* src, int w, int h, int stride, int channels, int rgb_order, uint8_t* gray,
                      int gray_stride, uint8_t* rgba)
{
    return ofpsb::launch_frame_convert(src, w, h, stride, channels, rgb_order, gray, gray_stride, rgba, nullptr, nullptr);
}
int emu_frame_resize(const uint8_t* src, int sw, int sh, int stride, int channels, uint8_t* dst, int dw, int dh)
{
    return ofpsb::launch_frame_resize(src, sw, sh, stride, channels, dst, dw, dh, dw * channels, nullptr, nullptr);
}
int emu_contrast_mask(const uint8_t* gray, int w, int h, int stride, uint8_t* mask, int mask_stride)
{
    return ofpsb::launch_contrast_mask(gray, w, h, stride, mask, mask_stride, nullptr, nullptr);
}
int emu_flow_entries(const float* flow, size_t flow_stride, const uint8_t* mask, size_t mask_stride, int w, int h, size_t gw,
                     size_t gh, ofps_mv* entries, size_t cap, unsigned long long* count)
{
    static ofpsb::FlowScratch scratch;
    return ofpsb::launch_flow_entries(flow, flow_stride, mask, mask_stride, w, h, gw, gh, entries, cap, count, scratch,
                                      nullptr, nullptr);
}
}
