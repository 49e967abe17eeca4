//! Raw bindings of `include/ofps_b200.h` (hand-written; the header is small and stable).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct ofps_mv {
    pub px: f32,
    pub py: f32,
    pub mx: f32,
    pub my: f32,
}

#[repr(C)]
pub struct ofpsb_ctx {
    _private: [u8; 0],
}

#[repr(C)]
pub struct ofpsb_stream {
    _private: [u8; 0],
}

pub const OFPSB_OK: c_int = 0;
pub const OFPSB_METRIC_SAD: c_int = 0;
pub const OFPSB_METRIC_SSD: c_int = 1;

extern "C" {
    pub fn ofpsb_create(device: c_int, out: *mut *mut ofpsb_ctx) -> c_int;
    pub fn ofpsb_destroy(ctx: *mut ofpsb_ctx);
    pub fn ofpsb_last_error() -> *const c_char;
    pub fn ofpsb_block_match(
        ctx: *mut ofpsb_ctx, prev: *const u8, cur: *const u8, w: c_int, h: c_int, stride: c_int, block: c_int,
        range: c_int, metric: c_int, mv_xy: *mut i16, cost: *mut u32, entries: *mut ofps_mv, n_blocks: *mut usize,
    ) -> c_int;
    pub fn ofpsb_stream_open(
        ctx: *mut ofpsb_ctx, w: c_int, h: c_int, block: c_int, range: c_int, metric: c_int, depth: c_int,
        out: *mut *mut ofpsb_stream,
    ) -> c_int;
    pub fn ofpsb_stream_close(s: *mut ofpsb_stream);
    pub fn ofpsb_stream_blocks(s: *mut ofpsb_stream) -> usize;
    pub fn ofpsb_stream_push(
        s: *mut ofpsb_stream, frame: *const u8, stride: usize, entries: *mut ofps_mv, n_entries: *mut usize,
    ) -> c_int;
    pub fn ofpsb_block_dim(min_size: f32, subdivide: usize, dim: *mut usize) -> c_int;
    pub fn ofpsb_detect_block_motion(
        ctx: *mut ofpsb_ctx, entries: *const ofps_mv, n: usize, min_size: f32, subdivide: usize, target_motion: f32,
        has_motion: *mut c_int, area: *mut usize, dim: *mut usize, field_xy: *mut f32, field_cap_cells: usize,
    ) -> c_int;
    pub fn ofpsb_densify(
        ctx: *mut ofpsb_ctx, entries: *const ofps_mv, n: usize, gw: usize, gh: usize, field_xy: *mut f32,
        counts: *mut f32,
    ) -> c_int;
    pub fn ofpsb_almeida(
        ctx: *mut ofpsb_ctx, entries: *const ofps_mv, n: usize, aspect: f32, fov_y_deg: f32, use_ransac: c_int,
        num_iters: usize, inlier_angle_deg: f32, ransac_samples: usize, seed: u64, quat_wijk: *mut f32,
    ) -> c_int;
    pub fn ofpsb_set_stream(ctx: *mut ofpsb_ctx, stream: *mut c_void) -> c_int;
    // cv-decoder dense-flow front end (cv-decoder/src/lib.rs:84-291)
    pub fn ofpsb_mfield_size(
        frame_w: usize, frame_h: usize, ar_x: usize, ar_y: usize, max_w: usize, max_h: usize, dx: *mut usize,
        dy: *mut usize,
    ) -> c_int;
    pub fn ofpsb_frame_convert(
        ctx: *mut ofpsb_ctx, src: *const u8, w: c_int, h: c_int, stride: c_int, channels: c_int, rgb_order: c_int,
        gray: *mut u8, rgba: *mut u8,
    ) -> c_int;
    pub fn ofpsb_frame_resize(
        ctx: *mut ofpsb_ctx, src: *const u8, sw: c_int, sh: c_int, stride: c_int, channels: c_int, dst: *mut u8,
        dw: c_int, dh: c_int,
    ) -> c_int;
    pub fn ofpsb_contrast_mask(ctx: *mut ofpsb_ctx, gray: *const u8, w: c_int, h: c_int, stride: c_int, mask: *mut u8) -> c_int;
    pub fn ofpsb_flow_entries(
        ctx: *mut ofpsb_ctx, flow_xy: *const f32, mask: *const u8, w: c_int, h: c_int, gw: usize, gh: usize,
        entries: *mut ofps_mv, cap: usize, n: *mut usize,
    ) -> c_int;
    pub fn ofpsb_cv_flow_frame(
        ctx: *mut ofpsb_ctx, gray: *const u8, gray_stride: c_int, flow_xy: *const f32, w: c_int, h: c_int,
        use_mask: c_int, gw: usize, gh: usize, entries: *mut ofps_mv, cap: usize, n: *mut usize,
    ) -> c_int;
    // flow-extract dense field (flow-extract/src/main.rs:72-83)
    pub fn ofpsb_flow_field(ctx: *mut ofpsb_ctx, entries: *const ofps_mv, n: usize, w: usize, h: usize, field_xy: *mut f32) -> c_int;
    pub fn ofpsb_interpolate_empty_cells(sums_xy: *mut f32, counts_xy: *mut f32, w: usize, h: usize) -> c_int;
}

/// Owning handle.  `Send` (the reference requires plugins to be `Send`), not `Sync`: the library allows a
/// context to move between threads but not to be used concurrently.
pub struct Context(pub *mut ofpsb_ctx);
unsafe impl Send for Context {}

impl Context {
    pub fn new(device: i32) -> Result<Self, String> {
        let mut p = std::ptr::null_mut();
        let rc = unsafe { ofpsb_create(device, &mut p) };
        if rc != OFPSB_OK {
            return Err(last_error());
        }
        Ok(Self(p))
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        unsafe { ofpsb_destroy(self.0) }
    }
}

pub fn last_error() -> String {
    unsafe { std::ffi::CStr::from_ptr(ofpsb_last_error()).to_string_lossy().into_owned() }
}

/// `MotionEntry` is a Rust tuple `(Point2<f32>, Vector2<f32>)` whose field order is not ABI-guaranteed:
/// copy into the C layout instead of transmuting.
pub fn to_c(motion: &[((f32, f32), (f32, f32))]) -> Vec<ofps_mv> {
    motion.iter().map(|&((px, py), (mx, my))| ofps_mv { px, py, mx, my }).collect()
}
