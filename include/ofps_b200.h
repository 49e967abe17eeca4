/*
 * ofps_b200.h — C ABI of libofps_b200.so, the B200 (sm_100a) hot path of h33p/ofps.
 *
 * This is the drop-in boundary: the entry points a Rust `cdylib` shim (one per plugin,
 * see rust/ and INTEGRATION.md) binds with `extern "C"` to implement the reference's
 * plugin traits.  Reference interface each group replaces (paths relative to the
 * reference checkout):
 *
 *   ofpsb_block_match*          -> Decoder::process_frame            ofps/src/decoder.rs:54-59
 *                                  (output convention of av-decoder/src/lib.rs:404-419;
 *                                  the SAD search itself is what the H.264 encoder did
 *                                  upstream of av-decoder — the reference has no such code)
 *   ofpsb_densify*              -> MotionFieldDensifier::add_vector + MotionField::from
 *                                  ofps/src/motion_field.rs:164-190, 297-308
 *   ofpsb_detect_block_motion*  -> Detector::detect_motion           ofps/src/detection.rs:11
 *                                  as implemented by block-motion-detector/src/lib.rs:49-119
 *   ofpsb_almeida*              -> Estimator::estimate               ofps/src/estimator.rs:19-24
 *                                  as implemented by almeida-estimator/src/lib.rs:100-251
 *   ofpsb_mvec_* / ofpsb_flo_*  -> .mvec wire format motion-extract/src/main.rs:23-35,
 *                                  motion-loader/src/lib.rs:46-65; .flo writer used by
 *                                  flow-extract/src/main.rs:122
 *
 * Conventions
 *   - Plain pointers and sizes only.  Entry points without a suffix take HOST pointers and
 *     are synchronous on return; `_dev` entry points take DEVICE pointers, enqueue on the
 *     context's stream and return without synchronising (call ofpsb_sync()).
 *   - Every entry point returns 0 (OFPSB_OK) or a negative OFPSB_E_* code; the message is
 *     available through ofpsb_last_error() (thread-local).
 *   - A context is bound to one device, may be used from any host thread, never from two
 *     threads at once (the reference's plugin objects are Send, not Sync:
 *     ofps/src/plugins/mod.rs:78-85).
 *   - There is NO CPU fallback: without a usable sm_100-class GPU ofpsb_create() fails.
 */
#ifndef OFPS_B200_H
#define OFPS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OFPSB_OK 0
#define OFPSB_E_INVALID (-1)   /* bad argument */
#define OFPSB_E_CUDA (-2)      /* CUDA runtime / launch failure */
#define OFPSB_E_NOMEM (-3)     /* allocation failure */
#define OFPSB_E_NODEVICE (-4)  /* no CUDA device / not an sm_100 part */
#define OFPSB_E_IO (-5)        /* file I/O failure (.mvec / .flo) */
#define OFPSB_E_CAPACITY (-6)  /* caller-provided output buffer too small */

#define OFPSB_METRIC_SAD 0
#define OFPSB_METRIC_SSD 1

/* MotionEntry (ofps/src/decoder.rs:40) flattened as motion-extract writes it
 * (motion-extract/src/main.rs:27-29): position then motion, normalised [0,1] units. */
typedef struct { float px, py, mx, my; } ofps_mv;

typedef struct ofpsb_ctx ofpsb_ctx;

/* ---------------------------------------------------------------- context */
int ofpsb_create(int device, ofpsb_ctx **out);
void ofpsb_destroy(ofpsb_ctx *ctx);
const char *ofpsb_last_error(void);
const char *ofpsb_version(void);
/* Borrow an external cudaStream_t (e.g. the caller's current stream); NULL restores the
 * context's own non-blocking stream. */
int ofpsb_set_stream(ofpsb_ctx *ctx, void *cuda_stream);
int ofpsb_sync(ofpsb_ctx *ctx);
/* sm_count, L2 bytes, total global memory bytes; any pointer may be NULL. */
int ofpsb_device_info(ofpsb_ctx *ctx, int *sm_count, size_t *l2_bytes, size_t *mem_bytes, int *cc_major, int *cc_minor);
/* Number of kernels launched by this context since creation (for gpu_launches accounting). */
uint64_t ofpsb_launch_count(ofpsb_ctx *ctx);
/* Counters of the LAST pruned block-match launch (needs option "block_match_stats" = 1; synchronises):
 * out[0] = blocks seen, out[1] = blocks decided by the pruning pass, out[2] = exact SAD evaluations it
 * spent, out[3] = blocks sent to the exhaustive work list. */
int ofpsb_block_match_stats(ofpsb_ctx *ctx, uint64_t out[4]);
/* Device time of the kernels of the LAST default-path (fused SEA) block-match launch, from CUDA events recorded on the
 * launching stream (needs option "block_match_profile" = 1; synchronises): out[0] = SEA kernel, out[1] = exhaustive
 * work-list kernel, milliseconds. */
int ofpsb_block_match_kernel_ms(ofpsb_ctx *ctx, float out[2]);
/* The stream the context currently enqueues on (a cudaStream_t), for event timing by the caller. */
void *ofpsb_get_stream(ofpsb_ctx *ctx);
/* Tuning / test knobs; unknown keys return OFPSB_E_INVALID.
 *   "densify_path"        0 = by size (default), 1 = force the scan path, 2 = force the sort path
 *   "block_match_kernel"  0 = best instance (TMA-staged, else LDG-staged, else generic; default),
 *                         1 = force the generic kernel, 2 = force the LDG-staged tile kernel,
 *                         3 = TMA-staged, alternative tile shape
 *   "batch_chunk_pairs"   pairs per pipelined chunk in ofpsb_block_match_batch (0 = automatic)
 *   "block_match_prune"   1 = exact successive-elimination pruning in front of the exhaustive SAD search
 *                         (default; same results, data-dependent speed), 0 = always exhaustive
 *   "block_match_stats"   1 = count blocks / decided blocks / exact evaluations of the pruning pass
 *   "block_match_chunk_pairs"  pairs per chunk of the round-1 pruning pipeline (0 = whole batch in one chunk, default)
 *   "block_match_pruner"  0 = fused SEA kernel (default: SAD, 8x8 / 16x16 blocks, +-8 / +-16 / +-32), 1 = round-1 pipeline
 *   "block_match_adaptive" 1 = skip the SEA kernel for 15 launches after one that left most blocks undecided (default)
 *   "block_match_tile_h"  0 = by launch size (default), 32 / 64 = tile rows of the SEA kernel (+-32 always uses 32)
 *   "block_match_prefetch_tiles"  L2 prefetch distance of the SEA kernel in tiles (-1 = default)
 *   "block_match_profile" 1 = time the SEA / work-list kernels with events (ofpsb_block_match_kernel_ms)
 *   "detect_union_find"   1 = union-find detector for every size (default: one-warp flood fill up to 32 x 32 cells)
 *   "almeida_stepwise"    1 = one launch per solver step instead of the persistent cooperative grid
 *   "almeida_cluster"     0 = never use the one-cluster solver (default 1: fields of up to 16,384 vectors run in one
 *                         thread-block cluster, one entry per thread, partial sums exchanged through DSMEM)
 * Environment: OFPSB_COPY_THREADS = host threads (caller included) that copy a pageable frame into the pinned
 * staging ring of ofpsb_stream_* (default: half the cores, at most 8). */
int ofpsb_set_option(ofpsb_ctx *ctx, const char *key, long long value);

/* Pinned host memory (page-locked; makes the batched host entry points copy asynchronously) and
 * plain device memory helpers for callers without their own CUDA runtime binding. */
int ofpsb_host_alloc(void **out, size_t bytes);
void ofpsb_host_free(void *p);
int ofpsb_dev_alloc(ofpsb_ctx *ctx, void **out, size_t bytes);
void ofpsb_dev_free(ofpsb_ctx *ctx, void *p);
/* Asynchronous on the context's stream (ofpsb_sync() to wait). */
int ofpsb_copy_to_device(ofpsb_ctx *ctx, void *d_dst, const void *h_src, size_t bytes);
int ofpsb_copy_to_host(ofpsb_ctx *ctx, void *h_dst, const void *d_src, size_t bytes);

/* ----------------------------------------------------------- block matcher
 * Exhaustive block matching of `cur` against `prev` (u8 luma, `stride` bytes per row).
 * Blocks block x block at (bx*block, by*block) in cur, full blocks only; candidates
 * (dx,dy) in [-range,range]^2 whose window lies fully inside prev; cost SAD or SSD;
 * winner = lexicographic min of (cost, dx^2+dy^2, dy, dx).  Outputs, per block in raster
 * order (any may be NULL): mv_xy[2i] = dx, mv_xy[2i+1] = dy; cost[i]; entries[i] =
 * { pos = (block centre + (dx,dy)) * (1/W,1/H), motion = (dx,dy) * -(1/W,1/H) }.
 * block in {4..64, multiple of 4}, range in [0,63].  *n_blocks (optional) receives the
 * block count (w/block)*(h/block). */
int ofpsb_block_match(ofpsb_ctx *ctx, const uint8_t *prev, const uint8_t *cur, int w, int h, int stride,
                      int block, int range, int metric,
                      int16_t *mv_xy, uint32_t *cost, ofps_mv *entries, size_t *n_blocks);

/* Batched host entry point: n_pairs independent pairs, pair i at prev + i*pair_stride
 * (bytes).  Outputs are n_pairs * n_blocks long.  Copies and kernels are pipelined on
 * two streams; pinned host memory makes the copies asynchronous. */
int ofpsb_block_match_batch(ofpsb_ctx *ctx, const uint8_t *prev, const uint8_t *cur, int w, int h, int stride,
                            size_t pair_stride, int n_pairs, int block, int range, int metric,
                            int16_t *mv_xy, uint32_t *cost, ofps_mv *entries, size_t *n_blocks);

/* Device entry point (batched).  Output pointers are device memory, n_pairs*n_blocks. */
int ofpsb_block_match_dev(ofpsb_ctx *ctx, const uint8_t *d_prev, const uint8_t *d_cur, int w, int h, int stride,
                          size_t pair_stride, int n_pairs, int block, int range, int metric,
                          int16_t *d_mv_xy, uint32_t *d_cost, ofps_mv *d_entries);

/* Device entry point for one horizontal strip of a spatially tiled frame (multi-GPU).
 * d_cur: first row of the strip (strip_h rows, strip_h % block == 0 except for the last
 * strip).  d_prev: the same rows of the previous frame, with halo_top valid rows stored
 * BEFORE d_prev (negative row offsets) and halo_bottom valid rows after the strip.
 * y_offset / full_h place the strip in the whole frame (entries are normalised by full_h;
 * a candidate is legal iff its window lies inside [y_offset-halo_top,
 * y_offset+strip_h+halo_bottom), which equals the whole-frame rule when the halos are
 * min(range, rows available)). */
int ofpsb_block_match_strip_dev(ofpsb_ctx *ctx, const uint8_t *d_prev, const uint8_t *d_cur, int w, int strip_h,
                                int stride, int halo_top, int halo_bottom, int y_offset, int full_h,
                                int block, int range, int metric,
                                int16_t *d_mv_xy, uint32_t *d_cost, ofps_mv *d_entries);

/* Batched strip: n_pairs pairs of one strip, pair i at d_prev + i*pair_stride / d_cur + i*pair_stride (for a
 * stream stored as [halo_top | strip | halo_bottom] rows per frame, d_cur = d_prev + pair_stride).  Outputs
 * are n_pairs * (w/block)*(strip_h/block) long. */
int ofpsb_block_match_strip_batch_dev(ofpsb_ctx *ctx, const uint8_t *d_prev, const uint8_t *d_cur, int w, int strip_h,
                                      int stride, size_t pair_stride, int n_pairs, int halo_top, int halo_bottom,
                                      int y_offset, int full_h, int block, int range, int metric,
                                      int16_t *d_mv_xy, uint32_t *d_cost, ofps_mv *d_entries);

/* -------------------------------------------------------------- streaming decoder
 * The device side of `Decoder::process_frame` (ofps/src/decoder.rs:45-73) for a block-matching decoder: frames go in
 * one at a time, the MotionEntry list of (previous frame, this frame) comes out.  Each frame is uploaded ONCE (the
 * previous one is still in HBM); `frame` may be pageable memory (it is staged through a pinned ring by helper threads)
 * or page-locked (handed to the copy engine as is: it must then stay unchanged until the pair it completes has been
 * collected).  One caller thread at a time (the reference's plugins are Send,
 * not Sync).  depth = frames in flight (>= 3; 4 or more lets submit run one frame ahead of collect). */
typedef struct ofpsb_stream ofpsb_stream;
int ofpsb_stream_open(ofpsb_ctx *ctx, int w, int h, int block, int range, int metric, int depth, ofpsb_stream **out);
void ofpsb_stream_close(ofpsb_stream *s);
/* Entries per pair: (w/block)*(h/block). */
size_t ofpsb_stream_blocks(ofpsb_stream *s);
/* Synchronous form: *n_entries = 0 for the first frame of the stream (`Ok(false)`: no motion vectors yet), else the
 * entry count, `entries` filled (host memory, ofpsb_stream_blocks long). */
int ofpsb_stream_push(ofpsb_stream *s, const uint8_t *frame, size_t stride, ofps_mv *entries, size_t *n_entries);
/* Pipelined form: submit enqueues upload + kernels + read-back and returns; collect waits for the OLDEST outstanding
 * pair (*n_entries = 0 when nothing is outstanding).  At most depth-2 pairs may be outstanding. */
int ofpsb_stream_submit(ofpsb_stream *s, const uint8_t *frame, size_t stride);
int ofpsb_stream_collect(ofpsb_stream *s, ofps_mv *entries, size_t *n_entries);

/* -------------------------------------------------------------- spatial tiling over the GPUs of one node
 * One LARGE frame pair cut into horizontal strips of whole block rows, one rank (process or thread) per GPU
 * (the multi-GPU form of the Decoder boundary, ofps/src/decoder.rs:45-73: `process_frame` of a
 * tiled decoder calls ofpsb_tiled_upload / _publish / _match on its strip).  There is no exchange step: every rank
 * exports its frame buffer once, maps its two neighbours' buffers, and the matching kernel reads the `range` halo
 * rows of the previous frame straight from the neighbours' HBM over NVLink (tensor maps on peer-mapped memory).
 *
 *   create (same w, h, block, range, n_slots on every rank)  ->  export  ->  [exchange the 128-byte blobs]  ->
 *   connect (IPC, other processes)  or  connect_local (ranks of ONE process)  ->  per frame: upload / write the slot,
 *   publish  ->  match(prev_slot, cur_slot).
 *
 * Strips: block rows split evenly, the first (h/block) % world ranks take one more; the last rank also keeps the
 * frame's remainder rows (h % block).  SAD only.  Outputs cover this rank's strip: nby*nbx blocks in raster order
 * (strip order = raster order of the whole frame), entries normalised by the whole frame (av-decoder/src/lib.rs:404-419).
 * Bit-identical to ofpsb_block_match on the whole frame. */
typedef struct ofpsb_tiled ofpsb_tiled;
#define OFPSB_TILED_HANDLE_BYTES 128
int ofpsb_tiled_create(ofpsb_ctx *ctx, int rank, int world, int w, int h, int block, int range, int n_slots,
                       ofpsb_tiled **out);
void ofpsb_tiled_destroy(ofpsb_tiled *t);
/* Geometry of this rank's strip: first pixel row, pixel rows matched (nby*block), rows stored (>= rows: the last
 * rank keeps the remainder), blocks per row / block rows, byte stride of the device slots.  Any pointer may be NULL. */
int ofpsb_tiled_info(ofpsb_tiled *t, int *y0, int *rows, int *own_rows, int *nbx, int *nby, int *stride);
/* OFPSB_TILED_HANDLE_BYTES describing this rank's buffer for its neighbours (cudaIpcMemHandle + geometry). */
int ofpsb_tiled_export(ofpsb_tiled *t, void *handle);
/* Map the neighbours (blobs from THEIR ofpsb_tiled_export; NULL at the frame's top / bottom). */
int ofpsb_tiled_connect(ofpsb_tiled *t, const void *up_handle, const void *down_handle);
/* Same for ranks that live in this process (one thread per GPU, or several strips on one GPU in tests). */
int ofpsb_tiled_connect_local(ofpsb_tiled *t, ofpsb_tiled *up, ofpsb_tiled *down);
/* Device pointer to the first own row of frame slot `slot` (own_rows rows of `stride` bytes) — for producers that
 * write the strip on the device.  ofpsb_tiled_upload copies own_rows rows from host memory (async on the context's
 * stream). */
void *ofpsb_tiled_slot_ptr(ofpsb_tiled *t, int slot);
int ofpsb_tiled_upload(ofpsb_tiled *t, int slot, const uint8_t *host_rows, size_t host_stride);
/* After the slot's rows are written (stream order): tell both neighbours (a 4-byte epoch in THEIR memory). */
int ofpsb_tiled_publish(ofpsb_tiled *t, int slot);
/* Match the strip of (prev_slot, cur_slot); device outputs (any may be NULL), asynchronous on the context's stream.
 * wait_neighbours != 0: first wait (on the device) until both neighbours have published prev_slot as often as this
 * rank has; 0: the caller has synchronised the ranks itself. */
int ofpsb_tiled_match(ofpsb_tiled *t, int prev_slot, int cur_slot, ofps_mv *d_entries, int16_t *d_mv_xy, uint32_t *d_cost,
                      int wait_neighbours);
/* A stream of tiled frames in consecutive slots: pairs (first_slot + i, first_slot + i + 1), i < n_pairs, in ONE launch
 * sequence (the per-launch latency of a strip is paid once per batch).  Outputs: n_pairs * nby*nbx, pair-major. */
int ofpsb_tiled_match_stream(ofpsb_tiled *t, int first_slot, int n_pairs, ofps_mv *d_entries, int16_t *d_mv_xy,
                             uint32_t *d_cost, int wait_neighbours);

/* -------------------------------------------------------------- densifier
 * field_xy: gw*gh*2 floats, cell-major [x0,y0,x1,y1,...], cell = y*gw+x
 * (MotionField::as_slice, ofps/src/motion_field.rs:42-49).  counts (optional): same
 * shape, the densifier's count rows.  Bit-exact with the reference order of additions. */
int ofpsb_densify(ofpsb_ctx *ctx, const ofps_mv *entries, size_t n, size_t gw, size_t gh,
                  float *field_xy, float *counts);
int ofpsb_densify_dev(ofpsb_ctx *ctx, const ofps_mv *d_entries, size_t n, size_t gw, size_t gh,
                      float *d_field_xy, float *d_counts);

/* The flow-extract dense pipeline (flow-extract/src/main.rs:72-83): densify into a w x h field (GPU, bit-exact
 * order), MotionFieldDensifier::interpolate_empty_cells (ofps/src/motion_field.rs:193-294), sum ./ counts.
 * The hole fill is strictly sequential by definition (ordered set keyed by the changing neighbour counts) and
 * runs on the host, as in the reference; see csrc/hole_fill.cu.  field_xy: w*h*2 floats. */
int ofpsb_flow_field(ofpsb_ctx *ctx, const ofps_mv *entries, size_t n, size_t w, size_t h, float *field_xy);
/* The hole fill alone, in place on a densifier state (sums and counts, 2*w*h floats each).  Host arithmetic
 * only: needs no context and no device. */
int ofpsb_interpolate_empty_cells(float *sums_xy, float *counts_xy, size_t w, size_t h);

/* --------------------------------------------------------------- detector
 * BlockMotionDetection::detect_motion.  *has_motion = 1 for Some, 0 for None; *area =
 * cell count of the winning island (0 when None); *dim = block_dim; field_xy receives
 * dim*dim*2 floats (zeros when None) and must hold field_cap_cells cells
 * (OFPSB_E_CAPACITY otherwise; *dim is still written). */
int ofpsb_block_dim(float min_size, size_t subdivide, size_t *dim);
int ofpsb_detect_block_motion(ofpsb_ctx *ctx, const ofps_mv *entries, size_t n,
                              float min_size, size_t subdivide, float target_motion,
                              int *has_motion, size_t *area, size_t *dim,
                              float *field_xy, size_t field_cap_cells);
/* entries in device memory; results still returned to host (synchronous). */
int ofpsb_detect_block_motion_dev(ofpsb_ctx *ctx, const ofps_mv *d_entries, size_t n,
                                  float min_size, size_t subdivide, float target_motion,
                                  int *has_motion, size_t *area, size_t *dim,
                                  float *field_xy, size_t field_cap_cells);

/* -------------------------------------------------------------- estimator
 * AlmeidaEstimator::estimate.  Camera passed as (aspect, fov_y_deg) — recoverable from
 * StandardCamera::aspect_ratio() and fov().1 (ofps/src/camera.rs:166-177).
 * use_ransac = 0: solve_ypr_given; 1: solve_ypr_ransac with the seeded counter RNG
 * documented in DESIGN.md (the reference's thread_rng is not reproducible).
 * quat_wijk: rotation as (w,i,j,k); translation is always zero in the reference. */
int ofpsb_almeida(ofpsb_ctx *ctx, const ofps_mv *entries, size_t n, float aspect, float fov_y_deg,
                  int use_ransac, size_t num_iters, float inlier_angle_deg, size_t ransac_samples,
                  uint64_t seed, float quat_wijk[4]);
int ofpsb_almeida_dev(ofpsb_ctx *ctx, const ofps_mv *d_entries, size_t n, float aspect, float fov_y_deg,
                      int use_ransac, size_t num_iters, float inlier_angle_deg, size_t ransac_samples,
                      uint64_t seed, float quat_wijk[4]);

/* --------------------------------------------------- fused per-frame paths
 * Detection-tab frame (ofps-suite/src/app/detection.rs:92-168): frame pair -> motion
 * vectors -> detector, one call, intermediates stay in HBM.  entries (optional, host)
 * receives the n_blocks motion entries. */
int ofpsb_frame_detect(ofpsb_ctx *ctx, const uint8_t *prev, const uint8_t *cur, int w, int h, int stride,
                       int block, int range, int metric,
                       float min_size, size_t subdivide, float target_motion,
                       ofps_mv *entries, size_t *n_blocks,
                       int *has_motion, size_t *area, size_t *dim, float *field_xy, size_t field_cap_cells);

/* ------------------------------------------- cv-decoder dense-flow front end
 * What CvDecoder::process_frame does around the third-party optical-flow call
 * (cv-decoder/src/lib.rs:84-291); the flow itself (OpenCV Farneback / RLOF) is an input.
 *
 * ofpsb_mfield_size: motion-field size from the frame size, the aspect-ratio scale and the
 * "Width" / "Height" properties (cv-decoder/src/lib.rs:90-118).  Host arithmetic only. */
int ofpsb_mfield_size(size_t frame_w, size_t frame_h, size_t ar_x, size_t ar_y, size_t max_w, size_t max_h,
                      size_t *dx, size_t *dy);
/* cvtColor(frame, COLOR_BGR2GRAY) (cv-decoder/src/lib.rs:138; OpenCV 4 fixed point, bit-exact) and the
 * RGBA out_frame (:145-153).  src: u8, `channels` = 3 or 4 interleaved, `stride` bytes per row;
 * rgb_order != 0 reads R,G,B instead of B,G,R.  gray: w*h bytes, rgba: w*h*4 bytes; either may be NULL. */
int ofpsb_frame_convert(ofpsb_ctx *ctx, const uint8_t *src, int w, int h, int stride, int channels, int rgb_order,
                        uint8_t *gray, uint8_t *rgba);
int ofpsb_frame_convert_dev(ofpsb_ctx *ctx, const uint8_t *d_src, int w, int h, int stride, int channels,
                            int rgb_order, uint8_t *d_gray, int gray_stride, uint8_t *d_rgba);
/* imgproc::resize(frame, (dw, dh), INTER_LINEAR) on the 8-bit frame, the "Process Fullres" = off path
 * (cv-decoder/src/lib.rs:127-135): OpenCV's fixed-point bilinear arithmetic, bit-exact.  Reductions only
 * (dw <= sw, dh <= sh — the reference never enlarges; OFPSB_E_INVALID otherwise).  dst: dw*dh*channels bytes. */
int ofpsb_frame_resize(ofpsb_ctx *ctx, const uint8_t *src, int sw, int sh, int stride, int channels,
                       uint8_t *dst, int dw, int dh);
int ofpsb_frame_resize_dev(ofpsb_ctx *ctx, const uint8_t *d_src, int sw, int sh, int stride, int channels,
                           uint8_t *d_dst, int dw, int dh, int dst_stride);
/* Contrast mask of the Farneback path (cv-decoder/src/lib.rs:204-236): Sobel(CV_32F, 1, 1, ksize 5) ->
 * threshold(20, 255, BINARY) -> dilate(11x11 MORPH_ELLIPSE), all with BORDER_REFLECT_101, fused in one
 * kernel.  mask: w*h bytes, 255 where the reference's f32 mask is 255, else 0 (bit-exact with OpenCV). */
int ofpsb_contrast_mask(ofpsb_ctx *ctx, const uint8_t *gray, int w, int h, int stride, uint8_t *mask);
int ofpsb_contrast_mask_dev(ofpsb_ctx *ctx, const uint8_t *d_gray, int w, int h, int stride,
                            uint8_t *d_mask, int mask_stride);
/* Dense flow image -> MotionEntry list (cv-decoder/src/lib.rs:238-291).  flow: h rows of w (fx,fy) f32
 * pairs in pixels; mask (optional, w*h bytes): pixels with mask == 0 are skipped (the RLOF path passes
 * NULL).  gw == gh == 0 ("Process Fullres" off): one entry per kept pixel in raster order, pos =
 * (x+0.5, y+0.5) .* (1/w, 1/h), motion = flow .* (1/w, 1/h).  Otherwise the kept pixels go through a
 * gw x gh MotionFieldDensifier and one entry per touched cell comes out in (x, y) lexicographic order
 * (the reference's BTreeSet), pos = cell centre, motion = cell mean; sums follow the reference's raster
 * order in un-fused f32 (bit-exact).  *n receives the entry count; OFPSB_E_CAPACITY if it exceeds cap
 * (the first cap entries are still written). */
int ofpsb_flow_entries(ofpsb_ctx *ctx, const float *flow_xy, const uint8_t *mask, int w, int h,
                       size_t gw, size_t gh, ofps_mv *entries, size_t cap, size_t *n);
/* Device variant: strides in elements (floats / bytes); d_entries device memory, 16-byte aligned (cudaMalloc /
 * ofpsb_dev_alloc pointers are); *n returned to host (one stream synchronisation).  16-byte aligned flow / mask rows
 * (base and pitch) take the asynchronous-copy staging path, anything else the element-wise one. */
int ofpsb_flow_entries_dev(ofpsb_ctx *ctx, const float *d_flow_xy, size_t flow_stride, const uint8_t *d_mask,
                           size_t mask_stride, int w, int h, size_t gw, size_t gh,
                           ofps_mv *d_entries, size_t cap, size_t *n);
/* The whole post-flow stage of one frame in one call: gray (+ flow) in, entries out; mask and cell sums
 * stay in HBM.  use_mask = 1 for the Farneback path, 0 for RLOF. */
int ofpsb_cv_flow_frame(ofpsb_ctx *ctx, const uint8_t *gray, int gray_stride, const float *flow_xy, int w, int h,
                        int use_mask, size_t gw, size_t gh, ofps_mv *entries, size_t cap, size_t *n);

/* ------------------------------------------------------ interchange files
 * .mvec: per frame u32 LE count, then count x 4 f32 LE (motion-extract/src/main.rs:23-35). */
int ofpsb_mvec_append(const char *path, const ofps_mv *entries, size_t n, int truncate);
/* Reads frame `frame_index`; *n receives its entry count; entries may be NULL to query. */
int ofpsb_mvec_read(const char *path, size_t frame_index, ofps_mv *entries, size_t cap, size_t *n);
/* Middlebury .flo ("PIEH", w, h, then w*h*2 f32), as OpenCV's writeOpticalFlow emits. */
int ofpsb_flo_write(const char *path, const float *field_xy, size_t w, size_t h);

#ifdef __cplusplus
}
#endif
#endif /* OFPS_B200_H */
