/*
 * moonb200.h - C ABI of libmoonb200.so, the B200-native replacement for the
 * data-parallel hot path of MoonRTX (albireo77/moonrtx).
 *
 * The reference has no FFI of its own: its boundary is the Python object
 * `self.rt` (plotoptix.TkOptiX, created at moonrtx/moon_renderer.py:571-575) and
 * three free functions of moonrtx/data_loader.py.  Each entry point below names
 * the reference call(s) it serves.  The Python side (moonrtx_b200/) binds these
 * with ctypes only; there are no torch types, no callbacks and no C++ types in
 * any signature.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; mrtx_last_error() returns
 *     the message of the last failing call on the calling thread;
 *   - pointers are HOST pointers unless the parameter name ends in `_dev`;
 *   - one context per GPU; a context is not thread-safe (the Python drop-in
 *     serialises calls with its `_padlock`, like PlotOptiX);
 *   - all work of a context is issued on the context's stream
 *     (mrtx_set_stream / mrtx_get_stream); functions that return data to host
 *     memory synchronise that stream before returning, all others are asynchronous;
 *   - images are row-major, row 0 = top of the frame.
 */
#ifndef MOONB200_H
#define MOONB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRTX_ABI_VERSION 1

typedef struct mrtx_ctx mrtx_ctx;

enum {
    MRTX_OK = 0,
    MRTX_ERR_INVALID = -1,      /* bad argument (ValueError on the Python side)            */
    MRTX_ERR_CUDA = -2,         /* CUDA runtime failure                                    */
    MRTX_ERR_STATE = -3,        /* call out of order (e.g. render before set_displacement) */
    MRTX_ERR_NCCL = -4
};

/* ---- library / context ------------------------------------------------------ */
int         mrtx_abi_version(void);
const char* mrtx_last_error(void);
int  mrtx_device_count(int* count);
/* TkOptiX(...) constructor, moon_renderer.py:571: one context per GPU.              */
int  mrtx_create(int device, mrtx_ctx** out_ctx);
/* rt.close(), moon_renderer.py:883                                                   */
int  mrtx_destroy(mrtx_ctx* ctx);
int  mrtx_synchronize(mrtx_ctx* ctx);
/* Issue the context's work on a caller-owned cudaStream_t (e.g. torch's current
 * stream, so torch.cuda.Event brackets it).  NULL restores the context's own.       */
int  mrtx_set_stream(mrtx_ctx* ctx, void* cuda_stream);
int  mrtx_get_stream(mrtx_ctx* ctx, void** cuda_stream);
int  mrtx_device_props(mrtx_ctx* ctx, int* sm_count, int* l2_bytes, size_t* hbm_bytes);

/* CUDA-event stopwatch on the context's stream (bench.py times kernels with it).    */
int  mrtx_timer_start(mrtx_ctx* ctx);
int  mrtx_timer_stop(mrtx_ctx* ctx, float* elapsed_ms);   /* synchronises the stop event */

/* ---- raw device / pinned memory (so the host side needs no torch) ------------- */
int  mrtx_dev_alloc(mrtx_ctx* ctx, size_t bytes, void** out_dev);
int  mrtx_dev_free(mrtx_ctx* ctx, void* ptr_dev);
int  mrtx_host_alloc(size_t bytes, void** out_pinned);
int  mrtx_host_free(void* pinned);
int  mrtx_h2d(mrtx_ctx* ctx, void* dst_dev, const void* src, size_t bytes);   /* async  */
int  mrtx_d2h(mrtx_ctx* ctx, void* dst, const void* src_dev, size_t bytes);   /* syncs  */
int  mrtx_l2_flush(mrtx_ctx* ctx);   /* overwrite a scratch buffer larger than L2 (bench hygiene) */

/* ---- data_loader hot path ------------------------------------------------------
 * load_elevation_data, moonrtx/data_loader.py:215-242 (array part: int16 block mean
 * in numpy's two-stage float32 order, *scale, +1, /max).  W, H = source size; both
 * must be divisible by ds (else MRTX_ERR_INVALID, the reference's reshape ValueError).
 * out = float32 [H/ds][W/ds]; *radius_scale = the float32 maximum before division.
 * Bit-exact against the reference for every ds >= 1.                                 */
int  mrtx_downscale_i16(mrtx_ctx* ctx, const int16_t* src, int W, int H, int ds,
                        float* out, float* radius_scale);
/* The same straight from the LDEM file (replaces plotoptix.utils.read_image, data_loader.py:206, for the file the
 * reference ships: an uncompressed little-endian single-channel 16-bit strip TIFF, classic or BigTIFF).  mrtx_tiff_info
 * parses the directory on the host (no device work): size, bits per sample, and whether the strips can be streamed as they
 * lie.  mrtx_downscale_tiff_i16 then reads them with pread() into the pinned staging buffers of the banded upload - no
 * decoded copy of the 8.5 GB map in host memory - and, if npy_cache_path is given, writes the result into that file in
 * numpy's .npy format while it comes down from the device (the downscale cache, data_loader.py:88-95); on any error no
 * cache file is left behind.  Not streamable (compressed, tiled, big-endian ...): MRTX_ERR_INVALID, decode on the host.  */
int  mrtx_tiff_info(const char* path, int* W, int* H, int* bits_per_sample, int* streamable);
int  mrtx_downscale_tiff_i16(mrtx_ctx* ctx, const char* path, int ds, float* out, float* radius_scale,
                             const char* npy_cache_path);
int  mrtx_downscale_i16_dev(mrtx_ctx* ctx, const int16_t* src_dev, int W, int H, int ds,
                            float* out_dev, float* radius_scale /* host, may be NULL */);

/* load_color_data, moonrtx/data_loader.py:331 (cv2 IMREAD_REDUCED_COLOR_k of a TIFF =
 * round-half-up mean of the central 2x2 of each k x k block) followed by
 * _moon_texture, :345-368 (256-entry LUT, BGR -> RGBA, alpha 255).
 * bgr = uint8 [H][W][3]; k in {1,2,4,8}; W, H divisible by k; out = uint8 [H/k][W/k][4]. */
int  mrtx_color_reduce_lut(mrtx_ctx* ctx, const uint8_t* bgr, int W, int H, int k,
                           const uint8_t lut[256], uint8_t* out_rgba);
int  mrtx_color_reduce_lut_dev(mrtx_ctx* ctx, const uint8_t* bgr_dev, int W, int H, int k,
                               const uint8_t lut[256], uint8_t* out_rgba_dev);

/* Synthetic LOLA-shaped inputs generated in HBM (bench/test data; the real 9 GB
 * LDEM cannot be downloaded offline).                                                 */
int  mrtx_synth_ldem_i16_dev(mrtx_ctx* ctx, int16_t* out_dev, int W, int H, uint32_t seed);
int  mrtx_synth_color_bgr_dev(mrtx_ctx* ctx, uint8_t* out_dev, int W, int H, uint32_t seed);

/* ---- scene (the rt.* calls of moon_renderer.py:570-650, 852-860) ----------------- */
/* rt.set_displacement("moon", float32[h][w]), moon_renderer.py:624.  Copies the map to
 * the device and builds the max-height pyramid.                                       */
int  mrtx_set_displacement_f32(mrtx_ctx* ctx, const float* map, int W, int H);
int  mrtx_set_displacement_f32_dev(mrtx_ctx* ctx, const float* map_dev, int W, int H, int copy);
/* Same surface from the raw LDEM counts: D = fl32(fl32(fl32(c*scale)+1)/radius_scale),
 * i.e. exactly the float32 map the reference would hold at downscale 1, at half the
 * memory (8.5 GB instead of 17 GB for 92160x46080).                                   */
int  mrtx_set_displacement_i16(mrtx_ctx* ctx, const int16_t* map, int W, int H,
                               float scale, float radius_scale);
int  mrtx_set_displacement_i16_dev(mrtx_ctx* ctx, const int16_t* map_dev, int W, int H,
                                   float scale, float radius_scale, int copy);
/* rt.set_texture_2d("moon_color" | "frame_overlay", uint8[h][w][4]),
 * moon_renderer.py:614, renderer_video.py:137.  slot 0 = moon_color (bilinear),
 * slot 1 = frame_overlay (nearest, must match the frame size); NULL clears.           */
int  mrtx_set_texture_rgba8(mrtx_ctx* ctx, int slot, const uint8_t* rgba, int W, int H);
/* rt.set_graph / update_graph / delete_geometry (renderer_labels.py:263-305, 324-325, renderer_pins.py:18-55, 131): the
 * overlay tubes of the grid, its number labels, the feature labels and the pins.  ALL visible segments of all graphs in one
 * list, scene space: n x 12 floats = (ax, ay, az, r, bx, by, bz, 0, red, green, blue, 0).  A segment is a capsule of radius
 * r; it is flat-shaded (the colour is the radiance of a camera sample that meets it before the surface), it casts no shadow
 * and receives none (the reference's material lets shadow rays through, renderer_labels.py:133-139).  n = 0 removes them.
 * The list is binned to 16 x 16-pixel screen tiles for the camera of every launch.                                        */
int  mrtx_set_tubes(mrtx_ctx* ctx, const float* segments, int n);
/* What rays that miss the Moon see.
 * rt.set_background_mode("TextureEnvironment") + rt.set_background(float32[h][w][3] in [0, 1], gamma=g,
 * rt_format="UByte4"), moon_renderer.py:602-609: the star map becomes an 8-bit environment texture of linear radiance
 * v^g (slot 2 of mrtx_set_texture_rgba8 takes such a texture directly), looked up by ray direction - equirectangular,
 * scene +Z up, -Y at longitude 0, bilinear; NULL = rt.set_background(0), black.
 * rt.set_data / update_data("sun_disk", pos=, r=, c=), moon_renderer.py:643-650, 855: the visible Sun disk, a flat-shaded
 * sphere in scene space that primary rays see and shadow rays do not (radius <= 0 removes it).
 * mrtx_resize_cubic_f32: cv2.resize(INTER_CUBIC) + clip of load_starmap (data_loader.py:412-415), host buffers.      */
int  mrtx_set_background_f32(mrtx_ctx* ctx, const float* rgb, int W, int H, float gamma);
int  mrtx_read_background_rgba8(mrtx_ctx* ctx, uint8_t* out, int* W, int* H);
int  mrtx_set_sun_disk(mrtx_ctx* ctx, const double center[3], double radius, const float color[3]);
int  mrtx_resize_cubic_f32(mrtx_ctx* ctx, const float* src, int W, int H, int channels, float* dst, int w, int h);
/* rt.set_data/update_data("moon", pos, u, v, r), moon_renderer.py:620-621, 854:
 * u = scene direction of the body +Z (north pole), v = scene direction of lon 0.      */
int  mrtx_set_frame(mrtx_ctx* ctx, const double pos[3], const double u[3], const double v[3],
                    double radius);
/* rt.setup_camera / update_camera (Pinhole), moon_renderer.py:627-635, 568; fov is the
 * vertical field of view in degrees.                                                  */
int  mrtx_set_camera(mrtx_ctx* ctx, const double eye[3], const double target[3],
                     const double up[3], double fov_deg);
/* rt.setup_light / update_light("sun", color=, pos=, radius=), moon_renderer.py:640, 860.
 * radiance = the light "color" (brightness * SUN_BRIGHTNESS_SCALE).                    */
int  mrtx_set_light(mrtx_ctx* ctx, const double pos[3], double radius, double radiance);
/* rt.set_float: scene_epsilon, tonemap_exposure, tonemap_gamma (marching_step,
 * marching_step_eps are accepted and ignored: intersection here is exact).
 * rt.set_uint: path_seg_range (2, 2 + n: n diffuse interreflection bounces after the camera hit), jitter, shadows.
 * Engine switches (not part of the PlotOptiX surface): debug_hits (float64 hit records for
 * tests), start_levels (a, b: pyramid levels the primary / shadow walks start at), long_walk (nodes
 * after which a walk is handed to the referee, default 2048), referee_budget (default 1500), and
 * kernel = 0 float64 one thread per pixel, 1 float64 persistent warps, 2 (default) filtered
 * float32 kernel + float64 referee, 3 the same arithmetic as a wavefront pipeline of dense
 * generation / streaming walk / dense shading kernels over ray records;
 * shadow_queue = 4 (default) kernel 2 as hit queue -> shade_kernel -> shadow queue with the batched
 * pool kernels for primary and shadow rays, 3 / 2 / 1 / 0 its earlier forms (streaming shadow kernel;
 * warp-bound primary kernel; shading in-kernel; everything in one kernel), kept for A/B tests: the
 * same rays and decisions, frames equal bit for bit.                                     */
int  mrtx_set_float(mrtx_ctx* ctx, const char* name, double value);
int  mrtx_set_uint(mrtx_ctx* ctx, const char* name, unsigned a, unsigned b);
/* frame size (TkOptiX(width=, height=)); reallocates accumulation / hit / output.     */
int  mrtx_resize(mrtx_ctx* ctx, int width, int height);

/* ---- render ------------------------------------------------------------------------
 * One launch of the path tracer over the pixel rectangle [x0,x1) x [y0,y1): samples
 * sample0 .. sample0+nsamples-1 of every pixel (primary ray, sun shadow ray, Lambert
 * shading) are ADDED to the float4 accumulation buffer; reset != 0 clears it first.
 * Sample s of pixel p uses a counter-based RNG keyed on (p, s); with jitter off every
 * sample is the pixel centre and the light centre.                                    */
int  mrtx_render(mrtx_ctx* ctx, int x0, int y0, int x1, int y1,
                 unsigned sample0, unsigned nsamples, int reset);
/* Gamma + Overlay post-processing (moon_renderer.py:597-600, renderer_video.py:137-144):
 * rgba8 = overlay over (exposure * accum / weight)^(1/gamma).                          */
int  mrtx_resolve(mrtx_ctx* ctx);
int  mrtx_read_rgba8(mrtx_ctx* ctx, uint8_t* out);          /* [H][W][4]                */
int  mrtx_read_accum_f32(mrtx_ctx* ctx, float* out);        /* [H][W][4] sum r,g,b,count */
int  mrtx_read_hit_f32(mrtx_ctx* ctx, float* out);          /* [H][W][4] x,y,z,dist      */
/* Pipelined frames for the F11 export loop (renderer_video.py:276-364: overlay, update_view,
 * accumulate, grab).  mrtx_frame_submit queues ONE whole frame with the scene as it is now -
 * overlay upload from pinned memory (or none), nsamples samples of every pixel, resolve, RGBA8
 * read-back into the caller's pinned buffer - and returns a ticket (0 / 1) without waiting; the
 * copies run on a second stream against the tracing of the neighbouring frames.  At most two
 * frames are in flight: a third submit reuses the first one's slot.  mrtx_frame_wait blocks
 * until that frame's pixels are in the buffer given at submit.                            */
int  mrtx_frame_submit(mrtx_ctx* ctx, const uint8_t* overlay_rgba_pinned, unsigned nsamples,
                       uint8_t* out_rgba_pinned, int* ticket);
int  mrtx_frame_wait(mrtx_ctx* ctx, int ticket);
/* Frame-parallel time-lapse across GPUs (renderer_video.py:276-364 feeds ONE encoder, in frame order): a rank that is
 * not the consumer queues its frame with mrtx_frame_submit_to - as mrtx_frame_submit, but the resolved RGBA8 frame
 * leaves with ncclSend to rank dst_rank (over NVLink) instead of going to this rank's host memory; mrtx_frame_wait
 * returns when it has left.  The consumer posts one mrtx_frame_recv per such frame, in the order it wants them: the frame
 * is received into a device staging buffer and copied on to the caller's pinned buffer; at most two receives are pending,
 * mrtx_frame_recv_wait blocks until the pixels of that ticket are in host memory.  Needs mrtx_comm_init.           */
int  mrtx_frame_submit_to(mrtx_ctx* ctx, const uint8_t* overlay_rgba_pinned, unsigned nsamples, int dst_rank, int* ticket);
int  mrtx_frame_recv(mrtx_ctx* ctx, int src_rank, uint8_t* out_rgba_pinned, int* ticket);
int  mrtx_frame_recv_wait(mrtx_ctx* ctx, int ticket);
/* The same delivery through peer memory instead of NCCL kernels.  The trace kernels are persistent and own every
 * register of every SM, so an ncclSend / ncclRecv kernel only ever starts at one of their boundaries - on BOTH GPUs at
 * once; at 8 GPUs the consumer cannot take 7 frames per frame time that way.  mrtx_p2p_open allocates this rank's
 * mailbox in its HBM (two frame slots of slot_bytes and two sequence words per sending rank) and returns its CUDA IPC
 * handle as 64 opaque bytes; the host exchanges the handles of all ranks (as it does the NCCL unique id) and gives the
 * nranks x 64 bytes, in rank order, to mrtx_p2p_connect.  From then on mrtx_frame_submit_to copies the frame into the
 * consumer's slot with the copy engine (NVLink), followed by the 4-byte sequence number; mrtx_frame_recv makes the
 * consumer's stream wait for that word (stream memory operation), copy the slot to the pinned buffer and hand the slot
 * back.  No SM takes part.  Same tickets, same ordering rules as above.  Works with or without mrtx_comm_init.       */
int  mrtx_p2p_open(mrtx_ctx* ctx, int nranks, int rank, size_t slot_bytes, uint8_t handle_out[64]);
int  mrtx_p2p_connect(mrtx_ctx* ctx, const uint8_t* handles);
/* unmaps the peers' mailboxes and frees this rank's: frame delivery goes back to NCCL point-to-point (every rank must do the
 * same, with no frame in flight)                                                                                          */
int  mrtx_p2p_close(mrtx_ctx* ctx);
/* rt._get_hit_at(x, y) -> (hx, hy, hz, hd), moon_renderer.py:1138; hd <= 0 = miss.     */
int  mrtx_hit_at(mrtx_ctx* ctx, int x, int y, float out4[4]);
/* device views of the frame buffers (for collectives and zero-copy consumers)        */
int  mrtx_frame_buffers_dev(mrtx_ctx* ctx, void** accum_dev, void** rgba8_dev, void** hit_dev);
/* Per-sample first-hit record of the last mrtx_render with nsamples == 1 and
 * debug_hits on: float64 [H][W][4] = (s_hit, radius, lon, lat) for parity tests.      */
int  mrtx_read_hit_f64(mrtx_ctx* ctx, double* out);

/* counters since the last reset: [0] primary rays, [1] primary rays entering the
 * bounding sphere, [2] primary hits, [3] shadow rays, [4] shadow rays occluded,
 * [5] pyramid node visits, [6] exact patch tests, [7] rays that ran out of steps,
 * warp-phase statistics of the persistent kernel (lanes / executions = SIMD occupancy of a phase):
 * [8] float64 test phases executed, [9] lanes in them, [10] traversal steps executed (per warp),
 * [11] lanes in them, [12] ray-start phases, [13] lanes in them, [14] refills, [15] pixels culled */
int  mrtx_counters(mrtx_ctx* ctx, uint64_t out[16], int reset);
/* Stopwatch of the trace path (engine switch "profile" = 1): CUDA events at the kernel boundaries of every mrtx_render
 * since the last reset, summed: out_ms[0] cull_kernel, [1] beam_kernel, [2] the primary-ray kernel (trace_kernel_pool; trace_kernel_fast with shadow_queue <= 2), [3] shadow_kernel,
 * [4] trace_kernel_referee, [5] fold_kernel (first sample chunk / pixel wave of each launch), [6] launches measured,
 * [7] shade_kernel.
 * bench.py reads the dominant kernel's launch duration from it (roofline.achieved).                              */
int  mrtx_kernel_times(mrtx_ctx* ctx, double out_ms[8], int reset);
/* the filtered kernel's deferrals since the last reset: [0] samples handed to the exact kernel;
 * [r] / [16 + r] primary / shadow rays deferred for reason r: 1 next to the polar axis or map too
 * coarse, 2 ray enters the cell below the surface, 3-4 window start on the surface, 5-6 middle / end
 * of the window on the surface, 7 grazing double root, 8-9 root search left its bracket, 10
 * ill-conditioned root, 11 residual too large, 12-13 root before the cell's longitude / latitude
 * range, 14 walk-back found no crossing, 15 walk longer than 2048 nodes (the referee's warp walks such
 * a ray in pieces, side by side)                                                                  */
int  mrtx_defer_stats(mrtx_ctx* ctx, uint64_t out[32], int reset);

/* ---- multi-GPU (one process per GPU) ----------------------------------------------
 * NCCL is loaded at run time from `libnccl_path` (the torch-bundled libnccl.so.2).     */
int  mrtx_comm_unique_id(const char* libnccl_path, uint8_t id128[128]);
int  mrtx_comm_init(mrtx_ctx* ctx, const char* libnccl_path, int nranks, int rank,
                    const uint8_t id128[128]);
int  mrtx_comm_destroy(mrtx_ctx* ctx);
/* progressive-sample split: sum the float4 accumulation buffers of all ranks          */
int  mrtx_allreduce_accum(mrtx_ctx* ctx);
/* screen-tile split: rank r rendered the rows y with (y / tile_rows) % nranks == r;
 * gather every rank's rows of the resolved RGBA8 frame into every rank's frame.       */
int  mrtx_allgather_rows(mrtx_ctx* ctx, int tile_rows);

/* screen-tile split, interleaved square tiles (BASELINE config 5): mrtx_render_tiles is mrtx_render over the
 * whole frame restricted to the tiles t = ty * tiles_x + tx with t mod nranks == rank (ONE launch: the cull pass drops
 * the other ranks' pixels); mrtx_allgather_tiles tone-maps the owned tiles straight into the send buffer, exchanges them
 * with one ncclAllGather and writes every rank's frame in frame order.  tile = side in pixels, a power of two.       */
int  mrtx_render_tiles(mrtx_ctx* ctx, int tile, unsigned sample0, unsigned nsamples, int reset);
int  mrtx_allgather_tiles(mrtx_ctx* ctx, int tile);

#ifdef __cplusplus
}
#endif
#endif /* MOONB200_H */
