// Shared declarations of libmoonb200.so (internal; the public ABI is include/moonb200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>

#include "../../include/moonb200.h"

// ---- error plumbing -------------------------------------------------------------
void mrtx_set_error(const char* fmt, ...);

#define MRTX_CUDA(call)                                                              \
    do {                                                                             \
        cudaError_t e_ = (call);                                                     \
        if (e_ != cudaSuccess) {                                                     \
            mrtx_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),   \
                           __FILE__, __LINE__);                                      \
            return MRTX_ERR_CUDA;                                                    \
        }                                                                            \
    } while (0)

#define MRTX_REQUIRE(cond, ...)                                                      \
    do {                                                                             \
        if (!(cond)) {                                                               \
            mrtx_set_error(__VA_ARGS__);                                             \
            return MRTX_ERR_INVALID;                                                 \
        }                                                                            \
    } while (0)

#define MRTX_CTX(ctx)                                                                \
    do {                                                                             \
        if (!(ctx)) {                                                                \
            mrtx_set_error("null context");                                          \
            return MRTX_ERR_INVALID;                                                 \
        }                                                                            \
        MRTX_CUDA(cudaSetDevice((ctx)->device));                                     \
    } while (0)

// ---- scene state ------------------------------------------------------------------
#define MRTX_PROF_MAX 256
#ifndef MRTX_TILED
#define MRTX_TILED 0            // development switch: pyramid levels also in 8 x 8-cell tiles, walked by the filtered kernels (measured: no gain)
#endif
#define MRTX_P2P_MAX_RANKS 64
#define MRTX_TUBE_TILE_LOG2 4    // screen tiles of the overlay-tube bins: 16 x 16 pixels ...
#define MRTX_TUBE_TILE_CAP 62     // ... listing up to this many segments each (more: every segment is tested)
#define MRTX_PROF_EVENTS 8       // before cull, after cull, beam, trace_kernel_fast, shade_kernel, shadow_kernel, referee, fold
#define MRTX_MAX_LEVELS 20
#define MRTX_DIL_MIN_LEVEL 2     // lowest level that has a dilated copy (beam pre-pass)

// Height field + max pyramid as the kernels see it.  Level 0 (the 2x2-corner max of
// every bilinear patch) is never stored: the patch's own four texels give it.
// Level k >= 1 holds, per cell, the max texel over the (2^k+1)^2 corner footprint of
// the 2^k x 2^k level-0 cells it covers (wrap in longitude, clamp in latitude).
struct HeightField {
    const void* base;           // int16 or float32 [H][W]
    int   is_i16;
    int   W, H;                 // texels
    int   top;                  // highest level used by traversal (0 = plain DDA)
    float scale, radius_scale;  // int16 decode: D = ((c*scale)+1)/radius_scale, each op f32
    const void* level[MRTX_MAX_LEVELS];   // level[k], k = 1..top; same dtype as base
    const void* dil[MRTX_MAX_LEVELS];     // dil[k], k = MRTX_DIL_MIN_LEVEL..top: level k dilated by one cell (pyramid.cu)
    // the same levels as element offsets from lvl_base (all levels live in one allocation): off[k] = level k,
    // off[MRTX_MAX_LEVELS + k] = dil k.  Kernels copy this table to shared memory: a per-lane level index into a
    // kernel-parameter array costs a dozen instructions per access, into shared memory one.
    const void* lvl_base;
    unsigned off[3 * MRTX_MAX_LEVELS];      // [2 * MRTX_MAX_LEVELS + k]: level k again in 8 x 8-cell tiles (one 128-byte line each, int16)
    int   nx[MRTX_MAX_LEVELS], ny[MRTX_MAX_LEVELS];   // cells per level (level 0: W, H-1)
    float dmax, dmin;           // global max / min displacement factor
    // wall tables (one allocation, hf_tables_owned): cell walls are the half-planes of constant
    // longitude through texel-centre column i and the cones of constant latitude through
    // texel-centre row j
    const float2*  lon32;       // [W+1] (cos, sin) of lambda_i = ((i+0.5)/W - 0.5) * 2 pi
    const double2* lon64;
    const float*   lat32;       // [H]   sin(phi_j), phi_j = (0.5 - (j+0.5)/H) * pi
    const double2* lat64;       // [H]   (sin, cos)(phi_j)
    const float2*  latsc32;     // [H]   (sin, cos)(phi_j): the directional walk tests polar walls on the cosine
};

struct Texture8 {
    const uchar4* data;
    int W, H;
};

struct Camera {
    double eye[3], w[3], right[3], up[3];   // orthonormal view basis (scene space)
    double tan_half_fov;                    // vertical
};

struct SceneParams {
    // scene -> body rotation rows (body x = u x v, y = -v, z = u) and sphere centre
    double ex[3], ey[3], ez[3], pos[3];
    double radius;                          // sphere radius (scene units), 10 in MoonRTX
    double light_pos[3], light_radius, light_radiance;
    double scene_epsilon;
    double sun_disk_pos[3], sun_disk_radius;   // the visible Sun disk (moon_renderer.py:643-650): flat-shaded sphere, scene space; radius <= 0: none
    float  sun_disk_color[3];
    float  exposure, inv_gamma;
    unsigned jitter, shadows, debug_hits;
    unsigned start_primary, start_shadow;   // filtered kernel: primary rays start at level top - start_primary, shadow rays at start_shadow
    unsigned long_walk, referee_budget;     // walks longer than long_walk nodes go to the referee; a referee lane spends referee_budget on a piece
    unsigned blocks_per_sm;                 // development: cap on resident blocks per SM of the two walk kernels (0 = what fits)
    unsigned hard_rays;                     // shadow rays too long for one referee warp are finished by the whole grid (default on)
    unsigned n_bounce;                      // diffuse interreflection bounces after the camera hit (path_seg_range max - 2; 0 = direct light)
    unsigned shadow_queue;                  // 4 (default): as 3, shadow rays by shadow_kernel_pool (batches of 32 queue entries + straggler pool);
                                            // 3: as 2, primary rays by trace_kernel_pool (undecided rays parked in a per-warp pool);
                                            // 2: primary hits -> hit queue -> shade_kernel -> shadow queue -> shadow_kernel;
                                            // 1: shading inside trace_kernel_fast, shadow rays through the queue; 0: everything inside trace_kernel_fast
    unsigned ceiling;                       // shadow rays: ceiling test from this level upwards (0 = off)
    unsigned beam, beam_drop;               // beam pre-pass of the filtered kernel (launches of >= 4 samples); samples start beam_drop levels below the beam's
    unsigned kernel;                        // 2 = filtered float32 kernel + exact kernel on what it defers (default),
                                            // 1 = exact persistent kernel only, 0 = exact, one thread per pixel
};

struct mrtx_ctx {
    int device;
    int sm_count, l2_bytes;
    size_t hbm_bytes;
    cudaStream_t own_stream, stream;
    cudaEvent_t ev0, ev1;

    // data_loader scratch
    unsigned* d_max_bits;       // running max of the un-normalised elevation (as uint bits)
    void* flush_buf; size_t flush_bytes;
    float* h_rs; float* h_rs_dev;   // mapped pinned word the normalise kernel writes radius_scale to
    void* stage[2]; size_t stage_bytes; cudaEvent_t stage_ev[2];   // pinned staging of host-buffer calls (chunked copies)

    // scene
    HeightField hf;
    void* hf_owned_base;        // non-null when the context owns the base map
    void* hf_levels_owned;      // one allocation holding all pyramid levels
    void* hf_tables_owned;      // wall tables
    Texture8 tex[3];            // 0 moon_color, 1 frame_overlay, 2 environment (star map: what rays that miss the Moon see)
    void* tex_owned[3];
    Camera cam;
    SceneParams sp;

    // frame
    int width, height;
    float4* accum; uchar4* rgba8; float4* hit; double4* hit64;
    unsigned long long* d_counters;
    unsigned long long* d_defer_stats;   // 32 entries
    unsigned* d_work;           // trace work counter + list length
    unsigned* pixel_list;       // width * height entries
    uint2* defer_list;          // width * height entries: samples the filtered kernel hands to the exact one
    unsigned* defer_mask;       // width * height words: the samples of a pixel that found that list full
    unsigned long long* accfix; // 3 * width * height: order-independent radiance sums of a launch (folded into accum at its end)
    void* sq_buf; size_t sq_cap; // shadow queue (allocated on first use): sq_cap ray records + aux entries ...
    void* hq_buf; size_t hq_cap; // ... and the hit queue in front of it: hq_cap slots
    void* hard_buf;                  // shadow rays trace_kernel_referee hands to referee_hard_kernel (allocated with the context)
    void* pool_buf; size_t pool_bytes;   // trace_kernel_pool's straggler pools
    void* bq_buf[2]; size_t bq_cap;  // bounce-ray queues (interreflection; allocated when path_seg_range asks for bounces)
    double* beam_s;             // width * height entries (by position in the pixel list): where the pixel's samples start ...
    unsigned char* beam_l;      // ... and the level the beam pre-pass stopped at
    // wavefront pipeline scratch (allocated on first use): ray / hit records, radiance slots, shadow queue, deferred items
    void* wave_buf; size_t wave_items;

    // overlay tubes (grid lines, labels, pins: rt.set_graph, renderer_labels.py:263-305, renderer_pins.py:18-55): capsule
    // segments in scene space, flat-shaded, never occluders of the sun.  tube_seg: 3 float4 per segment (a.xyz, r; b.xyz, -;
    // colour.rgb, -); tube_tiles: per 16 x 16-pixel screen tile a count and up to MRTX_TUBE_TILE_CAP segment indices,
    // rebuilt for the camera of every launch (tube_bin_kernel)
    float4* tube_seg; unsigned n_tubes, tube_cap;
    unsigned* tube_tiles; int tube_tx, tube_ty;

    // pipelined frames (mrtx_frame_submit / mrtx_frame_wait): a copy stream moves frame j's overlay in and frame j-1's
    // RGBA8 out while the main stream traces; overlay and output are double-buffered, one set of events per slot
    cudaStream_t copy_stream;
    uchar4* pipe_overlay[2]; uchar4* pipe_rgba8[2];
    cudaEvent_t pipe_ev_upload[2], pipe_ev_resolve[2], pipe_ev_d2h[2];
    int pipe_slot, pipe_w, pipe_h;
    int pipe_busy[2];           // slot queued and not yet waited for
    // frames delivered to another rank (mrtx_frame_submit_to) / received from one (mrtx_frame_recv): NCCL point-to-point
    // on a third stream; two staging buffers and events on the receiving side
    cudaStream_t comm_stream;
    uchar4* recv_buf[2]; cudaEvent_t recv_ev[2]; int recv_slot, recv_busy[2]; size_t recv_bytes;

    // the same delivery through peer memory (mrtx_p2p_open / mrtx_p2p_connect): every rank owns a MAILBOX in its HBM -
    // per sending rank two frame slots and two sequence words - which its peers map with CUDA IPC.  A frame crosses NVLink
    // as a copy-engine copy into the consumer's slot followed by a 4-byte copy of the sequence number; the consumer's
    // stream waits for that word (stream memory operation, no kernel) and copies the slot to pinned host memory, then
    // hands the slot back by writing the sender's "free" word.  No SM is involved on either side: the trace kernels are
    // persistent and own every register of every SM, so an NCCL send / recv kernel only ever starts at their boundaries.
    void* p2p_box; size_t p2p_slot_bytes, p2p_stride; int p2p_on;
    void* p2p_peer[MRTX_P2P_MAX_RANKS];
    unsigned* p2p_seq;          // p2p_seq[i] = i: source of the 4-byte sequence copies
    unsigned p2p_sent[MRTX_P2P_MAX_RANKS], p2p_rcvd[MRTX_P2P_MAX_RANKS];
    void* p2p_wait_fn;          // cuStreamWaitValue32 (driver entry point), p2p_wait_flags: GEQ (| FLUSH where supported)
    unsigned p2p_wait_flags;

    // per-kernel stopwatch of the trace path (mrtx_set_uint("profile", 1), mrtx_kernel_times): events at the kernel
    // boundaries of every mrtx_render, read and summed on request
    int prof_on, prof_n;
    cudaEvent_t* prof_ev;       // MRTX_PROF_MAX launches x MRTX_PROF_EVENTS events
    double prof_ms[8]; unsigned prof_launches;

    // comm
    void* nccl_lib; void* nccl_comm; int nranks, rank;
    void* gather_buf; size_t gather_bytes;
    int tile_log2;              // > 0 while a tile-sharded launch is being issued (mrtx_render_tiles)
};

// ---- kernels' host launchers (one per .cu) ---------------------------------------------
int launch_downscale_i16(mrtx_ctx* ctx, const int16_t* src_dev, int W, int H, int ds, float* out_dev, float* host_rs_dev);
int downscale_begin(mrtx_ctx* ctx);
int downscale_band(mrtx_ctx* ctx, const int16_t* src_dev, int W, int Hb, int ds, float* out_dev);
int downscale_finish(mrtx_ctx* ctx, float* out_dev, size_t n, float* host_rs_dev);
int launch_color_reduce(mrtx_ctx* ctx, const uint8_t* bgr_dev, int W, int H, int k,
                        const uint8_t* lut_dev, uint8_t* out_dev);
int launch_synth_ldem(mrtx_ctx* ctx, int16_t* out_dev, int W, int H, uint32_t seed);
int launch_synth_color(mrtx_ctx* ctx, uint8_t* out_dev, int W, int H, uint32_t seed);
int launch_background_texture(mrtx_ctx* ctx, const float* rgb_dev, int W, int H, float gamma, uint8_t* rgba_dev);
int launch_resize_cubic(mrtx_ctx* ctx, const float* src_dev, int W, int H, int C, float* dst_dev, int w, int h);
int build_pyramid(mrtx_ctx* ctx);
int launch_trace(mrtx_ctx* ctx, int x0, int y0, int x1, int y1, unsigned s0, unsigned ns);
int launch_resolve(mrtx_ctx* ctx);
int prof_mark(mrtx_ctx* ctx, int which);
int comm_send_bytes(mrtx_ctx* ctx, const void* buf_dev, size_t bytes, int peer, cudaStream_t st);
int p2p_send_frame(mrtx_ctx* ctx, const void* frame_dev, size_t bytes, int dst, cudaStream_t st);    // mailbox transport (comm.cu)
int p2p_recv_frame(mrtx_ctx* ctx, void* out_pinned, size_t bytes, int src, cudaStream_t st);
void p2p_release(mrtx_ctx* ctx);
int comm_recv_bytes(mrtx_ctx* ctx, void* buf_dev, size_t bytes, int peer, cudaStream_t st);     // record event `which` of the current launch (no-op unless profiling)
int launch_resolve_to(mrtx_ctx* ctx, const uchar4* overlay_dev, uchar4* out_dev);
void free_heightfield(mrtx_ctx* ctx);
