// A/B kernels kept for the tests (mrtx_set_uint("kernel", 0 | 1 | 3)): the float64 one-thread-per-pixel
// kernel, the float64 persistent state machine and the wavefront pipeline.  The production path
// (kernel 2) lives in trace.cu; nothing here is on it.

#include "trace_common.cuh"

namespace {

// ---- reference kernel: one thread per pixel, rays traced to completion one after another -------------
template <bool I16>
__global__ void __launch_bounds__(128)
trace_kernel_simple(const __grid_constant__ RenderArgs A) {
    const int x = A.x0 + blockIdx.x * blockDim.x + threadIdx.x;
    const int y = A.y0 + blockIdx.y * blockDim.y + threadIdx.y;
    Counters cnt = {0u, 0u, 0u};
    RayStats rs = {0u, 0u, 0u, 0u, 0u};
    if (x < A.x1 && y < A.y1) {
        const uint32_t pixel = (uint32_t)y * (uint32_t)A.width + (uint32_t)x;
        float3 acc = make_float3(0.f, 0.f, 0.f);
        for (unsigned sm = A.sample0; sm < A.sample0 + A.nsamples; ++sm) {
            Ray64 R;
            primary_ray(A, x, y, pixel, sm, R);
            ++rs.primary;
            TraceOut h;
            const unsigned nodes_before = cnt.nodes;
            trace_ray<I16>(A.hf, A.sp.radius, R, 0.0, false, A.hf.top - 3, h, cnt);
            if (cnt.nodes != nodes_before) ++rs.inside;
            if (h.hit) {
                ++rs.hits;
                float3 lit;
                Ray64 S;
                if (shade_hit(A, R, h, x, y, pixel, sm, lit, S)) {
                    TraceOut sh;
                    ++rs.shadow;
                    trace_ray<I16>(A.hf, A.sp.radius, S, 0.0, true, 2, sh, cnt);
                    if (sh.hit) { lit = make_float3(0.f, 0.f, 0.f); ++rs.occluded; }
                }
                acc.x += lit.x; acc.y += lit.y; acc.z += lit.z;
            } else {
                write_miss(A, x, y, sm == A.hit_sample);
            }
        }
        float4* ap = A.accum + (size_t)y * A.width + x;
        float4 old = *ap;
        old.x += acc.x; old.y += acc.y; old.z += acc.z; old.w += (float)A.nsamples;
        *ap = old;
    }
    flush_counters(A, rs, cnt, (threadIdx.y * blockDim.x + threadIdx.x) & 31);
}

// ---- production kernel: persistent warps, per-lane ray state machine, dynamic refill ---------------------
// Rays differ wildly in cost (64 % of a whole-disk frame misses the Moon, limb and terminator rays walk
// hundreds of cells), so a pixel->thread mapping leaves most lanes idle.  Here every lane owns one pixel
// at a time and steps a small state machine; idle lanes are refilled from a global pixel counter (one
// atomic per warp and refill), and the warp alternates between phases that all active lanes can share:
//   START (ray generation + sphere clip)  ->  TRAV (float32 pyramid steps, primary and shadow rays alike)
//   ->  CAND (float64 exact patch test [+ shading, shadow-ray set-up])  ->  next sample / next pixel.
enum { M_IDLE = 0, M_START = 1, M_TRAV = 2, M_CAND = 3, M_BEGIN = 4 };
constexpr int TRAV_BURST = 16;
constexpr int CAND_GROUP = 20;     // run the float64 phase once this many lanes wait for it

template <bool I16>
__global__ void __launch_bounds__(128, 3)
trace_kernel_persistent(const __grid_constant__ RenderArgs A) {
    const int lane = threadIdx.x & 31;
    const unsigned n_limb = A.work_counter[4];
    const unsigned total = n_limb + A.work_counter[1];    // list length, written by cull_kernel
    const float Rf = (float)A.sp.radius;

    Counters cnt = {0u, 0u, 0u};
    RayStats rs = {0u, 0u, 0u, 0u, 0u};
    unsigned ph[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};   // lane 0: phase executions / lanes in them
    int mode = M_IDLE;
    bool shadow = false, exhausted = false;
    int x = 0, y = 0;
    uint32_t pixel = 0;
    unsigned sm = 0;
    float3 acc = make_float3(0.f, 0.f, 0.f), lit = make_float3(0.f, 0.f, 0.f);
    Ray64 R;
    TravState st;
    Patch P;
    float sx = 0.f;
    int face = 4;

    auto retire_sample = [&]() {
        // next sample of the same pixel, or write the pixel back and free the lane
        if (++sm < A.sample0 + A.nsamples) mode = M_START;
        else {
            float4* ap = A.accum + (size_t)y * A.width + x;
            float4 old = *ap;
            old.x += acc.x; old.y += acc.y; old.z += acc.z; old.w += (float)A.nsamples;
            *ap = old;
            mode = M_IDLE;
        }
    };

    for (;;) {
        // Phase census.  Whatever phase most lanes are waiting for runs next, so the expensive phases
        // (float64 patch tests) execute with many lanes at once instead of whenever one lane needs them.
        int n_idle = __popc(__ballot_sync(0xffffffffu, mode == M_IDLE));
        int n_start = __popc(__ballot_sync(0xffffffffu, mode == M_START || mode == M_BEGIN));
        int n_trav = __popc(__ballot_sync(0xffffffffu, mode == M_TRAV));
        int n_cand = __popc(__ballot_sync(0xffffffffu, mode == M_CAND));
        if (n_idle == 32 && exhausted) break;

        // ---- refill idle lanes (batched: at least a quarter warp, or nothing else left to run) ---------
        if (!exhausted && n_idle > 0 && (n_idle >= 8 || n_idle + n_start == 32 || n_trav + n_cand == 0)) {
            const unsigned idle = __ballot_sync(0xffffffffu, mode == M_IDLE);
            unsigned base = 0;
            ++ph[6];
            if (lane == 0) base = atomicAdd(A.work_counter, (unsigned)n_idle);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base + (unsigned)n_idle >= total) exhausted = true;
            if (mode == M_IDLE) {
                const unsigned p = base + (unsigned)__popc(idle & ((1u << lane) - 1u));
                if (p < total) {
                    const unsigned packed = list_pixel(A, p, n_limb);
                    x = (int)(packed & 0xffffu); y = (int)(packed >> 16);
                    pixel = (uint32_t)y * (uint32_t)A.width + (uint32_t)x;
                    sm = A.sample0;
                    acc = make_float3(0.f, 0.f, 0.f);
                    mode = M_START;
                }
            }
            n_start = __popc(__ballot_sync(0xffffffffu, mode == M_START || mode == M_BEGIN));
        }

        const bool others_blocked = exhausted || n_idle < 8;      // no refill possible right now
        if (n_start > 0 && (n_start >= 8 || (n_trav == 0 && (n_cand < CAND_GROUP || others_blocked)))) {
            // ---- START: generate the next primary ray, clip it to the bounding sphere ----------------------
            // (also where a freshly shaded hit starts its shadow ray: one trav_begin site)
            ++ph[4]; ph[5] += (unsigned)n_start;
            if (mode == M_START) {
                primary_ray(A, x, y, pixel, sm, R);
                ++rs.primary;
                shadow = false;
            }
            if (mode == M_START || mode == M_BEGIN) {
                if (trav_begin(A.hf, A.sp.radius, R, 0.0, shadow ? 2 : A.hf.top - 3, st)) {
                    mode = M_TRAV;
                    if (!shadow) ++rs.inside;
                } else {
                    if (shadow) { acc.x += lit.x; acc.y += lit.y; acc.z += lit.z; }
                    else write_miss(A, x, y, sm == A.hit_sample);
                    retire_sample();
                }
            }
        } else if (n_cand > 0 && (n_cand >= CAND_GROUP || n_trav == 0)) {
            // ---- CAND: exact patch test; a primary hit is shaded and may spawn its shadow ray -------------
            ++ph[0]; ph[1] += (unsigned)n_cand;
            // One warp-uniform loop: each trip every lane that still needs an evaluation of f takes it
            // at the same instruction, whatever piece / walk-back state it is in.
            ExactState X;
            bool run = false;
            if (mode == M_CAND) run = exact_begin<I16>(A.hf, A.sp.radius, R, st, P, sx, X, cnt);
            else X.found = 0;
            while (__any_sync(0xffffffffu, run)) {
                if (run) run = exact_step<I16>(A.hf, A.sp.radius, R, st, X, cnt);
            }
            if (mode == M_CAND) {
                if (X.found) {
                    if (shadow) { ++rs.occluded; retire_sample(); }
                    else {
                        ++rs.hits;
                        TraceOut h;
                        exact_result(A.hf, X, h);
                        Ray64 S;
                        const bool need_shadow = shade_hit(A, R, h, x, y, pixel, sm, lit, S);
                        if (need_shadow) {
                            ++rs.shadow;
                            R = S;
                            shadow = true;
                            mode = M_BEGIN;
                        } else {
                            acc.x += lit.x; acc.y += lit.y; acc.z += lit.z;
                            retire_sample();
                        }
                    }
                } else {
                    if (trav_advance(A.hf, st, sx, face)) mode = M_TRAV;
                    else {
                        if (shadow) { acc.x += lit.x; acc.y += lit.y; acc.z += lit.z; }
                        else write_miss(A, x, y, sm == A.hit_sample);
                        retire_sample();
                    }
                }
            }
        } else if (n_trav > 0) {
            // ---- TRAV: pyramid steps shared by primary and shadow rays, while they are the majority -------
#pragma unroll 1
            for (int it = 0; it < TRAV_BURST; ++it) {
                ++ph[2]; ph[3] += (unsigned)__popc(__ballot_sync(0xffffffffu, mode == M_TRAV));
                if (mode == M_TRAV) {
                    const int r = trav_step<I16>(A.hf, Rf, st, P, sx, face, cnt);
                    if (r == TR_CANDIDATE) mode = M_CAND;
                    else if (r == TR_END) {
                        // primary: missed the terrain; shadow: the sun is visible
                        if (shadow) { acc.x += lit.x; acc.y += lit.y; acc.z += lit.z; }
                        else write_miss(A, x, y, sm == A.hit_sample);
                        retire_sample();
                    }
                }
                const int nt = __popc(__ballot_sync(0xffffffffu, mode == M_TRAV));
                if (nt == 0 || __popc(__ballot_sync(0xffffffffu, mode == M_CAND)) >= CAND_GROUP) break;
            }
        }
    }
    flush_counters(A, rs, cnt, lane);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 6; ++i) if (ph[i]) atomicAdd(&A.counters[8 + i], (unsigned long long)ph[i]);
    }
    {
        const unsigned r6 = __reduce_add_sync(0xffffffffu, lane == 0 ? ph[6] : 0u), r7 = __reduce_add_sync(0xffffffffu, ph[7]);
        if (lane == 0) { atomicAdd(&A.counters[14], (unsigned long long)r6); atomicAdd(&A.counters[15], (unsigned long long)r7); }
    }
}

// ---- production path: filtered float32 kernel, one lane per (pixel, sample) -----------------------------
// The exact machinery above costs ~2 500 float64 instructions per patch test and ~300 per node, executed by
// ~10 of 32 lanes.  This kernel decides the same rays with trace_fast.cuh (float32 in a cell-local frame that
// ---- wavefront pipeline (production path, kernel 3) ---------------------------------------------------------------------
// trace_kernel_fast keeps a sample in one lane from the camera to the light and a warp busy until the LAST of its 32
// samples is decided: a grazing ray that walks 200 cells keeps 31 finished lanes waiting, and ray generation, patch
// tests and shading run with whatever lanes happen to need them (measured: 16 of 32 lanes active per instruction on
// primary rays, 7 on shadow rays).  Here the work of one wave of (pixel, sample) items is cut where its shape changes:
//   gen_kernel            dense, one item per thread: camera ray, bounding-sphere clip, first cell -> 64-byte ray record
//   trace_kernel_walk     streaming: each lane owns one ray at a time, walks the pyramid and tests candidate patches;
//                         a lane whose ray is decided writes a 32-byte hit record and takes the next ray of the queue.
//                         Nothing but the float32 walk state lives in registers - the float64 ray is read back from
//                         its record for the ~1.1 patch tests a ray needs.
//   shade_kernel          dense: normal, albedo, Lambert term -> the item's radiance slot; the shadow ray of a lit hit
//                         is clipped and appended to the shadow queue as another ray record
//   trace_kernel_walk     the same streaming kernel over the shadow queue: occluded -> zero the item's slot
//   trace_kernel_referee  the few samples the filter could not certify, traced again with the float64 referee
//   reduce_kernel         per pixel: slots summed in sample order, one accumulator update
// A sample's result does not depend on which lane, warp or launch produced it.
struct HitRec { double s; float fc, fr; int r0, c0; int status; unsigned pad; };      // 32 B; status -1: missed the bounding sphere
static_assert(sizeof(HitRec) == 32, "record layout");

// item -> pixel and sample of the wave
struct ItemId { int x, y; uint32_t pixel; unsigned pl, k, sm; };
__device__ __forceinline__ ItemId item_id(const RenderArgs& A, unsigned it, unsigned n_limb) {
    ItemId d;
    d.pl = it / A.nsamples; d.k = it - d.pl * A.nsamples; d.sm = A.sample0 + d.k;
    const unsigned packed = list_pixel(A, A.wave_p0 + d.pl, n_limb);
    d.x = (int)(packed & 0xffffu); d.y = (int)(packed >> 16);
    d.pixel = (uint32_t)d.y * (uint32_t)A.width + (uint32_t)d.x;
    return d;
}

__global__ void __launch_bounds__(256)
gen_kernel(const __grid_constant__ RenderArgs A) {
    const unsigned n_limb = A.work_counter[4];
    const unsigned nkept = n_limb + A.work_counter[1];
    if (A.wave_p0 >= nkept) return;
    const unsigned n_items = min(A.wave_np, nkept - A.wave_p0) * A.nsamples;
    for (unsigned it = blockIdx.x * blockDim.x + threadIdx.x; it < n_items; it += gridDim.x * blockDim.x) {
        const ItemId d = item_id(A, it, n_limb);
        Ray64 R;
        Walk st;
        primary_ray_fast(A, d.x, d.y, d.pixel, d.sm, R);
        const bool alive = walk_begin(A.hf, A.sp.radius, R, 0.0, A.lvl_primary, st);
        store_ray_rec(A.rays + it, R, st, alive);
        if (!alive) A.hits[it].status = -1;
    }
}

#ifndef MRTX_WALK_MINBLOCKS
#define MRTX_WALK_MINBLOCKS 8
#endif
#ifndef MRTX_WALK_CAND
#define MRTX_WALK_CAND 12
#endif
#ifndef MRTX_WALK_REFILL
#define MRTX_WALK_REFILL 4
#endif

enum { LM_EMPTY = 0, LM_WALK = 1, LM_CAND = 2 };

template <bool I16, bool SHADOW>
__global__ void __launch_bounds__(128, MRTX_WALK_MINBLOCKS)
trace_kernel_walk(const __grid_constant__ RenderArgs A) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    unsigned n_items;
    if (SHADOW) n_items = A.work_counter[5];
    else {
        const unsigned nkept = A.work_counter[4] + A.work_counter[1];
        if (A.wave_p0 >= nkept) return;
        n_items = min(A.wave_np, nkept - A.wave_p0) * A.nsamples;
    }
    const RayRec* const recs = SHADOW ? A.srays : A.rays;
    unsigned* const queue = A.work_counter + (SHADOW ? 6 : 2);
    const int L0 = SHADOW ? A.lvl_shadow : A.lvl_primary;
    const float Rf = A.K.R;
    Counters cnt = {0u, 0u, 0u};
    unsigned n_defer = 0, n_occluded = 0;

    int mode = LM_EMPTY, face = 4;
    unsigned ridx = 0;
    Walk st;
    RawPatch P;
    float sx = 0.f;
    bool exhausted = false;

    for (;;) {
        const unsigned m_walk = __ballot_sync(FULL, mode == LM_WALK);
        const unsigned m_cand = __ballot_sync(FULL, mode == LM_CAND);
        const unsigned m_empty = ~(m_walk | m_cand);
        const bool idle = (m_walk | m_cand) == 0u;
        if (!exhausted && (idle || __popc(m_empty) >= MRTX_WALK_REFILL)) {
            // ---- refill: the next rays of the queue, one atomic per warp ---------------------------------------
            const unsigned n = (unsigned)__popc(m_empty);
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(queue, n);
            base = __shfl_sync(FULL, base, 0);
            if (base + n >= n_items) exhausted = true;
            const unsigned idx = base + (unsigned)__popc(m_empty & lt);
            if (mode == LM_EMPTY && idx < n_items) {
                const RayRec* rec = recs + idx;
                const double2 tail = __ldg((const double2*)rec + 3);
                const float smax = __int_as_float(__double2loint(tail.y));
                if (smax >= 0.0f) {
                    const unsigned cell = (unsigned)__double2hiint(tail.y);
                    Ray64 R;
                    load_ray_rec(rec, R);
                    walk_setup(R, tail.x, smax, st);
                    st.L = L0; st.J = (int)(cell >> 16); st.I = (int)(cell & 0xffffu);
                    st.s = 0.0f; st.steps = 0; st.vnext = NAN;
                    ridx = idx;
                    mode = LM_WALK;
                }
            }
            continue;
        }
        if (idle) break;
        bool finished = false;
        int status = FT_MISS;
        FastHit fh;
        if (__popc(m_cand) >= MRTX_WALK_CAND || __popc(m_cand) >= __popc(m_walk)) {
            // ---- patch test ------------------------------------------------------------------------------------
            if (mode == LM_CAND) {
                ++cnt.tests;
                const RayRec* rec = recs + ridx;
                Ray64 R;
                load_ray_rec(rec, R);
                const double s_in = __ldg(&rec->s_in);
                status = fast_test<I16>(A.hf, A.K, R, s_in, 0.0, st.s, sx, st.smax, P, SHADOW, fh);
                if (status == FT_MISS && walk_advance(A.hf, st, sx, face)) mode = LM_WALK;
                else finished = true;
            }
        } else if (mode == LM_WALK) {
            // ---- walk step -------------------------------------------------------------------------------------
            const int r = walk_step<I16>(A.hf, Rf, A.inv_rs, st, P, sx, face, cnt);
            if (r == TR_END) finished = true;
            else if (st.steps > (int)A.sp.long_walk) { finished = true; status = FT_DEFER_R(15); }
            else if (r == TR_CANDIDATE) mode = LM_CAND;
        }
        if (finished) {
            mode = LM_EMPTY;
            if (SHADOW) {
                if (status != FT_MISS) {
                    const unsigned item = __ldg(A.sitem + ridx);
                    float* slot = A.rad + (size_t)item * 3;          // occluded (or undecided: the referee fills it in)
                    slot[0] = 0.f; slot[1] = 0.f; slot[2] = 0.f;
                    if ((status & 3) == FT_HIT) ++n_occluded;
                    else {
                        atomicAdd(&A.defer_stats[16 + (status >> 2)], 1ull);
                        const unsigned pl = item / A.nsamples;
                        A.defer_items[atomicAdd(&A.work_counter[3], 1u)] = make_uint2(A.wave_p0 + pl, item - pl * A.nsamples);
                        ++n_defer;
                    }
                }
            } else {
                HitRec* h = A.hits + ridx;
                if (status == FT_HIT) {
                    ((double2*)h)[0] = make_double2(fh.s, __hiloint2double(__float_as_int(fh.fr), __float_as_int(fh.fc)));
                    ((int4*)h)[1] = make_int4(fh.r0, fh.c0, status, 0);
                } else h->status = status;
            }
        }
    }
    const RayStats rs = {0u, 0u, 0u, 0u, n_occluded};
    flush_counters(A, rs, cnt, lane);
    if (SHADOW) {
        const unsigned nd = __reduce_add_sync(FULL, n_defer);
        if (lane == 0 && nd) {
            // the referee traces these samples from the camera again: take back what the shading pass counted for them
            const unsigned long long neg = 0ull - (unsigned long long)nd;
            atomicAdd(&A.defer_stats[0], (unsigned long long)nd);
            atomicAdd(&A.counters[0], neg); atomicAdd(&A.counters[1], neg); atomicAdd(&A.counters[2], neg); atomicAdd(&A.counters[3], neg);
        }
    }
}

#ifndef MRTX_SHADE_MINBLOCKS
#define MRTX_SHADE_MINBLOCKS 6
#endif
template <bool I16>
__global__ void __launch_bounds__(128, MRTX_SHADE_MINBLOCKS)
shade_kernel(const __grid_constant__ RenderArgs A) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned n_limb = A.work_counter[4];
    const unsigned nkept = n_limb + A.work_counter[1];
    if (A.wave_p0 >= nkept) return;
    const unsigned n_items = min(A.wave_np, nkept - A.wave_p0) * A.nsamples;
    const unsigned n_round = (n_items + 31u) & ~31u;                     // whole warps stay in the loop (ballots)
    RayStats rs = {0u, 0u, 0u, 0u, 0u};
    const Counters cnt = {0u, 0u, 0u};
    unsigned n_defer = 0;
    for (unsigned it = blockIdx.x * blockDim.x + threadIdx.x; it < n_round; it += gridDim.x * blockDim.x) {
        float3 lit = make_float3(0.f, 0.f, 0.f);
        bool spawn = false;
        Ray64 S;
        Walk sw;
        if (it < n_items) {
            const ItemId d = item_id(A, it, n_limb);
            const HitRec* h = A.hits + it;
            const int4 hb = __ldg((const int4*)h + 1);                  // r0, c0, status
            const int status = hb.z;
            if (status >= 0 && (status & 3) == FT_DEFER) {
                atomicAdd(&A.defer_stats[status >> 2], 1ull);
                A.defer_items[atomicAdd(&A.work_counter[3], 1u)] = make_uint2(A.wave_p0 + d.pl, d.k);
                ++n_defer;
            } else {
                ++rs.primary;
                if (status >= 0) ++rs.inside;
                if (status == FT_HIT) {
                    ++rs.hits;
                    const double2 ha = __ldg((const double2*)h);
                    FastHit fh;
                    fh.s = ha.x; fh.fc = __int_as_float(__double2loint(ha.y)); fh.fr = __int_as_float(__double2hiint(ha.y));
                    fh.r0 = hb.x; fh.c0 = hb.y;
                    RawPatch P;
                    load_raw_patch<I16>(A.hf, fh.r0, fh.c0, P);
                    fh.d00 = decode_exact<I16>(A.hf, P.v00); fh.d01 = decode_exact<I16>(A.hf, P.v01);
                    fh.d10 = decode_exact<I16>(A.hf, P.v10); fh.d11 = decode_exact<I16>(A.hf, P.v11);
                    Ray64 R;
                    load_ray_rec(A.rays + it, R);
                    if (shade_fast(A, R, fh, d.x, d.y, d.pixel, d.sm, lit, S)) {
                        ++rs.shadow;
                        spawn = walk_begin(A.hf, A.sp.radius, S, 0.0, A.lvl_shadow, sw);
                    }
                } else write_miss(A, d.x, d.y, d.sm == A.hit_sample);
            }
            float* slot = A.rad + (size_t)it * 3;
            slot[0] = lit.x; slot[1] = lit.y; slot[2] = lit.z;
        }
        // shadow rays of the warp go to the queue together, in lane order
        const unsigned m = __ballot_sync(FULL, spawn);
        if (m) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(&A.work_counter[5], (unsigned)__popc(m));
            base = __shfl_sync(FULL, base, 0);
            if (spawn) {
                const unsigned j = base + (unsigned)__popc(m & ((1u << lane) - 1u));
                store_ray_rec(A.srays + j, S, sw, true);
                A.sitem[j] = it;
            }
        }
    }
    flush_counters(A, rs, cnt, lane);
    const unsigned nd = __reduce_add_sync(FULL, n_defer);
    if (lane == 0 && nd) atomicAdd(&A.defer_stats[0], (unsigned long long)nd);
}

// per pixel of the wave: radiance slots summed in sample order -> accumulator
__global__ void __launch_bounds__(256)
reduce_kernel(const __grid_constant__ RenderArgs A) {
    const unsigned n_limb = A.work_counter[4];
    const unsigned nkept = n_limb + A.work_counter[1];
    if (A.wave_p0 >= nkept) return;
    const unsigned npix = min(A.wave_np, nkept - A.wave_p0), ns = A.nsamples;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
        const float* v = A.rad + (size_t)i * ns * 3;
        float3 acc = make_float3(0.f, 0.f, 0.f);
        for (unsigned k = 0; k < ns; ++k) { acc.x += v[3 * k]; acc.y += v[3 * k + 1]; acc.z += v[3 * k + 2]; }
        const unsigned px = list_pixel(A, A.wave_p0 + i, n_limb);
        float4* ap = A.accum + (size_t)(px >> 16) * A.width + (px & 0xffffu);
        float4 old = *ap;
        old.x += acc.x; old.y += acc.y; old.z += acc.z; old.w += (float)ns;
        *ap = old;
    }
}

}  // namespace

// scratch of the wavefront pipeline: one allocation, carved into the per-item arrays
static int ensure_wave_buffers(mrtx_ctx* ctx, size_t items) {
    if (ctx->wave_items >= items) return MRTX_OK;
    MRTX_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->wave_buf);
    ctx->wave_buf = nullptr; ctx->wave_items = 0;
    const size_t per_item = 2 * sizeof(RayRec) + sizeof(HitRec) + sizeof(uint2) + 3 * sizeof(float) + sizeof(unsigned);
    MRTX_CUDA(cudaMalloc(&ctx->wave_buf, items * per_item));
    ctx->wave_items = items;
    return MRTX_OK;
}


int launch_trace_alt(mrtx_ctx* ctx, int x0, int y0, int x1, int y1, unsigned s0, unsigned ns, unsigned kernel) {
    RenderArgs A;
    fill_render_args(ctx, x0, y0, x1, y1, s0, ns, A);
    const bool i16 = ctx->hf.is_i16 != 0;
    if (kernel == 0) {
        const dim3 block(8, 16);
        const dim3 grid((x1 - x0 + block.x - 1) / block.x, (y1 - y0 + block.y - 1) / block.y);
        if (i16) trace_kernel_simple<true><<<grid, block, 0, ctx->stream>>>(A);
        else     trace_kernel_simple<false><<<grid, block, 0, ctx->stream>>>(A);
        MRTX_CUDA(cudaGetLastError());
        return MRTX_OK;
    }
    MRTX_CUDA(cudaMemsetAsync(A.work_counter, 0, 8 * sizeof(unsigned), ctx->stream));
    A.list_cap = (unsigned)((size_t)ctx->width * ctx->height);
    {
        const unsigned tiles_x = (unsigned)(x1 - x0 + 7) / 8u, tiles_y = (unsigned)(y1 - y0 + 3) / 4u;
        const unsigned total = tiles_x * tiles_y * 32u;
        cull_kernel<<<(total + 255u) / 256u, 256, 0, ctx->stream>>>(A);
    }
    const long long npix = (long long)(x1 - x0) * (y1 - y0);
    if (kernel == 3) {
        // wavefront pipeline: sample chunks of <= 32, waves of <= WAVE_ITEMS items (bounded scratch memory)
        const size_t WAVE_ITEMS = (size_t)1 << 25;
        int rc = ensure_wave_buffers(ctx, WAVE_ITEMS);
        if (rc) return rc;
        {
            char* q = (char*)ctx->wave_buf;                              // largest alignment first
            A.rays = (RayRec*)q; q += WAVE_ITEMS * sizeof(RayRec);
            A.srays = (RayRec*)q; q += WAVE_ITEMS * sizeof(RayRec);
            A.hits = (HitRec*)q; q += WAVE_ITEMS * sizeof(HitRec);
            A.defer_items = (uint2*)q; q += WAVE_ITEMS * sizeof(uint2);
            A.rad = (float*)q; q += WAVE_ITEMS * 3 * sizeof(float);
            A.sitem = (unsigned*)q;
        }
        // the first cell travels in 16 + 16 bits: start no lower than the level whose grid fits
        int lvl_min = 0;
        while ((ctx->hf.W >> lvl_min) > 65536 && lvl_min < ctx->hf.top) ++lvl_min;
        const int top = ctx->hf.top;
        A.lvl_primary = std::min(std::max(top - (int)A.sp.start_primary, lvl_min), top);
        A.lvl_shadow = std::min(std::max((int)A.sp.start_shadow, lvl_min), top);
        void (*k_primary)(const RenderArgs) = i16 ? trace_kernel_walk<true, false> : trace_kernel_walk<false, false>;
        void (*k_shadow)(const RenderArgs) = i16 ? trace_kernel_walk<true, true> : trace_kernel_walk<false, true>;
        void (*k_shade)(const RenderArgs) = i16 ? shade_kernel<true> : shade_kernel<false>;
        void (*k_referee)(const RenderArgs) = i16 ? trace_kernel_referee<true, true> : trace_kernel_referee<false, true>;
        int per_sm = 0, per_sm_s = 0;
        MRTX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_primary, 128, 0));
        MRTX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_s, k_shadow, 128, 0));
        if (per_sm < 1) per_sm = 1;
        if (per_sm_s < 1) per_sm_s = 1;
        for (unsigned done = 0; done < ns; done += 32u) {
            const unsigned n = ns - done < 32u ? ns - done : 32u;
            A.sample0 = s0 + done; A.nsamples = n;
            const unsigned wave_np = (unsigned)(WAVE_ITEMS / n);
            for (long long p0 = 0; p0 < npix; p0 += wave_np) {          // waves past the end of the list return at once
                A.wave_p0 = (unsigned)p0; A.wave_np = wave_np;
                MRTX_CUDA(cudaMemsetAsync(A.work_counter + 2, 0, 2 * sizeof(unsigned), ctx->stream));
                MRTX_CUDA(cudaMemsetAsync(A.work_counter + 5, 0, 2 * sizeof(unsigned), ctx->stream));
                const long long items = (npix - p0 < (long long)wave_np ? npix - p0 : (long long)wave_np) * n;
                const long long warps_needed = (items + 31) / 32;
                long long blocks = (long long)ctx->sm_count * per_sm, sblocks = (long long)ctx->sm_count * per_sm_s;
                if (blocks * 4 > warps_needed) blocks = (warps_needed + 3) / 4;
                if (sblocks * 4 > warps_needed) sblocks = (warps_needed + 3) / 4;
                if (blocks < 1) blocks = 1;
                if (sblocks < 1) sblocks = 1;
                const long long dense_cap = (long long)ctx->sm_count * 16;
                long long gblocks = std::min((items + 255) / 256, dense_cap), hblocks = std::min((items + 127) / 128, dense_cap * 2);
                long long rblocks = std::min((items / n + 255) / 256, dense_cap);
                gblocks = std::max(gblocks, 1ll); hblocks = std::max(hblocks, 1ll); rblocks = std::max(rblocks, 1ll);
                gen_kernel<<<(unsigned)gblocks, 256, 0, ctx->stream>>>(A);
                k_primary<<<(unsigned)blocks, 128, 0, ctx->stream>>>(A);
                k_shade<<<(unsigned)hblocks, 128, 0, ctx->stream>>>(A);
                k_shadow<<<(unsigned)sblocks, 128, 0, ctx->stream>>>(A);
                k_referee<<<ctx->sm_count * 8, 64, 0, ctx->stream>>>(A);
                reduce_kernel<<<(unsigned)rblocks, 256, 0, ctx->stream>>>(A);
            }
        }
    } else if (kernel == 1) {
        int per_sm = 0;
        if (i16) MRTX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trace_kernel_persistent<true>, 128, 0));
        else     MRTX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trace_kernel_persistent<false>, 128, 0));
        if (per_sm < 1) per_sm = 1;
        const long long warps_needed = (npix + 31) / 32;
        long long blocks = (long long)ctx->sm_count * per_sm;
        if (blocks * 4 > warps_needed) blocks = (warps_needed + 3) / 4;     // small rectangles: fewer blocks
        if (blocks < 1) blocks = 1;
        if (i16) trace_kernel_persistent<true><<<(unsigned)blocks, 128, 0, ctx->stream>>>(A);
        else     trace_kernel_persistent<false><<<(unsigned)blocks, 128, 0, ctx->stream>>>(A);
    }
    MRTX_CUDA(cudaGetLastError());
    return MRTX_OK;
}
