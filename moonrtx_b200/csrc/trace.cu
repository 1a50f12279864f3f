// K5-K9: primary ray + sun shadow ray through the max-height pyramid, Lambert shading,
// progressive accumulation, tone map + overlay, hit buffer.
//
// Replaces what PlotOptiX does for MoonRTX's scene (moon_renderer.py:570-650): raygen,
// the "DisplacedSurface" intersection program (a fixed-step march, marching_step 5e-3 /
// marching_step_eps 3e-4, :85-101), the diffuse closest-hit with one spherical light
// (:611-617, 640), accumulation (:578) and the Gamma / Overlay post-processing (:597-600,
// renderer_video.py:137-144).  B200 has no RT cores; intersection here is exact:
//
//   * the ray is walked through (lon, lat) cells of the pyramid, top level first.  Cell
//     walls are planes through the polar axis (constant lon) and cones about it (constant
//     lat), so exits are a linear and a quadratic solve - no inverse trig in the loop;
//   * a cell is skipped when the ray stays above its max radius, else descended;
//   * at level 0 (one bilinear patch) the first root of f(s) = |p(s)| - R*D(u(s), v(s)) is
//     found in float64 over the exact cell interval, so the hit does not depend on the
//     float32 traversal that proposed the cell.
//
// Traversal state is float32 re-based at the bounding-sphere entry (H4); every float32
// decision carries a margin so that it can only add candidate cells, never drop one.

#include "trace_common.cuh"

namespace {

#ifndef MRTX_HARD_BLOCKS
#define MRTX_HARD_BLOCKS 1024
#endif
static_assert(sizeof(HardRay) * HARD_MAX_RAYS <= 256 * 2560, "hard_buf is allocated in api.cu");

// ---- beam pre-pass: one thread per listed pixel (trace_fast.cuh, BeamCtl) -------------------------------------------
// The samples of a pixel are ~15 texels apart at 4K on the full-resolution map: each walks its own cells near the
// surface, but above it they all cross the same empty coarse cells.  This pass crosses them ONCE per pixel, with the
// pixel's centre ray against the dilated pyramid, and leaves for every listed pixel the ray parameter before which no
// sample of it can be below the surface and the level it stopped at; trace_kernel_fast starts the samples there.
template <bool I16>
__global__ void __launch_bounds__(256)
beam_kernel(const __grid_constant__ RenderArgs A) {
    const unsigned n_limb = A.work_counter[4];
    const unsigned nkept = n_limb + A.work_counter[1];
    Counters cnt = {0u, 0u, 0u};
    for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < nkept; p += gridDim.x * blockDim.x) {
        const unsigned packed = list_pixel(A, p, n_limb);
        const int x = (int)(packed & 0xffffu), y = (int)(packed >> 16);
        Ray64 C;
        primary_ray_at(A, x, y, 0.5, 0.5, C);
        // half diagonal of a pixel in the image plane at distance 1 (an upper bound of the angle, the plane's points
        // being at distance >= 1 from the eye)
        const double delta = A.cam.tan_half_fov / (double)A.height * 1.4142135623730951;
        double s_start = 0.0;
        int level = A.hf.top;
        const bool alive = beam_walk<I16>(A.hf, A.K, A.sp.radius, C, delta, s_start, level, cnt);
        A.beam_s[p] = alive ? s_start : 1.0e300;
        A.beam_l[p] = (unsigned char)level;
    }
    const unsigned nodes = __reduce_add_sync(0xffffffffu, cnt.nodes);
    if ((threadIdx.x & 31) == 0 && nodes) atomicAdd(&A.counters[5], (unsigned long long)nodes);
}

// ---- production kernels --------------------------------------------------------------------------------------------
// trace_kernel_fast runs the arithmetic of the filtered float32 path (trace_fast.cuh; every candidate patch is re-based in
// float64) laid out for the SIMD width:
//   * a warp owns 32 >> g_log2 neighbouring pixels, 2^g_log2 lanes share one pixel and trace one sample each, so the
//     lanes of a warp walk the same pyramid nodes (coherent loads, similar trip counts: measured 67 % of the lanes of
//     a primary walk phase are busy);
//   * while-while: lanes traverse until each holds a candidate patch (or has left the sphere), then all candidates are
//     tested at one instruction;
//   * warps claim tasks from a counter, so limb warps that walk hundreds of cells do not leave SMs idle at the end.
// Shadow rays are a different population: next to the terminator one lane's sun ray clears the relief after ten nodes
// and its neighbour's skims it for three hundred.  Traced inside the same warp they kept 28 % of the lanes busy
// (measured: 60 % of all walk iterations of a frame for 38 % of its nodes).  With QUEUE the kernel therefore stops at
// the shaded hit: the shadow ray, set up to its first cell, goes to a queue in HBM as a 64-byte record (+ 16 bytes:
// the radiance it carries if the sun is visible, pixel and sample), and shadow_kernel streams the queue with lane
// refill - a lane whose ray is decided takes the next record.  Queue order is push order, i.e. the 32 rays of a warp's
// pixels stay together at first.  The radiance of a visible sample is added to the pixel with 64-bit fixed-point atomics
// (accfix_add): integer sums do not depend on their order, so the frame stays reproducible bit for bit.
// A sample the filter cannot certify (FT_DEFER, primary or shadow ray) contributes nothing here; it goes on the
// deferred list and trace_kernel_referee traces it again from the camera.
#ifndef MRTX_PHASE_STATS
#define MRTX_PHASE_STATS 0
#endif
#ifndef MRTX_DEFER_PER_SAMPLE
#define MRTX_DEFER_PER_SAMPLE 0      // (one list entry per deferred sample instead of per pixel: measured, no difference)
#endif
#ifndef MRTX_FAST_MINBLOCKS
#define MRTX_FAST_MINBLOCKS 8
#endif
// MODE (SceneParams::shadow_queue): 0 both rays and the shading in this kernel; 1 shading here, shadow ray -> shadow queue;
// 2 (production) the kernel ends at the primary hit, which goes to the hit queue: shade_kernel (dense, one thread per hit)
// shades it and pushes the shadow ray.  Cutting there takes the shading, the second first-cell lookup and the record
// store out of this kernel's 70 KB of code and out of its register budget (measured at config 3: 19.0 -> 14.0 ms for
// this kernel, + shade_kernel).
template <bool I16, int MODE>
__global__ void __launch_bounds__(128, MRTX_FAST_MINBLOCKS)
trace_kernel_fast(const __grid_constant__ RenderArgs A) {
    constexpr bool QUEUE = MODE == 1;
    __shared__ unsigned s_off[3 * MRTX_MAX_LEVELS];          // level offsets (HeightField::off) where a per-lane index is cheap
    if (threadIdx.x < 3 * MRTX_MAX_LEVELS) s_off[threadIdx.x] = A.hf.off[threadIdx.x];
    __syncthreads();
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int gl = A.g_log2, g = 1 << gl;
    const int sub = lane & (g - 1), pw = lane >> gl;         // lane within the pixel's group, pixel within the warp
    const unsigned ppw = 32u >> gl;
    const unsigned n_limb = A.work_counter[4];
    const unsigned nkept = n_limb + A.work_counter[1];
    // this launch's part of the pixel list (waves bound the shadow queue)
    const unsigned p_end = min(nkept, A.wave_p0 + A.wave_np);
    if (A.wave_p0 >= p_end) return;
    const unsigned ntasks = (p_end - A.wave_p0 + ppw - 1u) / ppw;
    const float Rf = A.K.R;
    const unsigned rounds = (A.nsamples + (unsigned)g - 1u) >> gl;

    Counters cnt = {0u, 0u, 0u};
    RayStats rs = {0u, 0u, 0u, 0u, 0u};
    unsigned n_defer = 0;

    for (;;) {
        unsigned task = 0;
        if (lane == 0) task = atomicAdd(&A.work_counter[2], 1u);
        task = __shfl_sync(FULL, task, 0);
        if (task >= ntasks) break;
        const unsigned p = A.wave_p0 + task * ppw + (unsigned)pw;
        const bool valid = p < p_end;
        const unsigned packed = valid ? list_pixel(A, p, n_limb) : 0u;
        const int x = (int)(packed & 0xffffu), y = (int)(packed >> 16);
        const uint32_t pixel = (uint32_t)y * (uint32_t)A.width + (uint32_t)x;
        float3 acc = make_float3(0.f, 0.f, 0.f);
        unsigned dmask = 0;                                  // group leader: deferred samples of this pixel
        // where the beam pre-pass lets this pixel's samples start (no pre-pass: at the bounding sphere, level top - start_primary)
        double s_beam = 0.0;
        int lvl_primary = A.hf.top - (int)A.sp.start_primary;
        if (A.beam_s && valid) {
            s_beam = A.beam_s[p];
            if (s_beam > 0.0) lvl_primary = max((int)A.beam_l[p] - A.beam_drop, 0);
        }

#pragma unroll 1
        for (unsigned rd = 0; rd < rounds; ++rd) {
            const unsigned k = rd * (unsigned)g + (unsigned)sub;
            const unsigned sm = A.sample0 + k;
            const bool active = valid && k < A.nsamples;
            Ray64 R;
            Walk st;
            FastHit fh;
            float3 lit = make_float3(0.f, 0.f, 0.f);
            bool defer = false, want = active, hit = false, entered = false, occluded = false, shadowed = false;
#pragma unroll 1
            for (int pass = 0; pass < (MODE == 2 ? 1 : 2); ++pass) {    // 0: primary ray, 1: shadow ray
                bool alive = false;
                if (want) {
                    if (pass == 0) primary_ray_fast(A, x, y, pixel, sm, R);
                    // (ONE call site for both rays: the first-cell lookup is 4 KB of code, and this kernel's instruction
                    //  footprint is what stalls it - profiles/r06: "no instruction" is its largest stall reason)
                    const int wb = walk_begin2(A.hf, A.sp.radius, R, pass ? 0.0 : s_beam,
                                               pass ? (QUEUE ? A.sq_level : (int)A.sp.start_shadow) : lvl_primary, st);
                    alive = wb == 2;
                    if (pass == 0) entered = wb != 0;
                }
                if (QUEUE && pass == 1) {
                    // the shadow ray leaves the kernel here: set up to its first cell, pushed with what it carries
                    const bool push = alive;
                    const unsigned pm = __ballot_sync(FULL, push);
                    if (pm) {
                        unsigned base = 0;
                        if (lane == __ffs(pm) - 1) base = atomicAdd(&A.work_counter[5], (unsigned)__popc(pm));
                        base = __shfl_sync(FULL, base, __ffs(pm) - 1);
                        if (push) {
                            const unsigned j = base + (unsigned)__popc(pm & lt);
                            store_ray_rec(A.sq_rays + j, R, st, true);
                            A.sq_aux[j] = make_uint4(__float_as_uint(lit.x), __float_as_uint(lit.y), __float_as_uint(lit.z), pixel | (k << 27));
                            lit = make_float3(0.f, 0.f, 0.f);        // shadow_kernel adds it if the sun is visible
                        }
                    }
                    break;
                }
                int res = FT_MISS;
                int ceil_next = pass && A.sp.ceiling ? (int)A.sp.ceiling : 0x7fffffff;
                while (__any_sync(FULL, alive)) {
                    RawPatch P;
                    float sx = 0.f;
                    int face = 4;
                    bool cand = false;
#if MRTX_PHASE_STATS
                    const int steps0 = alive ? st.steps : 0;
                    const bool was_alive = alive;
#endif
                    // walk phase, in lockstep: one node per lane and iteration, the vote at the loop head brings the warp
                    // back together after every node (left to itself the compiler lets lanes that took different
                    // branches of a node run on separately: measured 4 lanes per instruction in walk_step)
                    // (a ray that is still walking after long_walk nodes stops here and is deferred below: decided per ray,
                    //  whatever the other lanes do - trace_kernel_pool and shadow_kernel apply the same rule)
                    while (__any_sync(FULL, alive && !cand && st.steps <= (int)A.sp.long_walk)) {
                        if (alive && !cand && st.steps <= (int)A.sp.long_walk) {
                            bool clear = false;
                            if (!QUEUE && st.L >= ceil_next) {
                                ceil_next = st.L + 2;
                                clear = ceiling_clear<I16>(A.hf, A.K, A.inv_rs, st, A.hf.dmin, s_off);
                            }
                            if (clear) alive = false;
                            else {
                                const int r = walk_step<I16>(A.hf, Rf, A.inv_rs, st, P, sx, face, cnt, nullptr, s_off);
                                if (r == TR_END) alive = false;
                                else if (r == TR_CANDIDATE) cand = true;
                            }
                        }
                    }
#if MRTX_PHASE_STATS
                    {
                        const unsigned dn = was_alive ? (unsigned)(st.steps - steps0) : 0u;
                        const unsigned mx = __reduce_max_sync(FULL, dn), sm_ = __reduce_add_sync(FULL, dn);
                        const unsigned nt = __popc(__ballot_sync(FULL, cand)), na = __popc(__ballot_sync(FULL, was_alive));
                        if (lane == 0) {
                            atomicAdd(&A.counters[pass ? 12 : 10], (unsigned long long)mx);
                            atomicAdd(&A.counters[pass ? 13 : 11], (unsigned long long)sm_);
                            if (nt) { atomicAdd(&A.counters[8], 1ull); atomicAdd(&A.counters[9], (unsigned long long)nt); }
                            atomicAdd(&A.counters[14], (unsigned long long)na);     // lanes alive at the start of walk phases ...
                            atomicAdd(&A.defer_stats[31], 1ull);                     // ... and the number of walk phases
                        }
                    }
#endif
                    if ((alive || cand) && st.steps > (int)A.sp.long_walk) {
                        res = FT_DEFER; alive = false; cand = false;
                        atomicAdd(&A.defer_stats[15 + (pass ? 16 : 0)], 1ull);
                    }
                    if (cand) {
                        ++cnt.tests;
                        const int t = fast_test<I16>(A.hf, A.K, R, st.s_in, 0.0, st.s, sx, st.smax, P, pass != 0, fh);
                        if (t == FT_MISS) { if (!walk_advance(A.hf, st, sx, face)) alive = false; }
                        else {
                            res = t & 3; alive = false;
                            if (res == FT_DEFER) atomicAdd(&A.defer_stats[(t >> 2) + (pass ? 16 : 0)], 1ull);
                        }
                    }
                }
                if (res == FT_DEFER) defer = true;
                if (pass == 0) {
                    hit = res == FT_HIT;
                    want = false;
                    if (MODE == 2) {
                        // the hit leaves the kernel here.  A warp always takes 32 slots (lanes without a hit mark theirs empty),
                        // so that shade_kernel's warps see the hits of one warp's pixels together and push their shadow rays
                        // together: shadow_kernel's lanes are refilled from neighbouring queue entries, and rays of the same
                        // pixels walk the same nodes (measured: hits packed densely cost shadow_kernel 1 ms in 12)
                        const unsigned hm = __ballot_sync(FULL, hit);
                        if (hm) {
                            unsigned base = 0;
                            if (lane == 0) base = atomicAdd(&A.work_counter[7], 32u);
                            base = __shfl_sync(FULL, base, 0);
                            uint4* q = (uint4*)(A.hq + base + (unsigned)lane);
                            // (streaming stores: the queue is read once, much later, and must not push the pyramid out of L2)
                            __stcs(q + 1, make_uint4((unsigned)fh.r0, (unsigned)fh.c0, hit ? pixel | (k << 27) : 0xffffffffu, 0u));
                            if (hit) {
                                __stcs(q, make_uint4((unsigned)__double2loint(fh.s), (unsigned)__double2hiint(fh.s), __float_as_uint(fh.fc), __float_as_uint(fh.fr)));
                                __stcs(q + 2, make_uint4(__float_as_uint(fh.d00), __float_as_uint(fh.d01), __float_as_uint(fh.d10), __float_as_uint(fh.d11)));
                            }
                        }
                        break;
                    }
                    if (hit) {
                        Ray64 S;
                        if (tube_tile_count(A, x, y) && tube_nearest(A, x, y, pixel, sm, fh.s, lit)) want = false;   // an overlay tube in front
                        else want = shade_fast(A, R, fh, x, y, pixel, sm, lit, S);
                        shadowed = want;
                        if (want) R = S;
                    }
                    if (!__any_sync(FULL, want)) break;
                } else if (want) {
                    occluded = res == FT_HIT;
                }
            }
            if (active && !defer) {
                ++rs.primary;
                if (entered) ++rs.inside;
                if (hit) {
                    ++rs.hits;
                    if (shadowed) { ++rs.shadow; if (occluded) ++rs.occluded; }
                    if (!occluded) { acc.x += lit.x; acc.y += lit.y; acc.z += lit.z; }
                } else {
                    float3 tc;
                    if (tube_tile_count(A, x, y) && tube_nearest(A, x, y, pixel, sm, 1.0e300, tc)) { acc.x += tc.x; acc.y += tc.y; acc.z += tc.z; }
                    else {
                        write_miss(A, x, y, sm == A.hit_sample);
                        if (sees_background(A)) { const float3 m = miss_radiance_body(A, R); acc.x += m.x; acc.y += m.y; acc.z += m.z; }
                    }
                }
            }
            // a deferred sample -> its own entry of the deferred list (pixel, bit = sample index in this launch): the referee
            // gives every entry a warp, and the samples of a limb pixel - where several defer at once - are each a long chain
#if MRTX_DEFER_PER_SAMPLE
            if (defer) {
                defer_push(A, pixel, k);
                ++n_defer;
            }
#else
            const unsigned dm = __ballot_sync(FULL, defer);
            if (dm) {
                const unsigned gm = g == 32 ? dm : (dm >> (pw << gl)) & ((1u << g) - 1u);
                dmask |= gm << (rd << gl);
                if (defer) ++n_defer;
            }
#endif
        }
        // per-pixel sum over the group's lanes, one read-modify-write per pixel
        for (int o = g >> 1; o > 0; o >>= 1) {
            acc.x += __shfl_xor_sync(FULL, acc.x, o);
            acc.y += __shfl_xor_sync(FULL, acc.y, o);
            acc.z += __shfl_xor_sync(FULL, acc.z, o);
        }
        if (valid && sub == 0) {
            float4* ap = A.accum + (size_t)y * A.width + x;
            float4 old = *ap;
            old.x += acc.x; old.y += acc.y; old.z += acc.z; old.w += (float)A.nsamples;
            *ap = old;
            if (dmask) A.defer_list[atomicAdd(&A.work_counter[3], 1u)] = make_uint2(packed, dmask);
        }
    }
    flush_counters(A, rs, cnt, lane);
    const unsigned nd = __reduce_add_sync(FULL, n_defer);
    if (lane == 0 && nd) atomicAdd(&A.defer_stats[0], (unsigned long long)nd);
}

// ---- trace_kernel_pool: the hit-queue form with a straggler pool ------------------------------------------------------------
// In trace_kernel_fast a warp stays with its 32 rays until the last of them is decided.  Measured at config 3 (phase
// counters): 2.36 walk phases and 2.35 patch-test phases per warp task - the first with ~31 lanes, the others with 5.5: the
// few rays whose first candidate patch was a miss (grazing rays, 11 M of 47.5 M) hold the warp for 26 % of its walk
// iterations and 57 % of its test phases.  Here a warp gives every batch of rays ONE walk phase and ONE test phase; rays
// that are still undecided after it are parked in the warp's own pool (96-byte records in global memory: ray, walk position -
// walk_setup() rebuilds the rest exactly) and the warp goes on to its next task.  Whenever 32 rays are parked they form the
// next batch, with all lanes busy.  A batch of fresh rays reserves its 32 hit-queue slots as trace_kernel_fast does; a parked
// ray remembers its slot and fills it whenever it is decided, so shade_kernel and shadow_kernel see the queue they would
// have seen.  Everything else a finished ray causes leaves through atomics (fixed-point sums, deferred list).
struct PoolRec { double ox, oy, oz, dx, dy, dz, s_in; float smax, s; int L, J, I, steps; unsigned pix_k, slot, pad1, pad2; };
static_assert(sizeof(PoolRec) == 96, "record layout");
constexpr unsigned POOL_CAP = 64;                           // per warp: at most 31 parked + 32 parked again by a pool batch
#ifndef MRTX_POOL_T2
#define MRTX_POOL_T2 MRTX_POOL_T
#endif
#ifndef MRTX_POOL_MINBLOCKS
#define MRTX_POOL_MINBLOCKS MRTX_FAST_MINBLOCKS
#endif
#ifndef MRTX_POOL_CH
#define MRTX_POOL_CH 1
#endif
#ifndef MRTX_POOL_T
#define MRTX_POOL_T 8
#endif
#ifndef MRTX_POOL_CHECK
#define MRTX_POOL_CHECK 0
#endif
#if MRTX_POOL_CHECK
#define POOL_ASSERT(c, ...) do { if (!(c)) { printf("POOL_ASSERT line %d: ", __LINE__); printf(__VA_ARGS__); printf("\n"); } } while (0)
#else
#define POOL_ASSERT(c, ...) do { } while (0)
#endif

template <bool I16>
__global__ void __launch_bounds__(128, MRTX_POOL_MINBLOCKS)
trace_kernel_pool(const __grid_constant__ RenderArgs A) {
    __shared__ unsigned s_off[3 * MRTX_MAX_LEVELS];
    if (threadIdx.x < 3 * MRTX_MAX_LEVELS) s_off[threadIdx.x] = A.hf.off[threadIdx.x];
    __syncthreads();
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int gl = A.g_log2, g = 1 << gl;
    const int sub = lane & (g - 1), pw = lane >> gl;
    const unsigned ppw = 32u >> gl;
    const unsigned n_limb = A.work_counter[4];
    const unsigned nkept = n_limb + A.work_counter[1];
    const unsigned p_end = min(nkept, A.wave_p0 + A.wave_np);
    if (A.wave_p0 >= p_end) return;
    const unsigned ntasks = (p_end - A.wave_p0 + ppw - 1u) / ppw;
    const float Rf = A.K.R;
    const unsigned rounds = (A.nsamples + (unsigned)g - 1u) >> gl;
    PoolRec* const pool = (PoolRec*)A.pool + (size_t)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * POOL_CAP;
    unsigned npool = 0;                                      // warp-uniform
    unsigned task = 0, rd = rounds, claimed = 0;             // fresh batches: (task, round); rd == rounds: claim the next task
    bool exhausted = false;

    Counters cnt = {0u, 0u, 0u};
    RayStats rs = {0u, 0u, 0u, 0u, 0u};
    unsigned n_defer = 0;

    for (;;) {
        // ---- the next batch: 32 parked rays if there are that many, else a fresh round of a task, else what is parked
        bool from_pool = npool >= 32u;
        if (!from_pool && !exhausted && rd >= rounds) {
            // (tasks are claimed MRTX_POOL_CH at a time: a warp's pool then holds rays of neighbouring pixels)
            if (!claimed) {
                if (lane == 0) task = atomicAdd(&A.work_counter[2], (unsigned)MRTX_POOL_CH);
                task = __shfl_sync(FULL, task, 0);
                claimed = (unsigned)MRTX_POOL_CH;
            } else ++task;
            --claimed;
            if (task >= ntasks) exhausted = true; else rd = 0;
        }
        if (!from_pool && exhausted) {
            if (!npool) break;
            from_pool = true;
        }
        Ray64 R;
        Walk st;
        FastHit fh;
        uint32_t pixel = 0;
        unsigned k = 0, slot = 0xffffffffu;
        bool have = false, alive = false, entered = false;
        if (from_pool) {
            const unsigned take = npool < 32u ? npool : 32u;
            npool -= take;
            if ((unsigned)lane < take) {
                const double2* q = (const double2*)(pool + npool + lane);
                const double2 a = q[0], b = q[1], c = q[2], d = q[3];
                const uint4 e = *((const uint4*)q + 4), f = *((const uint4*)q + 5);
                R.ox = a.x; R.oy = a.y; R.oz = b.x; R.dx = b.y; R.dy = c.x; R.dz = c.y;
                R.oo = R.ox * R.ox + R.oy * R.oy + R.oz * R.oz; R.od = R.ox * R.dx + R.oy * R.dy + R.oz * R.dz;
                walk_setup(R, d.x, __int_as_float(__double2loint(d.y)), st);
                st.s = __int_as_float(__double2hiint(d.y));
                st.L = (int)e.x; st.J = (int)e.y; st.I = (int)e.z; st.steps = (int)e.w; st.vnext = NAN;
                pixel = f.x & 0x7ffffffu; k = f.x >> 27; slot = f.y;
                have = true; alive = true; entered = true;
                POOL_ASSERT(slot < A.hq_cap && pixel < (unsigned)(A.width * A.height) && st.L >= 0 && st.L <= A.hf.top && st.J >= 0 && st.I >= 0,
                            "restored slot %u cap %u pixel %u L %d J %d I %d npool %u lane %d", slot, A.hq_cap, pixel, st.L, st.J, st.I, npool, lane);
            }
            __syncwarp();
        } else {
            const unsigned p = A.wave_p0 + task * ppw + (unsigned)pw;
            const bool valid = p < p_end;
            const unsigned packed = valid ? list_pixel(A, p, n_limb) : 0u;
            const int x = (int)(packed & 0xffffu), y = (int)(packed >> 16);
            pixel = (uint32_t)y * (uint32_t)A.width + (uint32_t)x;
            k = rd * (unsigned)g + (unsigned)sub;
            have = valid && k < A.nsamples;
            if (valid && sub == 0 && rd == 0) A.accum[pixel].w += (float)A.nsamples;       // one lane per pixel and launch
            if (have) {
                double s_beam = 0.0;
                int lvl_primary = A.hf.top - (int)A.sp.start_primary;
                if (A.beam_s) {
                    s_beam = A.beam_s[p];
                    if (s_beam > 0.0) lvl_primary = max((int)A.beam_l[p] - A.beam_drop, 0);
                }
                primary_ray_fast(A, x, y, pixel, A.sample0 + k, R);
                const int wb = walk_begin2(A.hf, A.sp.radius, R, s_beam, lvl_primary, st);
                alive = wb == 2;
                entered = wb != 0;
            }
            ++rd;
        }
        const unsigned sm = A.sample0 + k;

        // ---- one walk phase (in lockstep), one test phase
        RawPatch P;
        float sx = 0.f;
        int face = 4, res = FT_MISS;
        bool cand = false;
        // (the walk phase ends when fewer than MRTX_POOL_T lanes are still walking: those rays are parked as they are and
        //  continue in a later batch with full lanes; while the pools are being emptied every ray walks to its end)
        const unsigned walk_min = exhausted ? 1u : (unsigned)(from_pool ? MRTX_POOL_T2 : MRTX_POOL_T);
        const int long_walk = (int)A.sp.long_walk;          // (a ray still walking after that many nodes goes to the referee)
        do {
            if (alive && !cand && st.steps <= long_walk) {
                const int r = walk_step<I16>(A.hf, Rf, A.inv_rs, st, P, sx, face, cnt, nullptr, s_off);
                if (r == TR_END) alive = false;
                else if (r == TR_CANDIDATE) cand = true;
            }
        } while ((unsigned)__popc(__ballot_sync(FULL, alive && !cand && st.steps <= long_walk)) >= walk_min);
        if ((alive || cand) && st.steps > (int)A.sp.long_walk) {
            res = FT_DEFER; alive = false; cand = false;
            atomicAdd(&A.defer_stats[15], 1ull);
        }
        if (cand) {
            ++cnt.tests;
            const int t = fast_test<I16>(A.hf, A.K, R, st.s_in, 0.0, st.s, sx, st.smax, P, false, fh);
            if (t == FT_MISS) { if (!walk_advance(A.hf, st, sx, face)) alive = false; }
            else {
                res = t & 3; alive = false;
                if (res == FT_DEFER) atomicAdd(&A.defer_stats[t >> 2], 1ull);
            }
        }

        const bool done = have && !alive;
        const bool hit = done && res == FT_HIT;
        // ---- a fresh batch reserves its 32 slots if any of its rays has, or may still get, a hit; lanes mark theirs empty
        if (!from_pool) {
            const unsigned need = __ballot_sync(FULL, hit || alive);
            if (need) {
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(&A.work_counter[7], 32u);
                base = __shfl_sync(FULL, base, 0);
                slot = base + (unsigned)lane;
                POOL_ASSERT(slot < A.hq_cap, "fresh slot %u cap %u", slot, A.hq_cap);
                if (!hit) __stcs((uint4*)(A.hq + slot) + 1, make_uint4(0u, 0u, 0xffffffffu, 0u));
            }
        }
        if (hit) {
            POOL_ASSERT(slot < A.hq_cap, "hit slot %u cap %u from_pool %d", slot, A.hq_cap, (int)from_pool);
            uint4* q = (uint4*)(A.hq + slot);
            __stcs(q, make_uint4((unsigned)__double2loint(fh.s), (unsigned)__double2hiint(fh.s), __float_as_uint(fh.fc), __float_as_uint(fh.fr)));
            __stcs(q + 1, make_uint4((unsigned)fh.r0, (unsigned)fh.c0, pixel | (k << 27), 0u));
            __stcs(q + 2, make_uint4(__float_as_uint(fh.d00), __float_as_uint(fh.d01), __float_as_uint(fh.d10), __float_as_uint(fh.d11)));
        }
        __syncwarp();

        // ---- decided rays without a hit: a miss sees what lies behind the Moon, a deferral goes to the referee
        if (done && !hit) {
            if (res == FT_DEFER) {
                defer_push(A, pixel, k);
                ++n_defer;
            } else {
                ++rs.primary;
                if (entered) ++rs.inside;
                const int x = (int)(pixel % (unsigned)A.width), y = (int)(pixel / (unsigned)A.width);
                float3 tc;
                if (tube_tile_count(A, x, y) && tube_nearest(A, x, y, pixel, sm, 1.0e300, tc)) accfix_add(A.accfix, pixel, tc);
                else {
                    write_miss(A, x, y, sm == A.hit_sample);
                    if (sees_background(A)) accfix_add(A.accfix, pixel, miss_radiance_body(A, R));
                }
            }
        } else if (hit) { ++rs.primary; ++rs.inside; ++rs.hits; }
        __syncwarp();

        // ---- undecided rays wait in the pool
        const unsigned pm = __ballot_sync(FULL, alive);
        if (alive) {
            POOL_ASSERT(npool + (unsigned)__popc(pm & lt) < POOL_CAP && slot != 0xffffffffu, "park at %u slot %u", npool + (unsigned)__popc(pm & lt), slot);
            double2* q = (double2*)(pool + npool + (unsigned)__popc(pm & lt));
            q[0] = make_double2(R.ox, R.oy); q[1] = make_double2(R.oz, R.dx); q[2] = make_double2(R.dy, R.dz);
            q[3] = make_double2(st.s_in, __hiloint2double(__float_as_int(st.s), __float_as_int(st.smax)));
            *((uint4*)q + 4) = make_uint4((unsigned)st.L, (unsigned)st.J, (unsigned)st.I, (unsigned)st.steps);
            *((uint4*)q + 5) = make_uint4(pixel | (k << 27), slot, 0u, 0u);
        }
        npool += (unsigned)__popc(pm);
        __syncwarp();
    }
    flush_counters(A, rs, cnt, lane);
    const unsigned nd = __reduce_add_sync(FULL, n_defer);
    if (lane == 0 && nd) atomicAdd(&A.defer_stats[0], (unsigned long long)nd);
}

// ---- shade_kernel: one thread per queued hit ---------------------------------------------------------------------------------
// Dense and coherent (hits of a warp's pixels are neighbours in the queue): the primary ray evaluated again from (pixel,
// sample), normal, albedo, Lambert term, light sample; the shadow ray set up to its first cell and appended to the shadow
// queue with the radiance it carries.  A lit sample that needs no shadow ray (shadows off, or a ray that starts outside the
// bounding sphere) is added to the pixel here.
// Interreflection (path_seg_range, moon_renderer.py:583).  BOUNCE = false: hits of camera rays.  With
// A.n_bounce > 0 every hit also continues the path: a cosine-distributed direction about its normal (so the Lambert term
// and the density cancel and the path's throughput is just multiplied by the albedo), pushed to the bounce queue; bounce_kernel
// traces that queue to its first hits, which come back here with BOUNCE = true: ray and throughput are read from the bounce
// ray's queue entry, the direct light at the new hit is weighted with the throughput and goes through the same shadow queue.
#ifndef MRTX_SHADE_THREADS
#define MRTX_SHADE_THREADS 256
#endif
#ifndef MRTX_SHADE_MINBLOCKS
#define MRTX_SHADE_MINBLOCKS 2
#endif
template <bool I16, bool BOUNCE, bool SPAWN>
__global__ void __launch_bounds__(MRTX_SHADE_THREADS, MRTX_SHADE_MINBLOCKS)
shade_kernel(const __grid_constant__ RenderArgs A) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    // slots of the queue: 32 per warp iteration that had a hit, some empty (or packed: bounce_kernel's)
    const unsigned n_hits = BOUNCE ? A.work_counter[7] : min(A.work_counter[7], A.hq_cap);
    const unsigned n_round = (n_hits + 31u) & ~31u;                     // whole warps stay in the loop (ballots)
    constexpr bool spawn = SPAWN;                                        // (this stage's hits continue their paths)
    RayStats rs = {0u, 0u, 0u, 0u, 0u};
    const Counters cnt = {0u, 0u, 0u};
    for (unsigned it = blockIdx.x * blockDim.x + threadIdx.x; it < n_round; it += gridDim.x * blockDim.x) {
        float3 lit = make_float3(0.f, 0.f, 0.f), thr = make_float3(1.f, 1.f, 1.f);
        bool push = false, bpush = false;
        Ray64 S, B;
        Walk sw, bw;
        uint32_t pixel = 0;
        unsigned k = 0;
        if (it < n_hits) {
            const uint4* q = (const uint4*)(A.hq + it);
            const uint4 b = __ldcs(q + 1);
            if (b.z != 0xffffffffu) {
            const uint4 a = __ldcs(q), c = __ldcs(q + 2);
            FastHit fh;
            fh.s = __hiloint2double((int)a.y, (int)a.x); fh.fc = __uint_as_float(a.z); fh.fr = __uint_as_float(a.w);
            fh.r0 = (int)b.x; fh.c0 = (int)b.y;
            fh.d00 = __uint_as_float(c.x); fh.d01 = __uint_as_float(c.y); fh.d10 = __uint_as_float(c.z); fh.d11 = __uint_as_float(c.w);
            pixel = b.z & 0x7ffffffu; k = b.z >> 27;
            const int x = (int)(pixel % (unsigned)A.width), y = (int)(pixel / (unsigned)A.width);
            const unsigned sm = A.sample0 + k;
            Ray64 R;
            if (BOUNCE) {
                load_ray_rec(A.bq_in_rays + b.w, R);
                R.oo = R.ox * R.ox + R.oy * R.oy + R.oz * R.oz; R.od = R.ox * R.dx + R.oy * R.dy + R.oz * R.dz;
                const uint4 ax = __ldg(A.bq_in_aux + b.w);
                thr = make_float3(__uint_as_float(ax.x), __uint_as_float(ax.y), __uint_as_float(ax.z));
            } else primary_ray_fast(A, x, y, pixel, sm, R);
            ShadeAux aux;
            if (!BOUNCE && tube_tile_count(A, x, y) && tube_nearest(A, x, y, pixel, sm, fh.s, lit)) {
                // an overlay tube in front of the surface: its flat colour is the sample (added below)
            } else {
                const unsigned dim0 = 2u + 5u * (unsigned)A.depth;      // random dimensions of this hit: light 2, bounce 2, roulette 1
                if (shade_fast(A, R, fh, x, y, pixel, sm, lit, S, dim0, !BOUNCE, spawn ? &aux : nullptr)) {
                    ++rs.shadow;
                    push = walk_begin(A.hf, A.sp.radius, S, 0.0, A.sq_level, sw);
                }
                lit.x *= thr.x; lit.y *= thr.y; lit.z *= thr.z;
                // A path that leaves a point deep in the night finds nothing lit (night_sin2(), host side): every further hit
                // of it would have zero direct light.  Exact, not a cut-off: such paths end here.
                bool night = false;
                if (spawn) {
                    const double rp2 = aux.px * aux.px + aux.py * aux.py + aux.pz * aux.pz;
                    const double lx = A.light_b[0] - aux.px, ly = A.light_b[1] - aux.py, lz = A.light_b[2] - aux.pz;
                    const double pl = aux.px * lx + aux.py * ly + aux.pz * lz;
                    night = pl < 0.0 && pl * pl > (double)A.night_sin2 * rp2 * (lx * lx + ly * ly + lz * lz);
                }
                if (spawn && !night) {
                    // Russian roulette: the path goes on with probability p = the albedo's largest component and carries
                    // albedo / p (unbiased; the Moon reflects 3 - 30 %, so one path in three to five is followed: measured at
                    // config 3 with (2, 4), 56 -> 44 ms per frame)
                    const float p = fminf(fmaxf(aux.alb.x, fmaxf(aux.alb.y, aux.alb.z)), 1.0f);
                    const bool go = p > 0.0f && (float)rnd(pixel, sm, dim0 + 4u) < p;
                    const float ip = go ? 1.0f / p : 0.0f;
                    thr.x *= aux.alb.x * ip; thr.y *= aux.alb.y * ip; thr.z *= aux.alb.z * ip;
                    if (go && fmaxf(thr.x, fmaxf(thr.y, thr.z)) > 0.0f) {
                        // cosine-distributed direction about the normal (branchless ONB, Duff et al. 2017)
                        const float u1 = (float)rnd(pixel, sm, dim0 + 2u), u2 = (float)rnd(pixel, sm, dim0 + 3u);
                        const float rr = sqrtf(u1), cz = sqrtf(fmaxf(1.0f - u1, 0.0f));
                        float st, ct;
                        sincospif(2.0f * u2, &st, &ct);
                        const float nx = aux.nx, ny = aux.ny, nz = aux.nz;
                        const float sg = nz >= 0.0f ? 1.0f : -1.0f, aa = -1.0f / (sg + nz), bb = nx * ny * aa;
                        const float b1x = 1.0f + sg * nx * nx * aa, b1y = sg * bb, b1z = -sg * nx;
                        const float b2x = bb, b2y = sg + ny * ny * aa, b2z = -ny;
                        const float lx = rr * ct, ly = rr * st;
                        double dx = (double)(lx * b1x + ly * b2x + cz * nx), dy = (double)(lx * b1y + ly * b2y + cz * ny), dz = (double)(lx * b1z + ly * b2z + cz * nz);
                        const double dn = d_rsqrt(dx * dx + dy * dy + dz * dz);
                        dx *= dn; dy *= dn; dz *= dn;
                        const double eps = A.sp.scene_epsilon;
                        B.ox = fma(eps, (double)nx, aux.px); B.oy = fma(eps, (double)ny, aux.py); B.oz = fma(eps, (double)nz, aux.pz);
                        B.dx = dx; B.dy = dy; B.dz = dz;
                        B.oo = B.ox * B.ox + B.oy * B.oy + B.oz * B.oz;
                        B.od = B.ox * B.dx + B.oy * B.dy + B.oz * B.dz;
                        bpush = walk_begin(A.hf, A.sp.radius, B, 0.0, A.sq_level, bw);
                    }
                }
            }
            }
        }
        const unsigned pm = __ballot_sync(FULL, push);
        if (pm) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(&A.work_counter[5], (unsigned)__popc(pm));
            base = __shfl_sync(FULL, base, 0);
            if (push) {
                const unsigned j = base + (unsigned)__popc(pm & lt);
                store_ray_rec(A.sq_rays + j, S, sw, true);
                A.sq_aux[j] = make_uint4(__float_as_uint(lit.x), __float_as_uint(lit.y), __float_as_uint(lit.z), pixel | (k << 27));
            }
        }
        if (!push) accfix_add(A.accfix, pixel, lit);                     // (adds nothing where lit is zero)
        if (spawn) {
            const unsigned bm = __ballot_sync(FULL, bpush);
            if (bm) {
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(&A.work_counter[10], (unsigned)__popc(bm));
                base = __shfl_sync(FULL, base, 0);
                if (bpush) {
                    const unsigned j = base + (unsigned)__popc(bm & lt);
                    store_ray_rec(A.bq_out_rays + j, B, bw, true);
                    A.bq_out_aux[j] = make_uint4(__float_as_uint(thr.x), __float_as_uint(thr.y), __float_as_uint(thr.z), pixel | (k << 27));
                }
            }
        }
    }
    flush_counters(A, rs, cnt, lane);
}

// the bounce rays just written become the next stage's input: counters [8] <- [10], [9] = [10] = 0; hit and shadow queue emptied
__global__ void queue_flip_kernel(unsigned* wc) {
    wc[8] = wc[10]; wc[9] = 0u; wc[10] = 0u;
    wc[5] = 0u; wc[6] = 0u; wc[7] = 0u; wc[11] = 0u;
}

// ---- shadow queue: streaming walk with lane refill -------------------------------------------------------------------
// Each lane owns one queued shadow ray at a time.  Only the float32 walk state lives in registers; the float64 ray is read
// back from its record for the patch tests (about one per ten rays).  Walk steps and patch tests alternate warp-wide:
// a test phase runs once enough lanes hold a candidate, empty lanes are refilled four at a time with one atomic.
#ifndef MRTX_SQ_MINBLOCKS
#define MRTX_SQ_MINBLOCKS 8
#endif
#ifndef MRTX_SQ_ASCEND
#define MRTX_SQ_ASCEND false
#endif
#ifndef MRTX_SQ_CAND
#define MRTX_SQ_CAND 10
#endif
#ifndef MRTX_SQ_REFILL
#define MRTX_SQ_REFILL 12
#endif
enum { SQ_EMPTY = 0, SQ_WALK = 1, SQ_CAND = 2 };

// CLOSEST (bounce_kernel): the same streaming walk over the bounce queue, but the ray's FIRST crossing is
// wanted (cells are walked front to back, so that is the first patch that reports one): it goes to the hit queue with the
// index of its ray.  A bounce ray the filter cannot certify is dropped (counted in defer_stats[30]): the float64 referee
// re-traces camera samples, and a 2e-5 share of a path's second-order light is far below one 8-bit step.
template <bool I16, bool CLOSEST>
__global__ void __launch_bounds__(128, MRTX_SQ_MINBLOCKS)
shadow_kernel(const __grid_constant__ RenderArgs A) {
    __shared__ unsigned s_off[3 * MRTX_MAX_LEVELS];          // level offsets (HeightField::off) where a per-lane index is cheap
    if (threadIdx.x < 3 * MRTX_MAX_LEVELS) s_off[threadIdx.x] = A.hf.off[threadIdx.x];
    __syncthreads();
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const unsigned n_items = A.work_counter[CLOSEST ? 8 : 5];
    unsigned* const queue = A.work_counter + (CLOSEST ? 9 : 6);
    const RayRec* const q_rays = CLOSEST ? A.bq_in_rays : A.sq_rays;
    const uint4* const q_aux = CLOSEST ? A.bq_in_aux : A.sq_aux;
    const float Rf = A.K.R;
    Counters cnt = {0u, 0u, 0u};
    unsigned n_defer = 0, n_occluded = 0;
    FastHit fh;

    int mode = SQ_EMPTY, face = 4;
    int ceil_next = 0x7fffffff;
    unsigned ridx = 0;
    Walk st;
    RawPatch P;
    float sx = 0.f;
    bool exhausted = false;

    for (;;) {
        const unsigned m_walk = __ballot_sync(FULL, mode == SQ_WALK);
        const unsigned m_cand = __ballot_sync(FULL, mode == SQ_CAND);
        const unsigned m_empty = ~(m_walk | m_cand);
        const bool idle = (m_walk | m_cand) == 0u;
        if (!exhausted && (idle || __popc(m_empty) >= MRTX_SQ_REFILL)) {
            // ---- refill: the next rays of the queue, one atomic per warp
            const unsigned n = (unsigned)__popc(m_empty);
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(queue, n);
            base = __shfl_sync(FULL, base, 0);
            if (base + n >= n_items) exhausted = true;
            const unsigned idx = base + (unsigned)__popc(m_empty & lt);
            if (mode == SQ_EMPTY && idx < n_items) {
                const RayRec* rec = q_rays + idx;
                const double2 tail = __ldg((const double2*)rec + 3);
                const float smax = __int_as_float(__double2loint(tail.y));
                const unsigned cell = (unsigned)__double2hiint(tail.y);
                Ray64 R;
                load_ray_rec(rec, R);
                walk_setup(R, tail.x, smax, st);
                st.L = A.sq_level; st.J = (int)(cell >> 16); st.I = (int)(cell & 0xffffu);
                st.s = 0.0f; st.steps = 0; st.vnext = NAN;
                ridx = idx;
                mode = SQ_WALK;
                ceil_next = A.sp.ceiling ? (int)A.sp.ceiling : 0x7fffffff;
            }
            continue;
        }
        if (idle) break;
        bool finished = false;
        int status = FT_MISS;
        if (__popc(m_cand) >= MRTX_SQ_CAND || __popc(m_cand) >= __popc(m_walk)) {
            // ---- patch test (any crossing occludes)
            if (mode == SQ_CAND) {
                ++cnt.tests;
                const RayRec* rec = q_rays + ridx;
                Ray64 R;
                load_ray_rec(rec, R);
                status = fast_test<I16>(A.hf, A.K, R, st.s_in, 0.0, st.s, sx, st.smax, P, !CLOSEST, fh);
                if (status == FT_MISS && walk_advance(A.hf, st, sx, face)) mode = SQ_WALK;
                else finished = true;
            }
        } else if (mode == SQ_WALK) {
            // ---- walk step
            if (st.L >= ceil_next) {
                ceil_next = st.L + 2;
                if (ceiling_clear<I16>(A.hf, A.K, A.inv_rs, st, A.hf.dmin, s_off)) finished = true;
            }
            if (!finished) {
                const int r = walk_step<I16, false, MRTX_SQ_ASCEND>(A.hf, Rf, A.inv_rs, st, P, sx, face, cnt, nullptr, s_off);
                if (r == TR_END) finished = true;
                else if (st.steps > (int)A.sp.long_walk) { finished = true; status = FT_DEFER_R(15); }
                else if (r == TR_CANDIDATE) mode = SQ_CAND;
            }
        }
        if (CLOSEST) {
            // first hits of bounce rays -> hit queue (one reservation per warp and step)
            const bool hitp = finished && (status & 3) == FT_HIT;
            const unsigned hm = __ballot_sync(FULL, hitp);
            if (hm) {
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(&A.work_counter[7], (unsigned)__popc(hm));
                base = __shfl_sync(FULL, base, 0);
                if (hitp) {
                    const uint4 aux = __ldg(q_aux + ridx);
                    uint4* q = (uint4*)(A.hq + base + (unsigned)__popc(hm & lt));
                    __stcs(q, make_uint4((unsigned)__double2loint(fh.s), (unsigned)__double2hiint(fh.s), __float_as_uint(fh.fc), __float_as_uint(fh.fr)));
                    __stcs(q + 1, make_uint4((unsigned)fh.r0, (unsigned)fh.c0, aux.w, ridx));
                    __stcs(q + 2, make_uint4(__float_as_uint(fh.d00), __float_as_uint(fh.d01), __float_as_uint(fh.d10), __float_as_uint(fh.d11)));
                }
            }
            if (finished) {
                mode = SQ_EMPTY;
                if ((status & 3) == FT_DEFER) { atomicAdd(&A.defer_stats[30], 1ull); }
            }
        } else if (finished) {
            mode = SQ_EMPTY;
            const uint4 aux = __ldg(q_aux + ridx);
            const uint32_t pixel = aux.w & 0x7ffffffu;
            if (status == FT_MISS) {
                accfix_add(A.accfix, pixel, make_float3(__uint_as_float(aux.x), __uint_as_float(aux.y), __uint_as_float(aux.z)));
            } else if ((status & 3) == FT_HIT) ++n_occluded;
            else {
                // undecided: the referee traces the whole sample again from the camera
                atomicAdd(&A.defer_stats[16 + (status >> 2)], 1ull);
                defer_push(A, pixel, aux.w >> 27);
                ++n_defer;
            }
        }
    }
    const RayStats rs = {0u, 0u, 0u, 0u, n_occluded};
    flush_counters(A, rs, cnt, lane);
    const unsigned nd = __reduce_add_sync(FULL, n_defer);
    if (lane == 0 && nd) {
        // the referee counts these samples again: take back what trace_kernel_fast counted for them
        const unsigned long long neg = 0ull - (unsigned long long)nd;
        atomicAdd(&A.defer_stats[0], (unsigned long long)nd);
        atomicAdd(&A.counters[0], neg); atomicAdd(&A.counters[1], neg); atomicAdd(&A.counters[2], neg); atomicAdd(&A.counters[3], neg);
    }
}

// ---- shadow_kernel_pool: the shadow queue in batches, with a straggler pool ----------------------------------------------------
// The streaming kernel above keeps every lane busy by refilling it, which mixes rays at different stages of their walks in
// one warp: walk_step runs at 14.6 of 32 lanes there (ncu, profiles/r08), against 22.9 in trace_kernel_pool, whose batches
// start together.  Queue entries are coherent as they lie - 32 consecutive entries are the shadow rays of one shade warp:
// neighbouring hits, the same Sun.  Here a warp takes 32 consecutive entries, gives them one walk phase in lockstep (ending
// when fewer than MRTX_SPOOL_T lanes still walk) and one patch test, and parks what is still undecided (32-byte records: queue
// index + walk position; walk_setup() rebuilds the rest from the queue entry exactly); 32 parked rays form a batch of their own.
// Any crossing occludes; decisions and sums are the streaming kernel's (same per-ray arithmetic, fixed-point sums).
struct SPoolRec { unsigned ridx; float s; int L, J, I, steps; unsigned pad0, pad1; };
static_assert(sizeof(SPoolRec) == 32, "record layout");
#ifndef MRTX_SPOOL_T
#define MRTX_SPOOL_T 14
#endif
#ifndef MRTX_SPOOL_T2
#define MRTX_SPOOL_T2 MRTX_SPOOL_T                          // ... of a batch of parked rays
#endif
#ifndef MRTX_SPOOL_MINBLOCKS
#define MRTX_SPOOL_MINBLOCKS MRTX_SQ_MINBLOCKS
#endif
#ifndef MRTX_SPOOL_CH
#define MRTX_SPOOL_CH 1
#endif
template <bool I16, bool CLOSEST>
__global__ void __launch_bounds__(128, MRTX_SPOOL_MINBLOCKS)
shadow_kernel_pool(const __grid_constant__ RenderArgs A) {
    __shared__ unsigned s_off[3 * MRTX_MAX_LEVELS];
    if (threadIdx.x < 3 * MRTX_MAX_LEVELS) s_off[threadIdx.x] = A.hf.off[threadIdx.x];
    __syncthreads();
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const unsigned n_items = A.work_counter[CLOSEST ? 8 : 5];
    unsigned* const queue = A.work_counter + (CLOSEST ? 9 : 6);
    const RayRec* const q_rays = CLOSEST ? A.bq_in_rays : A.sq_rays;
    const uint4* const q_aux = CLOSEST ? A.bq_in_aux : A.sq_aux;
    const float Rf = A.K.R;
    const int long_walk = (int)A.sp.long_walk;
    SPoolRec* const pool = (SPoolRec*)A.pool + (size_t)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * POOL_CAP;
    unsigned npool = 0, next = 0, claimed = 0;               // warp-uniform
    bool exhausted = false;
    Counters cnt = {0u, 0u, 0u};
    unsigned n_defer = 0, n_occluded = 0;
    FastHit fh;

    for (;;) {
        bool from_pool = npool >= 32u;
        unsigned base = 0;
        if (!from_pool && !exhausted) {
            // (MRTX_SPOOL_CH batches are claimed at a time: a warp's pool then holds rays of neighbouring hits)
            if (!claimed) {
                if (lane == 0) next = atomicAdd(queue, 32u * (unsigned)MRTX_SPOOL_CH);
                next = __shfl_sync(FULL, next, 0);
                claimed = (unsigned)MRTX_SPOOL_CH;
            }
            base = next; next += 32u; --claimed;
            if (base >= n_items) exhausted = true;
        }
        if (!from_pool && exhausted) {
            if (!npool) break;
            from_pool = true;
        }
        Walk st;
        unsigned ridx = 0;
        bool alive = false;
        if (from_pool) {
            const unsigned take = npool < 32u ? npool : 32u;
            npool -= take;
            if ((unsigned)lane < take) {
                const uint4* q = (const uint4*)(pool + npool + lane);
                const uint4 a = q[0], b = q[1];
                ridx = a.x;
                const RayRec* rec = q_rays + ridx;
                const double2 tail = __ldg((const double2*)rec + 3);
                Ray64 R;
                load_ray_rec(rec, R);
                walk_setup(R, tail.x, __int_as_float(__double2loint(tail.y)), st);
                st.s = __uint_as_float(a.y); st.L = (int)a.z; st.J = (int)a.w; st.I = (int)b.x; st.steps = (int)b.y; st.vnext = NAN;
                alive = true;
            }
            __syncwarp();
        } else {
            const unsigned idx = base + (unsigned)lane;
            if (idx < n_items) {
                const RayRec* rec = q_rays + idx;
                const double2 tail = __ldg((const double2*)rec + 3);
                const unsigned cell = (unsigned)__double2hiint(tail.y);
                Ray64 R;
                load_ray_rec(rec, R);
                walk_setup(R, tail.x, __int_as_float(__double2loint(tail.y)), st);
                st.L = A.sq_level; st.J = (int)(cell >> 16); st.I = (int)(cell & 0xffffu);
                st.s = 0.0f; st.steps = 0; st.vnext = NAN;
                ridx = idx;
                alive = true;
            }
        }
        const bool have = alive;

        RawPatch P;
        float sx = 0.f;
        int face = 4, status = FT_MISS;
        bool cand = false;
        const unsigned walk_min = exhausted ? 1u : (unsigned)(from_pool ? MRTX_SPOOL_T2 : MRTX_SPOOL_T);
        do {
            if (alive && !cand && st.steps <= long_walk) {
                const int r = walk_step<I16, false, MRTX_SQ_ASCEND>(A.hf, Rf, A.inv_rs, st, P, sx, face, cnt, nullptr, s_off);
                if (r == TR_END) alive = false;
                else if (r == TR_CANDIDATE) cand = true;
            }
        } while ((unsigned)__popc(__ballot_sync(FULL, alive && !cand && st.steps <= long_walk)) >= walk_min);
        if ((alive || cand) && st.steps > long_walk) { status = FT_DEFER_R(15); alive = false; cand = false; }
        if (cand) {
            ++cnt.tests;
            Ray64 R;
            load_ray_rec(q_rays + ridx, R);
            status = fast_test<I16>(A.hf, A.K, R, st.s_in, 0.0, st.s, sx, st.smax, P, !CLOSEST, fh);
            if (!(status == FT_MISS && walk_advance(A.hf, st, sx, face))) alive = false;
        }
        if (CLOSEST) {
            // first hits of bounce rays -> hit queue (one reservation per warp and batch); an undecided bounce ray is dropped
            const bool hitp = have && !alive && (status & 3) == FT_HIT;
            const unsigned hm = __ballot_sync(FULL, hitp);
            if (hm) {
                unsigned hbase = 0;
                if (lane == 0) hbase = atomicAdd(&A.work_counter[7], (unsigned)__popc(hm));
                hbase = __shfl_sync(FULL, hbase, 0);
                if (hitp) {
                    const uint4 aux = __ldg(q_aux + ridx);
                    uint4* q = (uint4*)(A.hq + hbase + (unsigned)__popc(hm & lt));
                    __stcs(q, make_uint4((unsigned)__double2loint(fh.s), (unsigned)__double2hiint(fh.s), __float_as_uint(fh.fc), __float_as_uint(fh.fr)));
                    __stcs(q + 1, make_uint4((unsigned)fh.r0, (unsigned)fh.c0, aux.w, ridx));
                    __stcs(q + 2, make_uint4(__float_as_uint(fh.d00), __float_as_uint(fh.d01), __float_as_uint(fh.d10), __float_as_uint(fh.d11)));
                }
            }
            if (have && !alive && (status & 3) == FT_DEFER) atomicAdd(&A.defer_stats[30], 1ull);
        } else if (have && !alive) {
            const uint4 aux = __ldg(q_aux + ridx);
            const uint32_t pixel = aux.w & 0x7ffffffu;
            if (status == FT_MISS) {
                accfix_add(A.accfix, pixel, make_float3(__uint_as_float(aux.x), __uint_as_float(aux.y), __uint_as_float(aux.z)));
            } else if ((status & 3) == FT_HIT) ++n_occluded;
            else {
                // undecided: the referee traces the whole sample again from the camera
                atomicAdd(&A.defer_stats[16 + (status >> 2)], 1ull);
                defer_push(A, pixel, aux.w >> 27);
                ++n_defer;
            }
        }
        __syncwarp();
        const unsigned pm = __ballot_sync(FULL, alive);
        if (alive) {
            uint4* q = (uint4*)(pool + npool + (unsigned)__popc(pm & lt));
            q[0] = make_uint4(ridx, __float_as_uint(st.s), (unsigned)st.L, (unsigned)st.J);
            q[1] = make_uint4((unsigned)st.I, (unsigned)st.steps, 0u, 0u);
        }
        npool += (unsigned)__popc(pm);
        __syncwarp();
    }
    const RayStats rs = {0u, 0u, 0u, 0u, n_occluded};
    flush_counters(A, rs, cnt, lane);
    const unsigned nd = __reduce_add_sync(FULL, n_defer);
    if (lane == 0 && nd) {
        // the referee counts these samples again: take back what the primary-ray kernel counted for them
        const unsigned long long neg = 0ull - (unsigned long long)nd;
        atomicAdd(&A.defer_stats[0], (unsigned long long)nd);
        atomicAdd(&A.counters[0], neg); atomicAdd(&A.counters[1], neg); atomicAdd(&A.counters[2], neg); atomicAdd(&A.counters[3], neg);
    }
}

// fixed-point sums of the launch -> float accumulators (and cleared for the next launch)
__global__ void __launch_bounds__(256)
fold_kernel(float4* __restrict__ accum, unsigned long long* __restrict__ accfix, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned long long* a = accfix + i * 3;
        const unsigned long long r = a[0], g = a[1], b = a[2];
        if (r | g | b) {
            float4 v = accum[i];
            v.x += (float)((double)r * (1.0 / (double)ACCFIX_SCALE));
            v.y += (float)((double)g * (1.0 / (double)ACCFIX_SCALE));
            v.z += (float)((double)b * (1.0 / (double)ACCFIX_SCALE));
            accum[i] = v;
            a[0] = 0ull; a[1] = 0ull; a[2] = 0ull;
        }
    }
}

// K8: Gamma post-process + Overlay alpha blend -> RGBA8
__global__ void resolve_kernel(const float4* __restrict__ accum, const uchar4* __restrict__ overlay,
                               uchar4* __restrict__ out, size_t n, float exposure, float inv_gamma) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float4 a = accum[i];
        const double wgt = a.w > 0.0f ? (double)a.w : 1.0;
        const double ch[3] = {a.x, a.y, a.z};
        unsigned c8[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            double c = (double)exposure * ch[q] / wgt;
            c = c > 0.0 ? pow(c, (double)inv_gamma) : 0.0;
            const double v = floor(c * 255.0 + 0.5);
            c8[q] = (unsigned)(v > 255.0 ? 255.0 : v);
        }
        if (overlay) {
            // exact alpha compositing on the tone-mapped bytes (renderer_video.py:21-25)
            const uchar4 o = overlay[i];
            const unsigned al = o.w, na = 255u - o.w;
            c8[0] = (o.x * al + c8[0] * na + 127u) / 255u;
            c8[1] = (o.y * al + c8[1] * na + 127u) / 255u;
            c8[2] = (o.z * al + c8[2] * na + 127u) / 255u;
        }
        out[i] = make_uchar4((unsigned char)c8[0], (unsigned char)c8[1], (unsigned char)c8[2], 255);
    }
}

}  // namespace

// shadow queue: ray records and their aux entries in one allocation
// hit queue: a warp takes 32 slots per round of 2^g samples, so a sample count that is no power of two leaves part of the
// last round's slots empty
static int ensure_queues(mrtx_ctx* ctx, size_t items, size_t hit_slots, bool bounces) {
    if (bounces && ctx->bq_cap < items) {
        MRTX_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int q = 0; q < 2; ++q) { cudaFree(ctx->bq_buf[q]); ctx->bq_buf[q] = nullptr; }
        ctx->bq_cap = 0;
        for (int q = 0; q < 2; ++q) MRTX_CUDA(cudaMalloc(&ctx->bq_buf[q], items * (sizeof(RayRec) + sizeof(uint4))));
        ctx->bq_cap = items;
    }
    if (ctx->sq_cap < items) {
        MRTX_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->sq_buf);
        ctx->sq_buf = nullptr; ctx->sq_cap = 0;
        MRTX_CUDA(cudaMalloc(&ctx->sq_buf, items * (sizeof(RayRec) + sizeof(uint4))));
        ctx->sq_cap = items;
    }
    if (ctx->hq_cap < hit_slots) {
        MRTX_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->hq_buf);
        ctx->hq_buf = nullptr; ctx->hq_cap = 0;
        MRTX_CUDA(cudaMalloc(&ctx->hq_buf, hit_slots * sizeof(HitQRec)));
        ctx->hq_cap = hit_slots;
    }
    return MRTX_OK;
}

// sin^2 of the Sun's depression below the horizon of the SPHERE beyond which a path with `left` more bounces cannot reach lit
// terrain: alpha = acos(Rmin / Rmax) is how far beyond the terminator terrain still sees the Sun and half of how far apart two
// surface points can see each other, so lit terrain lies within (1 + 2 left) alpha of the path's present hit (+ 1 degree for
// the Sun's disk and its finite distance)
static float night_sin2(const mrtx_ctx* ctx, int left) {
    const double alpha = acos(std::min(1.0, (double)ctx->hf.dmin / (double)ctx->hf.dmax));
    const double sn = sin(std::min((1.0 + 2.0 * left) * alpha + 0.0175, 1.5707));
    return (float)(sn * sn);
}

template <bool I16, int MODE>
static int launch_fast(mrtx_ctx* ctx, RenderArgs& A, long long npix) {
    int per_sm = 0;
    MRTX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trace_kernel_fast<I16, MODE>, 128, 0));
    if (per_sm < 1) per_sm = 1;
    if (ctx->sp.blocks_per_sm && (int)ctx->sp.blocks_per_sm < per_sm) per_sm = (int)ctx->sp.blocks_per_sm;
    const long long warps_needed = ((npix << A.g_log2) + 31) / 32;
    long long blocks = (long long)ctx->sm_count * per_sm;
    if (blocks * 4 > warps_needed) blocks = (warps_needed + 3) / 4;
    if (blocks < 1) blocks = 1;
    trace_kernel_fast<I16, MODE><<<(unsigned)blocks, 128, 0, ctx->stream>>>(A);
    return MRTX_OK;
}

template <bool I16>
static int launch_trace_t(mrtx_ctx* ctx, RenderArgs& A, unsigned s0, unsigned ns) {
    prof_mark(ctx, 0);
    if (ctx->n_tubes) {
        const int tx = (ctx->width + (1 << MRTX_TUBE_TILE_LOG2) - 1) >> MRTX_TUBE_TILE_LOG2, ty = (ctx->height + (1 << MRTX_TUBE_TILE_LOG2) - 1) >> MRTX_TUBE_TILE_LOG2;
        if (tx != ctx->tube_tx || ty != ctx->tube_ty || !ctx->tube_tiles) {
            MRTX_CUDA(cudaStreamSynchronize(ctx->stream));
            cudaFree(ctx->tube_tiles); ctx->tube_tiles = nullptr;
            MRTX_CUDA(cudaMalloc(&ctx->tube_tiles, (size_t)tx * ty * (MRTX_TUBE_TILE_CAP + 2) * sizeof(unsigned)));
            ctx->tube_tx = tx; ctx->tube_ty = ty;
            A.tube_tiles = ctx->tube_tiles; A.tube_tx = tx;
        }
        const size_t words = (size_t)ctx->tube_tx * ctx->tube_ty * (MRTX_TUBE_TILE_CAP + 2);
        MRTX_CUDA(cudaMemsetAsync(ctx->tube_tiles, 0, words * sizeof(unsigned), ctx->stream));
        tube_bin_kernel<<<(ctx->n_tubes + 127) / 128, 128, 0, ctx->stream>>>(ctx->tube_seg, ctx->n_tubes, ctx->tube_tiles, ctx->tube_tx, ctx->tube_ty,
                                                                          ctx->cam, ctx->width, ctx->height);
    }
    int rc = launch_cull(ctx, A);
    if (rc) return rc;
    prof_mark(ctx, 1);
    const long long npix = (long long)(A.x1 - A.x0) * (A.y1 - A.y0);
    if (ctx->sp.beam && ctx->sp.jitter && ns >= 4u && ctx->hf.top >= MRTX_DIL_MIN_LEVEL) {
        A.beam_s = ctx->beam_s; A.beam_l = ctx->beam_l;
        const unsigned blocks = (unsigned)std::min<long long>((npix + 255) / 256, (long long)ctx->sm_count * 16);
        beam_kernel<I16><<<blocks, 256, 0, ctx->stream>>>(A);
    }
    prof_mark(ctx, 2);
    // (the hit-queue form also serves a scene without shadow rays: shade_kernel then adds the radiance itself)
    const bool queue = ctx->sp.shadow_queue >= 2u || (ctx->sp.shadow_queue != 0 && ctx->sp.shadows != 0);
    const size_t SQ_MAX = (size_t)1 << 26;                  // 64 Mi queued rays = 5 GiB; larger launches run in waves of pixels
    int sq_blocks = 0, sq_pool_blocks = 0, n_bounce = 0;
    A.depth = 0; A.n_bounce = 0;
    if (queue) {
        const size_t chunk = ns < 32u ? ns : 32u;
        const size_t items = std::min<size_t>((size_t)npix * chunk, SQ_MAX);
        const bool pow2 = (chunk & (chunk - 1)) == 0 && (ns <= 32u || ns % 32u == 0);
        n_bounce = ctx->sp.shadow_queue >= 2u ? (int)ctx->sp.n_bounce : 0;       // (interreflection runs through the queues)
        rc = ensure_queues(ctx, items, ctx->sp.shadow_queue >= 2u ? items * (pow2 ? 1 : 2) + 2048 : 0, n_bounce > 0);
        if (rc) return rc;
        A.sq_rays = (RayRec*)ctx->sq_buf;
        A.sq_aux = (uint4*)((char*)ctx->sq_buf + ctx->sq_cap * sizeof(RayRec));
        A.hq = (HitQRec*)ctx->hq_buf; A.hq_cap = (unsigned)ctx->hq_cap;
        A.sq_cap = (unsigned)ctx->sq_cap;
        // a ray's first cell travels in 16 + 16 bits: start no lower than the level whose grid fits
        int lvl = (int)ctx->sp.start_shadow;
        while ((ctx->hf.W >> lvl) > 65536 && lvl < ctx->hf.top) ++lvl;
        A.sq_level = std::min(lvl, ctx->hf.top);
        int per_sm = 0;
        MRTX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, shadow_kernel<I16, false>, 128, 0));
        if (ctx->sp.blocks_per_sm && (int)ctx->sp.blocks_per_sm < per_sm) per_sm = (int)ctx->sp.blocks_per_sm;
        sq_blocks = ctx->sm_count * (per_sm < 1 ? 1 : per_sm);
        if (ctx->sp.shadow_queue >= 4u && !ctx->sp.ceiling) {
            MRTX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, shadow_kernel_pool<I16, false>, 128, 0));
            if (ctx->sp.blocks_per_sm && (int)ctx->sp.blocks_per_sm < per_sm) per_sm = (int)ctx->sp.blocks_per_sm;
            sq_pool_blocks = ctx->sm_count * (per_sm < 1 ? 1 : per_sm);
        }
    }
    // (shadow rays: the batched kernel with its straggler pool unless an engine switch asks for the streaming one)
    auto launch_shadow = [&]() {
        if (sq_pool_blocks) shadow_kernel_pool<I16, false><<<sq_pool_blocks, 128, 0, ctx->stream>>>(A);
        else shadow_kernel<I16, false><<<sq_blocks, 128, 0, ctx->stream>>>(A);
    };
    // chunks of <= 32 samples (one mask bit per sample in the deferred list); within a chunk, waves of pixels that the
    // shadow queue can hold; each followed by the referee over whatever was deferred, and the fold of the fixed-point sums
    for (unsigned done = 0; done < ns; done += 32u) {
        const unsigned n = ns - done < 32u ? ns - done : 32u;
        A.sample0 = s0 + done; A.nsamples = n;
        int gl = 0;
        while ((2u << gl) <= n && gl < 5) ++gl;
        A.g_log2 = gl;
        const long long wave_np = queue ? (long long)(ctx->sq_cap / n) : npix;
        for (long long p0 = 0; p0 < npix; p0 += wave_np) {              // waves past the end of the list return at once
            A.wave_p0 = (unsigned)p0; A.wave_np = (unsigned)std::min<long long>(wave_np, npix - p0);
            if (done || p0) {
                MRTX_CUDA(cudaMemsetAsync(A.work_counter + 2, 0, 2 * sizeof(unsigned), ctx->stream));
                MRTX_CUDA(cudaMemsetAsync(A.work_counter + 5, 0, 3 * sizeof(unsigned), ctx->stream));
                MRTX_CUDA(cudaMemsetAsync(A.work_counter + 11, 0, 2 * sizeof(unsigned), ctx->stream));
            }
            const bool first = !done && !p0;                // (the stopwatch brackets the first chunk and wave of a launch)
            const bool hitq = queue && ctx->sp.shadow_queue >= 2u;
            if (hitq && ctx->sp.shadow_queue >= 3u) {
                // the pooled form: one block per resident slot, each warp with its own pool
                int per_sm = 0;
                MRTX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trace_kernel_pool<I16>, 128, 0));
                if (per_sm < 1) per_sm = 1;
                if (ctx->sp.blocks_per_sm && (int)ctx->sp.blocks_per_sm < per_sm) per_sm = (int)ctx->sp.blocks_per_sm;
                long long blocks = (long long)ctx->sm_count * per_sm;
                const long long warps_needed = (((long long)A.wave_np << A.g_log2) + 31) / 32;
                if (blocks * 4 > warps_needed) blocks = (warps_needed + 3) / 4;
                if (blocks < 1) blocks = 1;
                const size_t need = (size_t)ctx->sm_count * 16 * 4 * POOL_CAP * sizeof(PoolRec);
                if (ctx->pool_bytes < need) {
                    MRTX_CUDA(cudaStreamSynchronize(ctx->stream));
                    cudaFree(ctx->pool_buf); ctx->pool_buf = nullptr; ctx->pool_bytes = 0;
                    MRTX_CUDA(cudaMalloc(&ctx->pool_buf, need));
                    ctx->pool_bytes = need;
                }
                A.pool = ctx->pool_buf;
                trace_kernel_pool<I16><<<(unsigned)blocks, 128, 0, ctx->stream>>>(A);
            } else
            rc = hitq ? launch_fast<I16, 2>(ctx, A, A.wave_np) : queue ? launch_fast<I16, 1>(ctx, A, A.wave_np) : launch_fast<I16, 0>(ctx, A, A.wave_np);
            if (rc) return rc;
            if (first) prof_mark(ctx, 3);
            A.depth = 0; A.n_bounce = n_bounce;
            if (n_bounce) {
                A.night_sin2 = night_sin2(ctx, n_bounce);
                A.bq_out_rays = (RayRec*)ctx->bq_buf[0]; A.bq_out_aux = (uint4*)((char*)ctx->bq_buf[0] + ctx->bq_cap * sizeof(RayRec));
                MRTX_CUDA(cudaMemsetAsync(A.work_counter + 8, 0, 3 * sizeof(unsigned), ctx->stream));
            }
            if (hitq) {
                if (n_bounce) shade_kernel<I16, false, true><<<ctx->sm_count * 8 * (256 / MRTX_SHADE_THREADS), MRTX_SHADE_THREADS, 0, ctx->stream>>>(A);
                else shade_kernel<I16, false, false><<<ctx->sm_count * 8 * (256 / MRTX_SHADE_THREADS), MRTX_SHADE_THREADS, 0, ctx->stream>>>(A);
            }
            if (first) prof_mark(ctx, 4);
            if (queue) launch_shadow();
            // interreflection: bounce rays -> first hits -> shading (direct light weighted with the path's throughput, next
            // bounce) -> shadow rays, once per bounce; the two bounce queues swap roles
            for (int d = 1; d <= n_bounce; ++d) {
                queue_flip_kernel<<<1, 1, 0, ctx->stream>>>(A.work_counter);
                A.depth = d;
                A.night_sin2 = night_sin2(ctx, n_bounce - d);
                void* in = ctx->bq_buf[(d - 1) & 1]; void* out = ctx->bq_buf[d & 1];
                A.bq_in_rays = (RayRec*)in; A.bq_in_aux = (uint4*)((char*)in + ctx->bq_cap * sizeof(RayRec));
                A.bq_out_rays = (RayRec*)out; A.bq_out_aux = (uint4*)((char*)out + ctx->bq_cap * sizeof(RayRec));
                // bounce_kernel: first hits -> hit queue
                if (sq_pool_blocks) shadow_kernel_pool<I16, true><<<sq_pool_blocks, 128, 0, ctx->stream>>>(A);
                else shadow_kernel<I16, true><<<sq_blocks, 128, 0, ctx->stream>>>(A);
                if (d < n_bounce) shade_kernel<I16, true, true><<<ctx->sm_count * 8 * (256 / MRTX_SHADE_THREADS), MRTX_SHADE_THREADS, 0, ctx->stream>>>(A);
                else shade_kernel<I16, true, false><<<ctx->sm_count * 8 * (256 / MRTX_SHADE_THREADS), MRTX_SHADE_THREADS, 0, ctx->stream>>>(A);
                launch_shadow();
            }
            A.depth = 0;
            if (first) prof_mark(ctx, 5);
            trace_kernel_referee<I16, false><<<ctx->sm_count * 8, 64, 0, ctx->stream>>>(A);
            if (A.hard) {
                // shadow rays too long for one warp (polar slivers): every one walked by 16 384 threads, a few at a time
                referee_hard_kernel<I16><<<dim3(MRTX_HARD_BLOCKS, 8), 64, 0, ctx->stream>>>(A);
                referee_hard_finish_kernel<<<1, 64, 0, ctx->stream>>>(A);
            }
            if (first) prof_mark(ctx, 6);
        }
    }
    {
        const size_t n = (size_t)ctx->width * ctx->height;
        fold_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ctx->accum, ctx->accfix, n);
    }
    prof_mark(ctx, 7);
    MRTX_CUDA(cudaGetLastError());
    return MRTX_OK;
}

int launch_trace(mrtx_ctx* ctx, int x0, int y0, int x1, int y1, unsigned s0, unsigned ns) {
    unsigned kernel = ctx->sp.kernel;
    if (kernel >= 2 && !make_fast_consts(ctx->hf, ctx->sp.radius).enabled) kernel = 1;   // map too coarse for the filter: everything would defer
    if (kernel != 2) return launch_trace_alt(ctx, x0, y0, x1, y1, s0, ns, kernel);
    RenderArgs A;
    fill_render_args(ctx, x0, y0, x1, y1, s0, ns, A);
    return ctx->hf.is_i16 ? launch_trace_t<true>(ctx, A, s0, ns) : launch_trace_t<false>(ctx, A, s0, ns);
}

int launch_resolve(mrtx_ctx* ctx) {
    return launch_resolve_to(ctx, ctx->tex[1].data, ctx->rgba8);
}

int launch_resolve_to(mrtx_ctx* ctx, const uchar4* overlay, uchar4* out) {
    const size_t n = (size_t)ctx->width * ctx->height;
    resolve_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ctx->accum, overlay, out, n, ctx->sp.exposure, ctx->sp.inv_gamma);
    MRTX_CUDA(cudaGetLastError());
    return MRTX_OK;
}
