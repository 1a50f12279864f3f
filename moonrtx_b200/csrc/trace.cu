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

#include <algorithm>

#include "trace_fast.cuh"

namespace {

using namespace mrtx_core;

// ---- sampling ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hash_u32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ double rnd(uint32_t pixel, uint32_t sample, uint32_t dim) {
    const uint32_t h = hash_u32(pixel ^ hash_u32(sample * 4u + dim + 0x9e3779b9u));
    return (double)(h >> 8) * (1.0 / 16777216.0);
}

__device__ float3 sample_albedo(const Texture8& tex, double lon, double lat) {
    if (!tex.data) return make_float3(1.0f, 1.0f, 1.0f);
    const int w = tex.W, h = tex.H;
    const float u = (float)((lon * (0.5 / PI_D) + 0.5) * w - 0.5), v = (float)((0.5 - lat * (1.0 / PI_D)) * h - 0.5);
    const float fu = floorf(u);
    int c0 = (int)fu;
    const float fc = u - fu;
    c0 = c0 < 0 ? c0 + w : (c0 >= w ? c0 - w : c0);
    const int c1 = c0 + 1 == w ? 0 : c0 + 1;
    const int r0 = min(max((int)floorf(v), 0), h - 2);
    const float fr = fminf(fmaxf(v - (float)r0, 0.0f), 1.0f);
    const uchar4 a = __ldg(tex.data + (size_t)r0 * w + c0), b = __ldg(tex.data + (size_t)r0 * w + c1);
    const uchar4 c = __ldg(tex.data + (size_t)(r0 + 1) * w + c0), d = __ldg(tex.data + (size_t)(r0 + 1) * w + c1);
    const float w00 = (1.0f - fc) * (1.0f - fr), w01 = fc * (1.0f - fr), w10 = (1.0f - fc) * fr, w11 = fc * fr;
    const float s = 1.0f / 255.0f;
    return make_float3((a.x * w00 + b.x * w01 + c.x * w10 + d.x * w11) * s,
                       (a.y * w00 + b.y * w01 + c.y * w10 + d.y * w11) * s,
                       (a.z * w00 + b.z * w01 + c.z * w10 + d.z * w11) * s);
}

struct RenderArgs {
    HeightField hf;
    Texture8 tex;
    Camera cam;
    SceneParams sp;
    int width, height, x0, y0, x1, y1;
    unsigned sample0, nsamples;
    float4* accum; float4* hit; double4* hit64;
    unsigned long long* counters;
    unsigned* work_counter;          // [0] persistent kernel: next unclaimed entry, [1] pixel list length,
                                     // [2] fast kernel: next unclaimed warp task, [3] deferred list length
    unsigned* pixel_list;            // pixels whose rays can touch the bounding sphere (x | y << 16): [0, work_counter[4])
                                     // limb pixels, [list_cap - work_counter[1], list_cap) the others, backwards
    unsigned list_cap;
    unsigned long long* defer_stats; // [reason + 16 * shadow]: why samples were deferred
    uint2* defer_list;               // (pixel, mask of samples sample0 + bit) the fast kernel could not certify
    // wavefront pipeline (kernel 3): one wave = list pixels [wave_p0, wave_p0 + wave_np) x nsamples samples
    unsigned wave_p0, wave_np;
    float* rad;                      // [item][3] radiance of every sample of the wave, item = (p - wave_p0) * nsamples + k
    struct RayRec* rays;             // [item] primary ray records
    struct HitRec* hits;             // [item] what the primary walk decided
    struct RayRec* srays;            // shadow rays spawned by the shading pass (work_counter[5] of them) ...
    unsigned* sitem;                 // ... and the item each belongs to
    uint2* defer_items;              // (list pixel p, sample k) the filter could not certify (work_counter[3] of them)
    int lvl_primary, lvl_shadow;     // pyramid level the walks start at
    FastConsts K;
    float inv_rs;
    int g_log2;                      // fast kernel: 2^g_log2 lanes share one pixel (one sample each per round)
    // eye and light centre in the body frame (host-computed once per launch)
    double eye_b[3], light_b[3];
};

struct RayStats { unsigned primary, inside, hits, shadow, occluded; };

// p-th pixel of the work list (limb pixels first)
__device__ __forceinline__ unsigned list_pixel(const RenderArgs& A, unsigned p, unsigned n_limb) {
    return A.pixel_list[p < n_limb ? p : A.list_cap - 1u - (p - n_limb)];
}

// Primary ray of (pixel x, y; sample sm) in the body frame.
__device__ __forceinline__ void primary_ray(const RenderArgs& A, int x, int y, uint32_t pixel, unsigned sm, Ray64& R) {
    const SceneParams& sp = A.sp;
    const Camera& cam = A.cam;
    const double aspect = (double)A.width / (double)A.height;
    const double jx = sp.jitter ? rnd(pixel, sm, 0) : 0.5, jy = sp.jitter ? rnd(pixel, sm, 1) : 0.5;
    const double sx = ((x + jx) / A.width * 2.0 - 1.0) * cam.tan_half_fov * aspect;
    const double sy = (1.0 - (y + jy) / A.height * 2.0) * cam.tan_half_fov;
    double d[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) d[a] = cam.w[a] + sx * cam.right[a] + sy * cam.up[a];
    const double dn = 1.0 / sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
#pragma unroll
    for (int a = 0; a < 3; ++a) d[a] *= dn;
    R.ox = A.eye_b[0]; R.oy = A.eye_b[1]; R.oz = A.eye_b[2];
    R.dx = sp.ex[0] * d[0] + sp.ex[1] * d[1] + sp.ex[2] * d[2];
    R.dy = sp.ey[0] * d[0] + sp.ey[1] * d[1] + sp.ey[2] * d[2];
    R.dz = sp.ez[0] * d[0] + sp.ez[1] * d[1] + sp.ez[2] * d[2];
    R.oo = R.ox * R.ox + R.oy * R.oy + R.oz * R.oz;
    R.od = R.ox * R.dx + R.oy * R.dy + R.oz * R.dz;
}

// Shade a primary hit: Lambert term towards the (sampled) sun point, albedo lookup, hit buffers.
// Returns the radiance the sample receives if the sun is visible and, when it faces the sun, the
// shadow ray S to decide that.
__device__ __forceinline__ bool shade_hit(const RenderArgs& A, const Ray64& R, const TraceOut& h, int x, int y,
                                          uint32_t pixel, unsigned sm, float3& lit, Ray64& S) {
    const SceneParams& sp = A.sp;
    const HitInfo& hi = h.info;
    const Patch& P = h.patch;
    const double px = R.ox + h.s * R.dx, py = R.oy + h.s * R.dy, pz = R.oz + h.s * R.dz;
    // normal of r(lon, lat) = R * D: n ~ e_r - (r_lon / (r cos lat)) e_lon - (r_lat / r) e_lat
    const double frc = hi.fr < 0.0 ? 0.0 : (hi.fr > 1.0 ? 1.0 : hi.fr);
    const double dD_dfc = ((double)P.d01 - (double)P.d00) * (1.0 - frc) + ((double)P.d11 - (double)P.d10) * frc;
    double dD_dfr = ((double)P.d10 - (double)P.d00) * (1.0 - hi.fc) + ((double)P.d11 - (double)P.d01) * hi.fc;
    if (hi.fr <= 0.0 || hi.fr >= 1.0) dD_dfr = 0.0;
    const double r_lon = sp.radius * dD_dfc * A.hf.W / (2.0 * PI_D);
    const double r_lat = -sp.radius * dD_dfr * A.hf.H / PI_D;
    const double rho = sqrt(px * px + py * py);
    const double cl = rho / hi.r, sl = pz / hi.r;
    const double so = rho > 0.0 ? px / rho : 0.0, co = rho > 0.0 ? -py / rho : 1.0;
    const double clc = cl > 1e-12 ? cl : 1e-12;
    const double a1 = r_lon / (hi.r * clc), a2 = r_lat / hi.r;
    double nx = cl * so - a1 * co - a2 * (-sl * so);
    double ny = -cl * co - a1 * so - a2 * (sl * co);
    double nz = sl - a2 * cl;
    const double nn = 1.0 / sqrt(nx * nx + ny * ny + nz * nz);
    nx *= nn; ny *= nn; nz *= nn;
    // light sample
    const double Lx = A.light_b[0], Ly = A.light_b[1], Lz = A.light_b[2];
    const double tx = Lx - px, ty = Ly - py, tz = Lz - pz;
    const double dist = sqrt(tx * tx + ty * ty + tz * tz);
    double gx = Lx, gy = Ly, gz = Lz;
    if (sp.jitter && sp.light_radius > 0.0) {
        // uniform point on the disk facing the hit (branchless ONB, Duff et al. 2017)
        const double cx = tx / dist, cy = ty / dist, cz = tz / dist;
        const double sg = cz >= 0.0 ? 1.0 : -1.0, a = -1.0 / (sg + cz), b = cx * cy * a;
        const double b1x = 1.0 + sg * cx * cx * a, b1y = sg * b, b1z = -sg * cx;
        const double b2x = b, b2y = sg + cy * cy * a, b2z = -cy;
        const double rr = sp.light_radius * sqrt(rnd(pixel, sm, 2)), th = 2.0 * PI_D * rnd(pixel, sm, 3);
        double st, ct;
        sincospi(2.0 * rnd(pixel, sm, 3), &st, &ct);      // = sincos(th), without the library's huge-argument path
        gx += rr * (ct * b1x + st * b2x); gy += rr * (ct * b1y + st * b2y); gz += rr * (ct * b1z + st * b2z);
    }
    double lx = gx - px, ly = gy - py, lz = gz - pz;
    const double ln = 1.0 / sqrt(lx * lx + ly * ly + lz * lz);
    lx *= ln; ly *= ln; lz *= ln;
    const double cosl = nx * lx + ny * ly + nz * lz;
    if (sm == A.sample0 && A.hit) {
        // scene = pos + R^T p_body
        const float hx = (float)(sp.pos[0] + sp.ex[0] * px + sp.ey[0] * py + sp.ez[0] * pz);
        const float hy = (float)(sp.pos[1] + sp.ex[1] * px + sp.ey[1] * py + sp.ez[1] * pz);
        const float hz = (float)(sp.pos[2] + sp.ex[2] * px + sp.ey[2] * py + sp.ez[2] * pz);
        A.hit[(size_t)y * A.width + x] = make_float4(hx, hy, hz, (float)h.s);
    }
    if (A.hit64) A.hit64[(size_t)y * A.width + x] = make_double4(h.s, hi.r, hi.lon, hi.lat);
    lit = make_float3(0.f, 0.f, 0.f);
    if (!(cosl > 0.0)) return false;
    const float3 alb = sample_albedo(A.tex, hi.lon, hi.lat);
    const double q = sp.light_radius / dist;
    const float E = (float)(sp.light_radiance * q * q * cosl);
    lit = make_float3(alb.x * E, alb.y * E, alb.z * E);
    S.ox = px + sp.scene_epsilon * nx; S.oy = py + sp.scene_epsilon * ny; S.oz = pz + sp.scene_epsilon * nz;
    S.dx = lx; S.dy = ly; S.dz = lz;
    S.oo = S.ox * S.ox + S.oy * S.oy + S.oz * S.oz;
    S.od = S.ox * S.dx + S.oy * S.dy + S.oz * S.dz;
    return sp.shadows != 0;
}

__device__ __forceinline__ void write_miss(const RenderArgs& A, int x, int y, bool first_sample) {
    if (first_sample && A.hit) A.hit[(size_t)y * A.width + x] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (A.hit64) A.hit64[(size_t)y * A.width + x] = make_double4(-1.0, 0.0, 0.0, 0.0);
}

__device__ __forceinline__ void flush_counters(const RenderArgs& A, const RayStats& rs, const Counters& cnt, int lane) {
    const unsigned vals[8] = {rs.primary, rs.inside, rs.hits, rs.shadow, rs.occluded, cnt.nodes, cnt.tests, cnt.overflow};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const unsigned v = __reduce_add_sync(0xffffffffu, vals[i]);
        if (lane == 0 && v) atomicAdd(&A.counters[i], (unsigned long long)v);
    }
}

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
                write_miss(A, x, y, sm == A.sample0);
            }
        }
        float4* ap = A.accum + (size_t)y * A.width + x;
        float4 old = *ap;
        old.x += acc.x; old.y += acc.y; old.z += acc.z; old.w += (float)A.nsamples;
        *ap = old;
    }
    flush_counters(A, rs, cnt, (threadIdx.y * blockDim.x + threadIdx.x) & 31);
}

// ---- pass 1 of the production path: whole-pixel cull + compaction -------------------------------------
// 64 % of a whole-disk frame never touches the Moon.  One thread per pixel (8x4 tiles, so the list
// keeps screen-space coherence) tests the pixel's centre ray against the bounding sphere grown by 1.5
// pixels; pixels that cannot hit are finished here, the rest are appended to the work list that the
// persistent kernel consumes - its lanes then only ever receive pixels with real work.
__global__ void __launch_bounds__(256)
cull_kernel(const __grid_constant__ RenderArgs A) {
    const int rw = A.x1 - A.x0, rh = A.y1 - A.y0;
    const unsigned tiles_x = (unsigned)(rw + 7) / 8u, tiles_y = (unsigned)(rh + 3) / 4u;
    const unsigned total = tiles_x * tiles_y * 32u;
    const unsigned p = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool keep = false, limb = false;
    unsigned culled = 0;
    int x = 0, y = 0;
    if (p < total) {
        const unsigned tile = p >> 5, within = p & 31u;
        x = A.x0 + (int)(tile % tiles_x) * 8 + (int)(within & 7u);
        y = A.y0 + (int)(tile / tiles_x) * 4 + (int)(within >> 3);
        if (x < A.x1 && y < A.y1) {
            const Camera& cam = A.cam;
            const double Rb = A.sp.radius * (double)A.hf.dmax;
            const double eye_dist = sqrt(A.eye_b[0] * A.eye_b[0] + A.eye_b[1] * A.eye_b[1] + A.eye_b[2] * A.eye_b[2]);
            const double cull_r = Rb + eye_dist * 3.0 * cam.tan_half_fov / A.height;
            const double aspect = (double)A.width / (double)A.height;
            const double cx = ((x + 0.5) / A.width * 2.0 - 1.0) * cam.tan_half_fov * aspect;
            const double cy = (1.0 - (y + 0.5) / A.height * 2.0) * cam.tan_half_fov;
            double d[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) d[a] = cam.w[a] + cx * cam.right[a] + cy * cam.up[a];
            const double dn = 1.0 / sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            const double bx = (A.sp.ex[0] * d[0] + A.sp.ex[1] * d[1] + A.sp.ex[2] * d[2]) * dn;
            const double by = (A.sp.ey[0] * d[0] + A.sp.ey[1] * d[1] + A.sp.ey[2] * d[2]) * dn;
            const double bz = (A.sp.ez[0] * d[0] + A.sp.ez[1] * d[1] + A.sp.ez[2] * d[2]) * dn;
            const double od = A.eye_b[0] * bx + A.eye_b[1] * by + A.eye_b[2] * bz;
            const double d2 = eye_dist * eye_dist - od * od;
            if (eye_dist > cull_r && (d2 > cull_r * cull_r || od > 0.0)) {
                culled = 1;                                     // every sample of this pixel misses
                write_miss(A, x, y, true);
                float4* ap = A.accum + (size_t)y * A.width + x;
                float4 old = *ap;
                old.w += (float)A.nsamples;
                *ap = old;
            } else {
                keep = true;
                const double core = A.sp.radius * (double)A.hf.dmin - eye_dist * 3.0 * cam.tan_half_fov / A.height;
                limb = !(core > 0.0 && d2 < core * core && od < 0.0);
            }
        }
    }
    // Rays that can pass through the relief shell without meeting the sphere below it walk hundreds to thousands of
    // cells (most of all over the poles, where equirectangular cells are slivers): those pixels go to the FRONT of
    // the list so that their long dependent walks start first and hide behind the bulk of the frame; the rest
    // is appended from the far end downwards.
    const unsigned ml = __ballot_sync(0xffffffffu, keep && limb), mi = __ballot_sync(0xffffffffu, keep && !limb);
    unsigned bl = 0, bi = 0;
    if (lane == 0) {
        if (ml) bl = atomicAdd(&A.work_counter[4], (unsigned)__popc(ml));
        if (mi) bi = atomicAdd(&A.work_counter[1], (unsigned)__popc(mi));
    }
    bl = __shfl_sync(0xffffffffu, bl, 0); bi = __shfl_sync(0xffffffffu, bi, 0);
    const unsigned below = (1u << lane) - 1u;
    if (keep) {
        const unsigned packed = (unsigned)x | ((unsigned)y << 16);
        if (limb) A.pixel_list[bl + (unsigned)__popc(ml & below)] = packed;
        else A.pixel_list[A.list_cap - 1u - (bi + (unsigned)__popc(mi & below))] = packed;
    }
    const unsigned nc = __reduce_add_sync(0xffffffffu, culled);
    if (lane == 0 && nc) {
        atomicAdd(&A.counters[0], (unsigned long long)nc * A.nsamples);
        atomicAdd(&A.counters[15], (unsigned long long)nc);
    }
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
                    else write_miss(A, x, y, sm == A.sample0);
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
                        else write_miss(A, x, y, sm == A.sample0);
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
                        else write_miss(A, x, y, sm == A.sample0);
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
// is re-based in float64 per candidate) in one tenth of the instructions, and is laid out for the SIMD
// width instead of around it:
//   * a warp owns 32 >> g_log2 neighbouring pixels, 2^g_log2 lanes share one pixel and trace one sample each,
//     so the lanes of a warp walk the same pyramid nodes (coherent loads, similar trip counts);
//   * while-while: lanes traverse until each holds a candidate patch (or has left the sphere), then all
//     candidates are tested at one instruction; primary and shadow rays run through the same loop body;
//   * per-pixel sums are reduced with shuffles: one accumulator read-modify-write per pixel;
//   * warps claim tasks from a counter, so limb / terminator warps that walk hundreds of cells do not leave
//     SMs idle at the end of the frame.
// A sample the filter cannot certify (FT_DEFER) contributes nothing here; its bit is set in the pixel's entry of
// the deferred list and trace_kernel_referee traces it again afterwards.
__device__ __forceinline__ void primary_ray_fast(const RenderArgs& A, int x, int y, uint32_t pixel, unsigned sm, Ray64& R) {
    const SceneParams& sp = A.sp;
    const Camera& cam = A.cam;
    const double aspect = (double)A.width / (double)A.height;
    const double jx = sp.jitter ? rnd(pixel, sm, 0) : 0.5, jy = sp.jitter ? rnd(pixel, sm, 1) : 0.5;
    const double sx = ((x + jx) / A.width * 2.0 - 1.0) * cam.tan_half_fov * aspect;
    const double sy = (1.0 - (y + jy) / A.height * 2.0) * cam.tan_half_fov;
    double d[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) d[a] = cam.w[a] + sx * cam.right[a] + sy * cam.up[a];
    const double dn = d_rsqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
#pragma unroll
    for (int a = 0; a < 3; ++a) d[a] *= dn;
    R.ox = A.eye_b[0]; R.oy = A.eye_b[1]; R.oz = A.eye_b[2];
    R.dx = sp.ex[0] * d[0] + sp.ex[1] * d[1] + sp.ex[2] * d[2];
    R.dy = sp.ey[0] * d[0] + sp.ey[1] * d[1] + sp.ey[2] * d[2];
    R.dz = sp.ez[0] * d[0] + sp.ez[1] * d[1] + sp.ez[2] * d[2];
    R.oo = R.ox * R.ox + R.oy * R.oy + R.oz * R.oz;
    R.od = R.ox * R.dx + R.oy * R.dy + R.oz * R.dz;
}

// hit64 debug record (tests): everything from the float64 hit point
__device__ __noinline__ void write_hit64(const RenderArgs& A, const Ray64& R, const FastHit& h, int x, int y) {
    const double px = R.ox + h.s * R.dx, py = R.oy + h.s * R.dy, pz = R.oz + h.s * R.dz;
    const double lon = ((h.c0 + 0.5 + (double)h.fc) / A.hf.W - 0.5) * (2.0 * PI_D);
    const double lat = (0.5 - (h.r0 + 0.5 + (double)h.fr) / A.hf.H) * PI_D;
    A.hit64[(size_t)y * A.width + x] = make_double4(h.s, sqrt(px * px + py * py + pz * pz), lon, lat);
}

// Lambert term, albedo and shadow ray of a primary hit; float32 except where positions near R are
// added or subtracted.  Same model as shade_hit().
__device__ __forceinline__ bool shade_fast(const RenderArgs& A, const Ray64& R, const FastHit& h, int x, int y,
                                           uint32_t pixel, unsigned sm, float3& lit, Ray64& S) {
    const SceneParams& sp = A.sp;
    const double px = fma(h.s, R.dx, R.ox), py = fma(h.s, R.dy, R.oy), pz = fma(h.s, R.dz, R.oz);
    const float fx = (float)px, fy = (float)py, fz = (float)pz;
    // normal of r(lon, lat) = R * D: n ~ e_r - (r_lon / (r cos lat)) e_lon - (r_lat / r) e_lat
    const float dD_dfc = fmaf(h.fr, (h.d11 - h.d10) - (h.d01 - h.d00), h.d01 - h.d00);
    float dD_dfr = fmaf(h.fc, (h.d11 - h.d01) - (h.d10 - h.d00), h.d10 - h.d00);
    if ((h.r0 == 0 && h.fr <= 0.0f) || (h.r0 == A.hf.H - 2 && h.fr >= 1.0f)) dD_dfr = 0.0f;      // polar cap: the rows clamp
    const float Rf = A.K.R;
    const float r_lon = Rf * dD_dfc * A.K.Kw, r_lat = -Rf * dD_dfr * A.K.Kh;
    const float rho2 = fmaf(fx, fx, fy * fy);
    const float irho = f_rsqrt(rho2), ir = f_rsqrt(fmaf(fz, fz, rho2));
    const float rho = rho2 * irho;
    const float cl = rho * ir, sl = fz * ir, so = fx * irho, co = -fy * irho;
    const float a1 = r_lon * irho, a2 = r_lat * ir;              // r_lon / (r cos lat), r_lat / r
    float nx = cl * so - a1 * co + a2 * sl * so;
    float ny = -cl * co - a1 * so - a2 * sl * co;
    float nz = sl - a2 * cl;
    const float nn = f_rsqrt(nx * nx + ny * ny + nz * nz);
    nx *= nn; ny *= nn; nz *= nn;
    // light sample
    double tx = A.light_b[0] - px, ty = A.light_b[1] - py, tz = A.light_b[2] - pz;
    const double idist = d_rsqrt(tx * tx + ty * ty + tz * tz);
    if (sp.jitter && sp.light_radius > 0.0) {
        // uniform point on the disk facing the hit (branchless ONB, Duff et al. 2017)
        const float cx = (float)(tx * idist), cy = (float)(ty * idist), cz = (float)(tz * idist);
        const float sg = cz >= 0.0f ? 1.0f : -1.0f, a = -1.0f / (sg + cz), b = cx * cy * a;
        const float b1x = 1.0f + sg * cx * cx * a, b1y = sg * b, b1z = -sg * cx;
        const float b2x = b, b2y = sg + cy * cy * a, b2z = -cy;
        const float rr = (float)sp.light_radius * sqrtf((float)rnd(pixel, sm, 2));
        float st, ct;
        sincospif(2.0f * (float)rnd(pixel, sm, 3), &st, &ct);
        tx += (double)(rr * (ct * b1x + st * b2x)); ty += (double)(rr * (ct * b1y + st * b2y)); tz += (double)(rr * (ct * b1z + st * b2z));
    }
    const double ln = d_rsqrt(tx * tx + ty * ty + tz * tz);
    const double lx = tx * ln, ly = ty * ln, lz = tz * ln;
    const float cosl = nx * (float)lx + ny * (float)ly + nz * (float)lz;
    if (sm == A.sample0 && A.hit) {
        // scene = pos + R^T p_body
        const float hx = (float)(sp.pos[0] + sp.ex[0] * px + sp.ey[0] * py + sp.ez[0] * pz);
        const float hy = (float)(sp.pos[1] + sp.ex[1] * px + sp.ey[1] * py + sp.ez[1] * pz);
        const float hz = (float)(sp.pos[2] + sp.ex[2] * px + sp.ey[2] * py + sp.ez[2] * pz);
        A.hit[(size_t)y * A.width + x] = make_float4(hx, hy, hz, (float)h.s);
    }
    if (A.hit64) write_hit64(A, R, h, x, y);
    lit = make_float3(0.f, 0.f, 0.f);
    if (!(cosl > 0.0f)) return false;
    float3 alb = make_float3(1.0f, 1.0f, 1.0f);
    if (A.tex.data) {
        const int w = A.tex.W, hgt = A.tex.H;
        const float u = ((float)h.c0 + 0.5f + h.fc) * ((float)w / (float)A.hf.W) - 0.5f;
        const float v = ((float)h.r0 + 0.5f + h.fr) * ((float)hgt / (float)A.hf.H) - 0.5f;
        const float fu = floorf(u);
        int c0 = (int)fu;
        const float fc = u - fu;
        c0 = c0 < 0 ? c0 + w : (c0 >= w ? c0 - w : c0);
        const int c1 = c0 + 1 == w ? 0 : c0 + 1;
        const int r0 = min(max((int)floorf(v), 0), hgt - 2);
        const float fr = fminf(fmaxf(v - (float)r0, 0.0f), 1.0f);
        const uchar4 ta = __ldg(A.tex.data + (size_t)r0 * w + c0), tb = __ldg(A.tex.data + (size_t)r0 * w + c1);
        const uchar4 tc = __ldg(A.tex.data + (size_t)(r0 + 1) * w + c0), td = __ldg(A.tex.data + (size_t)(r0 + 1) * w + c1);
        const float w00 = (1.0f - fc) * (1.0f - fr), w01 = fc * (1.0f - fr), w10 = (1.0f - fc) * fr, w11 = fc * fr;
        const float sc = 1.0f / 255.0f;
        alb = make_float3((ta.x * w00 + tb.x * w01 + tc.x * w10 + td.x * w11) * sc,
                          (ta.y * w00 + tb.y * w01 + tc.y * w10 + td.y * w11) * sc,
                          (ta.z * w00 + tb.z * w01 + tc.z * w10 + td.z * w11) * sc);
    }
    const float q = (float)(sp.light_radius * idist);
    const float E = (float)sp.light_radiance * q * q * cosl;
    lit = make_float3(alb.x * E, alb.y * E, alb.z * E);
    const double eps = sp.scene_epsilon;
    S.ox = fma(eps, (double)nx, px); S.oy = fma(eps, (double)ny, py); S.oz = fma(eps, (double)nz, pz);
    S.dx = lx; S.dy = ly; S.dz = lz;
    S.oo = S.ox * S.ox + S.oy * S.oy + S.oz * S.oz;
    S.od = S.ox * S.dx + S.oy * S.dy + S.oz * S.dz;
    return sp.shadows != 0;
}

#ifndef MRTX_FAST_MINBLOCKS
#define MRTX_FAST_MINBLOCKS 8
#endif
template <bool I16>
__global__ void __launch_bounds__(128, MRTX_FAST_MINBLOCKS)
trace_kernel_fast(const __grid_constant__ RenderArgs A) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int gl = A.g_log2, g = 1 << gl;
    const int sub = lane & (g - 1), pw = lane >> gl;         // lane within the pixel's group, pixel within the warp
    const unsigned ppw = 32u >> gl;
    const unsigned n_limb = A.work_counter[4];
    const unsigned nkept = n_limb + A.work_counter[1];
    const unsigned ntasks = (nkept + ppw - 1u) / ppw;
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
        const unsigned p = task * ppw + (unsigned)pw;
        const bool valid = p < nkept;
        const unsigned packed = valid ? list_pixel(A, p, n_limb) : 0u;
        const int x = (int)(packed & 0xffffu), y = (int)(packed >> 16);
        const uint32_t pixel = (uint32_t)y * (uint32_t)A.width + (uint32_t)x;
        float3 acc = make_float3(0.f, 0.f, 0.f);
        unsigned dmask = 0;                                  // group leader: deferred samples of this pixel

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
            for (int pass = 0; pass < 2; ++pass) {           // 0: primary ray, 1: shadow ray
                bool alive = false;
                if (want) {
                    if (pass == 0) primary_ray_fast(A, x, y, pixel, sm, R);
                    alive = walk_begin(A.hf, A.sp.radius, R, 0.0, pass ? (int)A.sp.start_shadow : A.hf.top - (int)A.sp.start_primary, st);
                    if (pass == 0) entered = alive;
                }
                int res = FT_MISS;
                while (__any_sync(FULL, alive)) {
                    RawPatch P;
                    float sx = 0.f;
                    int face = 4;
                    bool cand = false;
                    if (alive) {
                        for (;;) {
                            const int r = walk_step<I16>(A.hf, Rf, A.inv_rs, st, P, sx, face, cnt);
                            if (r == TR_CONTINUE) continue;
                            if (r == TR_END) alive = false; else cand = true;
                            break;
                        }
                    }
                    __syncwarp();
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
                    if (hit) {
                        Ray64 S;
                        want = shade_fast(A, R, fh, x, y, pixel, sm, lit, S);
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
                } else write_miss(A, x, y, sm == A.sample0);
            }
            // deferred samples of each pixel -> its leader's mask (bit = sample index in this launch)
            const unsigned dm = __ballot_sync(FULL, defer);
            if (dm) {
                const unsigned gm = g == 32 ? dm : (dm >> (pw << gl)) & ((1u << g) - 1u);
                dmask |= gm << (rd << gl);
                if (defer) ++n_defer;
            }
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
struct RayRec { double ox, oy, oz, dx, dy, dz, s_in; float smax; unsigned cell; };   // 64 B; smax < 0: nothing to walk
struct HitRec { double s; float fc, fr; int r0, c0; int status; unsigned pad; };      // 32 B; status -1: missed the bounding sphere
static_assert(sizeof(RayRec) == 64 && sizeof(HitRec) == 32, "record layout");

__device__ __forceinline__ void store_ray_rec(RayRec* dst, const Ray64& R, const Walk& st, bool alive) {
    double2* q = (double2*)dst;
    q[0] = make_double2(R.ox, R.oy); q[1] = make_double2(R.oz, R.dx); q[2] = make_double2(R.dy, R.dz);
    const float smax = alive ? st.smax : -1.0f;
    const unsigned cell = alive ? ((unsigned)st.J << 16) | (unsigned)st.I : 0u;
    q[3] = make_double2(alive ? st.s_in : 0.0, __hiloint2double((int)cell, __float_as_int(smax)));
}
__device__ __forceinline__ void load_ray_rec(const RayRec* src, Ray64& R) {
    const double2* q = (const double2*)src;
    const double2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    R.ox = a.x; R.oy = a.y; R.oz = b.x; R.dx = b.y; R.dy = c.x; R.dz = c.y;
}

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
                    st.s = 0.0f; st.steps = 0;
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
                } else write_miss(A, d.x, d.y, d.sm == A.sample0);
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

// ---- deferred samples: same walk, float64 referee per undecided patch ---------------------------------
// One WARP per deferred sample.  The sample is traced again from the start with the float32 walk and the
// filter; only where the filter says FT_DEFER does the float64 exact test of trace_core.cuh (exact in-cell
// pieces, walk-back through the neighbours) decide that patch.
// Deferred rays are the long grazing ones and there are only a few thousand of them, so what the launch takes is
// the longest serial chain in it.  The ray's path through the shell is therefore cut into pieces that lanes walk
// independently, 32 at a time, nearest first; the first hit is the hit of the nearest piece that has one.  (A piece
// that starts below the surface reports a hit at its start, which can only lose against the true crossing in an
// earlier piece.)
// Equal pieces are not equal work.  A sun ray at the horizon stays within the walk's 5 m margin of level ground for
// 4 km, and next to a pole those 4 km are tens of thousands of cells 10 cm wide: measured, ONE piece of ONE shadow
// ray 1.5 km from the south pole held 22 145 nodes and 14 642 patch tests and the launch took 25 ms instead of 2.
// A lane therefore walks a piece only as far as a budget lets it (SceneParams::referee_budget: nodes + 3 * patch
// tests, default 1500); what is left of the piece goes back on the warp's stack of intervals and is cut again.
constexpr int REFEREE_STACK = 96;           // pending intervals per warp
struct RefIv { double a, b; int depth; int pad; };        // depth > 0: what a lane left of a piece
enum { RS_CLEAR = 0, RS_HIT = 1, RS_MORE = 2 };

__device__ __forceinline__ double shfl_d(double v, int src) {
    return __hiloint2double(__shfl_sync(0xffffffffu, __double2hiint(v), src), __shfl_sync(0xffffffffu, __double2loint(v), src));
}

// One piece: the walk starts at s_lo (a little before the piece, where its first cell can be found safely); cells that
// end before s_own belong to the piece before and are only walked, not tested.
template <bool I16>
__device__ int trace_referee(const RenderArgs& A, const Ray64& R, double s_lo, double s_own, double s_hi, int start_level, float t0_rel,
                             bool any_hit, int budget, double& s_stop, bool& fast, FastHit& fh, TraceOut& h, Counters& cnt) {
    Walk w;
    if (!walk_begin(A.hf, A.sp.radius, R, s_lo, start_level, w, t0_rel)) return RS_CLEAR;
    w.smax = fminf(w.smax, (float)(s_hi - w.s_in));
    if (!(w.smax > 0.0f)) return RS_CLEAR;
    const float own_start = fmaxf((float)(s_own - w.s_in), 0.0f) + 1.0e-6f * A.K.R;
    const float own_from = own_start - 1.1f * t0_rel * A.K.R;
    int cost = 0;
    for (;;) {
        RawPatch P;
        float sx;
        int face;
        const int r = walk_step<I16>(A.hf, A.K.R, A.inv_rs, w, P, sx, face, cnt);
        if (r == TR_END) return RS_CLEAR;
        if (r == TR_CANDIDATE) {
            if (sx > own_from) {
                cost += 3;
                ++cnt.tests;
                const int t = fast_test<I16>(A.hf, A.K, R, w.s_in, s_lo, w.s, sx, w.smax, P, any_hit, fh) & 3;
                if (t == FT_HIT) { fast = true; return RS_HIT; }
                if (t == FT_DEFER) {
                    TravState st;
                    st.s_in = w.s_in; st.s_min = s_lo; st.s_end = w.s_in + (double)w.smax; st.s = w.s;
                    Patch Pd;
                    load_patch<I16>(A.hf, P.r0, P.c0, Pd);
                    if (exact_test<I16>(A.hf, A.sp.radius, R, st, Pd, sx, h, cnt)) { fast = false; return RS_HIT; }
                }
            }
            if (!walk_advance(A.hf, w, sx, face)) return RS_CLEAR;
        }
        // (what is handed back must be strictly shorter than the piece: only positions beyond its own start count)
        if (w.s > own_start && ++cost > budget) { s_stop = w.s_in + (double)w.s; return RS_MORE; }
    }
}

// First hit of R at s >= s_min by the whole warp.  Returns the lane that holds it (fast / fh / h valid there), or -1.
template <bool I16>
__device__ int referee_ray(const RenderArgs& A, RefIv* stack, const Ray64& R, double s_min, int start_level, bool any_hit,
                           bool& fast, FastHit& fh, TraceOut& h, Counters& cnt, bool& entered) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const double Rb = A.sp.radius * (double)A.hf.dmax;
    const double disc = R.od * R.od - (R.oo - Rb * Rb);
    entered = false;
    if (!(disc > 0.0)) return -1;
    const double sq = sqrt(disc);
    const double s1 = -R.od + sq;
    if (s1 <= s_min) return -1;
    const double s0 = fmax(s_min, -R.od - sq);
    entered = true;
    int top = 4;                                            // the path in four intervals, the nearest on top
    if (lane < 4) { RefIv& e = stack[3 - lane]; e.a = s0 + lane * 0.25 * (s1 - s0); e.b = lane == 3 ? s1 : s0 + (lane + 1) * 0.25 * (s1 - s0); e.depth = 0; }
    __syncwarp();
    int rounds = 0;
    while (top > 0) {
        // this round: the n nearest pending intervals, each cut into m pieces; lanes in order of distance
        // (one interval cut 32 ways at first; once a long chain has been split, its parts run side by side)
        const int n = any_hit ? min(top, 32) : 1, m = 32 / n;
        const int q = lane / m, j = lane - q * m;
        const bool work = q < n;
        const RefIv iv = stack[top - 1 - (work ? q : 0)];
        top -= n;
        __syncwarp();
        const double step = (iv.b - iv.a) / (double)m;
        const double own = iv.a + j * step, end = j == m - 1 ? iv.b : iv.a + (j + 1) * step;
        const bool first = !(own > s0);
        // Pieces overlap a little: a piece's first cell is found from a float32 position a step (t0) inside it, so the
        // walk starts 3 t0 early.  Among polar slivers that lead-in alone is thousands of cells: what comes back from a
        // lane that ran out of budget is cut with a tenth of it (still 30 times the float32 error of the position).
        const float t0_rel = iv.depth ? 1.0e-6f : 1.0e-5f;
        const double lap = 3.0 * (double)t0_rel * A.sp.radius;
        const double lo = first ? s_min : fmax(s_min, own - lap), hi = end >= s1 ? s1 + 1.0 : end;
        // (no room to split further, or splitting does not converge: walk it out)
        const int budget = top + 34 <= REFEREE_STACK && ++rounds < 512 ? (int)A.sp.referee_budget : 0x7fffffff;
        double s_stop = end;
        int st = RS_CLEAR;
        if (work) st = trace_referee<I16>(A, R, lo, first ? s_min : own, hi, first ? start_level : 2, t0_rel, any_hit, budget, s_stop, fast, fh, h, cnt);
        __syncwarp();
        const unsigned m_hit = __ballot_sync(FULL, st == RS_HIT), m_more = __ballot_sync(FULL, st == RS_MORE);
        const int first_hit = m_hit ? __ffs(m_hit) - 1 : 32;
        if (any_hit) {
            if (m_hit) return first_hit;                    // any crossing occludes
            if (st == RS_MORE) { RefIv& e = stack[top + __popc(m_more & ((1u << lane) - 1u))]; e.a = s_stop; e.b = end; e.depth = iv.depth + 1; }
            top += __popc(m_more);
            __syncwarp();
            continue;
        }
        // nearest hit: unfinished pieces in front of the first hit come first, then the piece that hit (traced again)
        const unsigned before = first_hit < 32 ? m_more & ((1u << first_hit) - 1u) : m_more;
        if (!before) {
            if (first_hit < 32) return first_hit;
            continue;
        }
        // (under it, what lies beyond that piece in this interval: only looked at should the piece not hit again)
        const int n_hit = first_hit < 32 ? 2 : 0;
        if (n_hit) {
            const double ha = shfl_d(own, first_hit), hb = shfl_d(end, first_hit);
            if (lane == 0) {
                stack[top].a = hb; stack[top].b = iv.b; stack[top].depth = iv.depth;
                stack[top + 1].a = ha; stack[top + 1].b = hb; stack[top + 1].depth = iv.depth;
            }
        }
        if (st == RS_MORE && lane < first_hit) {
            RefIv& e = stack[top + n_hit + __popc(before & ~((2u << lane) - 1u))];      // the nearest ends up on top
            e.a = s_stop; e.b = end; e.depth = iv.depth + 1;
        }
        top += n_hit + __popc(before);
        __syncwarp();
    }
    return -1;
}

// WAVE: entries are (list pixel, sample) items of the wavefront pipeline and the result goes to the item's
// radiance slot; otherwise (pixel, sample mask) entries of trace_kernel_fast and the result is added to the accumulator.
template <bool I16, bool WAVE>
__global__ void __launch_bounds__(64)
trace_kernel_referee(const __grid_constant__ RenderArgs A) {
    __shared__ RefIv stacks[2][REFEREE_STACK];
    RefIv* const stack = stacks[threadIdx.x >> 5];
    const unsigned total = A.work_counter[3];
    const int lane = threadIdx.x & 31;
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    Counters cnt = {0u, 0u, 0u};
    RayStats rs = {0u, 0u, 0u, 0u, 0u};                    // lane 0 counts rays
    const unsigned n_limb = A.work_counter[4];
    for (unsigned e = warp; e < total; e += nwarps) {
        const uint2 ent = WAVE ? A.defer_items[e] : A.defer_list[e];
        const unsigned packed = WAVE ? list_pixel(A, ent.x, n_limb) : ent.x;
        const int x = (int)(packed & 0xffffu), y = (int)(packed >> 16);
        const uint32_t pixel = (uint32_t)y * (uint32_t)A.width + (uint32_t)x;
        float3 acc = make_float3(0.f, 0.f, 0.f);            // lane 0 sums the samples in order
        for (unsigned mask = WAVE ? 1u << ent.y : ent.y; mask; mask &= mask - 1u) {
            const unsigned sm = A.sample0 + (unsigned)(__ffs(mask) - 1);
            Ray64 R, S;
            primary_ray_fast(A, x, y, pixel, sm, R);
            bool fast = false, entered = false;
            FastHit fh;
            TraceOut h;
            const int who = referee_ray<I16>(A, stack, R, 0.0, A.hf.top - 3, false, fast, fh, h, cnt, entered);
            if (lane == 0) { ++rs.primary; if (entered) ++rs.inside; }
            if (who < 0) { if (lane == 0) write_miss(A, x, y, sm == A.sample0); continue; }
            float3 lit = make_float3(0.f, 0.f, 0.f);
            bool need_shadow = false;
            if (lane == who) need_shadow = fast ? shade_fast(A, R, fh, x, y, pixel, sm, lit, S) : shade_hit(A, R, h, x, y, pixel, sm, lit, S);
            need_shadow = __shfl_sync(0xffffffffu, need_shadow ? 1 : 0, who) != 0;
            lit.x = __shfl_sync(0xffffffffu, lit.x, who); lit.y = __shfl_sync(0xffffffffu, lit.y, who); lit.z = __shfl_sync(0xffffffffu, lit.z, who);
            if (lane == 0) ++rs.hits;
            bool occluded = false;
            if (need_shadow) {
                S.ox = shfl_d(S.ox, who); S.oy = shfl_d(S.oy, who); S.oz = shfl_d(S.oz, who);
                S.dx = shfl_d(S.dx, who); S.dy = shfl_d(S.dy, who); S.dz = shfl_d(S.dz, who);
                S.oo = S.ox * S.ox + S.oy * S.oy + S.oz * S.oz;
                S.od = S.ox * S.dx + S.oy * S.dy + S.oz * S.dz;
                occluded = referee_ray<I16>(A, stack, S, 0.0, 2, true, fast, fh, h, cnt, entered) >= 0;
                if (lane == 0) { ++rs.shadow; if (occluded) ++rs.occluded; }
            }
            if (!occluded) { acc.x += lit.x; acc.y += lit.y; acc.z += lit.z; }
        }
        if (lane == 0) {
            if (WAVE) {
                float* slot = A.rad + ((size_t)(ent.x - A.wave_p0) * A.nsamples + ent.y) * 3;
                slot[0] = acc.x; slot[1] = acc.y; slot[2] = acc.z;
            } else {
                float4* ap = A.accum + (size_t)y * A.width + x; // the filtered kernel has counted the samples
                float4 old = *ap;
                old.x += acc.x; old.y += acc.y; old.z += acc.z;
                *ap = old;
            }
        }
    }
    __syncwarp();
    flush_counters(A, rs, cnt, lane);
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

static void to_body(const SceneParams& sp, const double* v, double* out) {
    out[0] = sp.ex[0] * v[0] + sp.ex[1] * v[1] + sp.ex[2] * v[2];
    out[1] = sp.ey[0] * v[0] + sp.ey[1] * v[1] + sp.ey[2] * v[2];
    out[2] = sp.ez[0] * v[0] + sp.ez[1] * v[1] + sp.ez[2] * v[2];
}

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

int launch_trace(mrtx_ctx* ctx, int x0, int y0, int x1, int y1, unsigned s0, unsigned ns) {
    RenderArgs A;
    A.hf = ctx->hf; A.tex = ctx->tex[0]; A.cam = ctx->cam; A.sp = ctx->sp;
    A.width = ctx->width; A.height = ctx->height;
    A.x0 = x0; A.y0 = y0; A.x1 = x1; A.y1 = y1;
    A.sample0 = s0; A.nsamples = ns;
    A.accum = ctx->accum; A.hit = ctx->hit;
    A.hit64 = ctx->sp.debug_hits ? ctx->hit64 : nullptr;
    A.counters = ctx->d_counters;
    A.work_counter = ctx->d_work;
    A.pixel_list = ctx->pixel_list;
    const double er[3] = {A.cam.eye[0] - A.sp.pos[0], A.cam.eye[1] - A.sp.pos[1], A.cam.eye[2] - A.sp.pos[2]};
    const double lr[3] = {A.sp.light_pos[0] - A.sp.pos[0], A.sp.light_pos[1] - A.sp.pos[1], A.sp.light_pos[2] - A.sp.pos[2]};
    to_body(A.sp, er, A.eye_b);
    to_body(A.sp, lr, A.light_b);
    A.defer_list = ctx->defer_list;
    A.rad = nullptr; A.rays = nullptr; A.hits = nullptr; A.srays = nullptr; A.sitem = nullptr; A.defer_items = nullptr;
    A.wave_p0 = 0; A.wave_np = 0; A.lvl_primary = 0; A.lvl_shadow = 0;
    A.defer_stats = ctx->d_defer_stats;
    A.K = make_fast_consts(ctx->hf, ctx->sp.radius);
    A.inv_rs = 1.0f / ctx->hf.radius_scale;
    A.g_log2 = 0;
    const bool i16 = ctx->hf.is_i16 != 0;
    unsigned kernel = ctx->sp.kernel;
    if (kernel >= 2 && !A.K.enabled) kernel = 1;             // map too coarse for the filter: everything would defer
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
    } else {
        // filtered kernel in chunks of <= 32 samples (one mask bit per sample in the deferred list), each
        // followed by the referee kernel over whatever it deferred
        int per_sm = 0;
        if (i16) MRTX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trace_kernel_fast<true>, 128, 0));
        else     MRTX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trace_kernel_fast<false>, 128, 0));
        if (per_sm < 1) per_sm = 1;
        for (unsigned done = 0; done < ns; done += 32u) {
            const unsigned n = ns - done < 32u ? ns - done : 32u;
            A.sample0 = s0 + done; A.nsamples = n;
            int gl = 0;
            while ((2u << gl) <= n && gl < 5) ++gl;
            A.g_log2 = gl;
            if (done) MRTX_CUDA(cudaMemsetAsync(A.work_counter + 2, 0, 2 * sizeof(unsigned), ctx->stream));
            const long long warps_needed = ((npix << gl) + 31) / 32;
            long long blocks = (long long)ctx->sm_count * per_sm;
            if (blocks * 4 > warps_needed) blocks = (warps_needed + 3) / 4;
            if (blocks < 1) blocks = 1;
            if (i16) trace_kernel_fast<true><<<(unsigned)blocks, 128, 0, ctx->stream>>>(A);
            else     trace_kernel_fast<false><<<(unsigned)blocks, 128, 0, ctx->stream>>>(A);
            if (i16) trace_kernel_referee<true, false><<<ctx->sm_count * 8, 64, 0, ctx->stream>>>(A);
            else     trace_kernel_referee<false, false><<<ctx->sm_count * 8, 64, 0, ctx->stream>>>(A);
        }
    }
    MRTX_CUDA(cudaGetLastError());
    return MRTX_OK;
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
