// Device helpers, launch arguments, the whole-pixel cull pass and the float64 referee, shared by
// trace.cu (production kernels) and trace_alt.cu (the A/B kernels 0, 1 and 3 kept for the tests).
// Everything lives in an anonymous namespace: each translation unit gets its own copy.
#pragma once

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
    Texture8 env;                    // environment texture (star map) seen by rays that miss the Moon; data null: black
    Camera cam;
    SceneParams sp;
    int width, height, x0, y0, x1, y1;
    unsigned sample0, nsamples;
    int tile_log2, tile_ranks, tile_rank;   // screen-tile sharding (mrtx_render_tiles): this launch renders the pixels of the
                                     // square tiles (side 2^tile_log2) t with t mod tile_ranks == tile_rank only
    unsigned hit_sample;             // the sample whose first hit goes to the hit buffer: sample 0 of the cycle, whichever launch,
                                     // chunk or rank traces it (rt._get_hit_at, moon_renderer.py:1138)
    float4* accum; float4* hit; double4* hit64;
    unsigned long long* counters;
    unsigned* work_counter;          // [0] persistent kernel: next unclaimed entry, [1] pixel list length,
                                     // [2] fast kernel: next unclaimed warp task, [3] deferred list length
    unsigned* pixel_list;            // pixels whose rays can touch the bounding sphere (x | y << 16): [0, work_counter[4])
                                     // limb pixels, [list_cap - work_counter[1], list_cap) the others, backwards
    unsigned list_cap;
    unsigned long long* defer_stats; // [reason + 16 * shadow]: why samples were deferred
    uint2* defer_list;               // (pixel, mask of samples sample0 + bit) the fast kernel could not certify
    unsigned* defer_mask;            // per pixel: samples that found the list full (defer_push); all zero between launches
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
    // shadow queue of the production path: rays pushed by trace_kernel_fast (work_counter[5] of them), their radiance and
    // (pixel | sample bit << 27), consumed by shadow_kernel (cursor work_counter[6]); accfix: see accfix_add()
    struct RayRec* sq_rays; uint4* sq_aux; unsigned sq_cap; int sq_level;
    // overlay tubes (mrtx_set_tubes): segments, per-tile lists (see mrtx_ctx::tube_tiles); n_tubes = 0: none
    const float4* tubes; unsigned n_tubes; const unsigned* tube_tiles; int tube_tx;
    // interreflection (path_seg_range): bounce rays in two queues of the shadow queue's layout - bq_in is
    // traced to its first hit by bounce_kernel (work_counter[8] rays, cursor [9]), bq_out is filled by shade_kernel
    // ([10] rays); their aux entries carry the path's throughput.  depth: 0 = camera hits, d = hits of the d-th bounce
    struct RayRec* bq_in_rays; uint4* bq_in_aux; struct RayRec* bq_out_rays; uint4* bq_out_aux;
    int depth, n_bounce;
    float night_sin2;                // sin^2 of the Sun's depression beyond which a path cannot reach lit terrain (shade_kernel)
    struct HardRay* hard;            // shadow rays handed from trace_kernel_referee to referee_hard_kernel (work_counter[12] of them)
    void* pool;                      // trace_kernel_pool: POOL_CAP parked rays per warp
    // hit queue (shadow_queue = 2): primary hits pushed by trace_kernel_fast (work_counter[7] of them), shaded by shade_kernel
    struct HitQRec* hq; unsigned hq_cap;
    unsigned long long* accfix;
    double* beam_s; unsigned char* beam_l;   // beam pre-pass, by position in the pixel list (null: no pre-pass)
    int beam_drop;
    // eye and light centre in the body frame (host-computed once per launch)
    double eye_b[3], light_b[3];
};

// ray record of a queue: the float64 ray, where its walk starts and the cell it starts in (J << 16 | I at the queue's start level)
struct RayRec { double ox, oy, oz, dx, dy, dz, s_in; float smax; unsigned cell; };   // 64 B; smax < 0: nothing to walk
static_assert(sizeof(RayRec) == 64, "record layout");

__device__ __forceinline__ void store_ray_rec(RayRec* dst, const Ray64& R, const Walk& st, bool alive) {
    double2* q = (double2*)dst;
    q[0] = make_double2(R.ox, R.oy); q[1] = make_double2(R.oz, R.dx); q[2] = make_double2(R.dy, R.dz);
    const float smax = alive ? st.smax : -1.0f;
    const unsigned cell = alive ? ((unsigned)st.J << 16) | (unsigned)st.I : 0u;
    q[3] = make_double2(alive ? st.s_in : 0.0, __hiloint2double((int)cell, __float_as_int(smax)));
}
// record of the hit queue: what fast_test() found (FastHit) + pixel | sample bit << 27.  The primary ray itself is not
// stored: it is a function of (pixel, sample) alone and shade_kernel evaluates it again, bit for bit.
struct HitQRec { double s; float fc, fr; int r0, c0; unsigned pix_k, ridx; float d00, d01, d10, d11; };   // 48 B; ridx: the bounce ray's queue entry
static_assert(sizeof(HitQRec) == 48, "record layout");

__device__ __forceinline__ void load_ray_rec(const RayRec* src, Ray64& R) {
    const double2* q = (const double2*)src;
    const double2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    R.ox = a.x; R.oy = a.y; R.oz = b.x; R.dx = b.y; R.dy = c.x; R.dz = c.y;
}


// accumulation that does not depend on the order of its terms: radiance sums in 2^-36 fixed point (64-bit atomics).
// fold_kernel adds them to the float accumulators at the end of a launch.
constexpr float ACCFIX_SCALE = 68719476736.0f;      // 2^36
__device__ __forceinline__ void accfix_add(unsigned long long* accfix, uint32_t pixel, float3 v) {
    unsigned long long* a = accfix + (size_t)pixel * 3;
    if (v.x > 0.f) atomicAdd(a + 0, __float2ull_rn(fminf(v.x, 6.0e7f) * ACCFIX_SCALE));
    if (v.y > 0.f) atomicAdd(a + 1, __float2ull_rn(fminf(v.y, 6.0e7f) * ACCFIX_SCALE));
    if (v.z > 0.f) atomicAdd(a + 2, __float2ull_rn(fminf(v.z, 6.0e7f) * ACCFIX_SCALE));
}

struct RayStats { unsigned primary, inside, hits, shadow, occluded; };

// One sample for the referee: an entry of the deferred list (list_cap = one per pixel of the frame).  Writers that make an
// entry per SAMPLE (shadow_kernel, trace_kernel_pool) can fill the list when nearly every ray defers (long_walk of a few
// steps); what does not fit is kept as a bit of the pixel's word in defer_mask, which the referee scans if - and only if -
// work_counter[3] has passed list_cap.  No sample is ever dropped.
__device__ __forceinline__ void defer_push(const RenderArgs& A, uint32_t pixel, unsigned k) {
    const unsigned slot = atomicAdd(&A.work_counter[3], 1u);
    if (slot < A.list_cap) {
        const unsigned px = pixel % (unsigned)A.width, py = pixel / (unsigned)A.width;
        A.defer_list[slot] = make_uint2(px | (py << 16), 1u << k);
    } else atomicOr(&A.defer_mask[pixel], 1u << k);
}

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
    if (sm == A.hit_sample && A.hit) {
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

// ---- overlay tubes ----------------------------------------------------------------------------------
// Grid lines, labels and pins are graphs of thin tubes floating 0.5 % above the sphere (moon_grid.py:188, 273), drawn with a
// flat material that shadow rays pass through (renderer_labels.py:133-139): a camera sample that meets a tube before the
// terrain takes the tube's colour as its radiance, nothing else changes.  A segment is a capsule (two end points, one
// radius); the nearest one along the sample's ray is found among the segments binned to the pixel's screen tile.
__device__ __forceinline__ unsigned tube_tile_count(const RenderArgs& A, int x, int y) {
    if (!A.n_tubes) return 0u;
    return __ldg(A.tube_tiles + (size_t)((y >> MRTX_TUBE_TILE_LOG2) * A.tube_tx + (x >> MRTX_TUBE_TILE_LOG2)) * (MRTX_TUBE_TILE_CAP + 2));
}

// first intersection of the ray o + t d (|d| = 1) with the capsule (a, b, r), or a negative number
__device__ __forceinline__ double capsule_hit(const double o[3], const double d[3], const float4 sa, const float4 sb) {
    const double ax = sa.x, ay = sa.y, az = sa.z, r = sa.w;
    const double bax = (double)sb.x - ax, bay = (double)sb.y - ay, baz = (double)sb.z - az;
    const double oax = o[0] - ax, oay = o[1] - ay, oaz = o[2] - az;
    const double baba = bax * bax + bay * bay + baz * baz, bard = bax * d[0] + bay * d[1] + baz * d[2];
    const double baoa = bax * oax + bay * oay + baz * oaz, rdoa = d[0] * oax + d[1] * oay + d[2] * oaz;
    const double oaoa = oax * oax + oay * oay + oaz * oaz;
    const double qa = baba - bard * bard, qb = baba * rdoa - baoa * bard, qc = baba * oaoa - baoa * baoa - r * r * baba;
    double best = -1.0;
    if (qa > 1.0e-12 * baba) {
        const double h = qb * qb - qa * qc;
        if (h >= 0.0) {
            const double t = (-qb - sqrt(h)) / qa;
            const double yy = baoa + t * bard;
            if (yy > 0.0 && yy < baba) return t;            // the cylinder between the caps (entry point: nearest of all)
        }
    }
    // the spheres at the two ends
    {
        const double B = rdoa, Cc = oaoa - r * r, h = B * B - Cc;
        if (h > 0.0) best = -B - sqrt(h);
    }
    {
        const double obx = oax - bax, oby = oay - bay, obz = oaz - baz;
        const double B = d[0] * obx + d[1] * oby + d[2] * obz, Cc = obx * obx + oby * oby + obz * obz - r * r, h = B * B - Cc;
        if (h > 0.0) { const double t = -B - sqrt(h); if (t > 0.0 && (best <= 0.0 || t < best)) best = t; }
    }
    return best;
}

// nearest tube along the primary ray of (x, y; sample sm) before s_max; writes the hit buffer like a surface hit would
__device__ __noinline__ bool tube_nearest(const RenderArgs& A, int x, int y, uint32_t pixel, unsigned sm, double s_max, float3& col) {
    const unsigned* L = A.tube_tiles + (size_t)((y >> MRTX_TUBE_TILE_LOG2) * A.tube_tx + (x >> MRTX_TUBE_TILE_LOG2)) * (MRTX_TUBE_TILE_CAP + 2);
    const unsigned cnt = __ldg(L);
    if (!cnt) return false;
    const Camera& cam = A.cam;
    const bool j = A.sp.jitter != 0;
    const double jx = j ? rnd(pixel, sm, 0) : 0.5, jy = j ? rnd(pixel, sm, 1) : 0.5;
    const double aspect = (double)A.width / (double)A.height;
    const double sx = ((x + jx) / A.width * 2.0 - 1.0) * cam.tan_half_fov * aspect;
    const double sy = (1.0 - (y + jy) / A.height * 2.0) * cam.tan_half_fov;
    double d[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) d[a] = cam.w[a] + sx * cam.right[a] + sy * cam.up[a];
    const double dn = 1.0 / sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
#pragma unroll
    for (int a = 0; a < 3; ++a) d[a] *= dn;
    const bool all = cnt > MRTX_TUBE_TILE_CAP;                 // list overflowed: every segment is a candidate
    const unsigned n = all ? A.n_tubes : cnt;
    double best = s_max;
    int which = -1;
    const float dfx = (float)d[0], dfy = (float)d[1], dfz = (float)d[2];
    const float efx = (float)cam.eye[0], efy = (float)cam.eye[1], efz = (float)cam.eye[2];
    for (unsigned i = 0; i < n; ++i) {
        const unsigned k = all ? i : __ldg(L + 1 + i);
        const float4 sa = __ldg(A.tubes + 3 * (size_t)k), sb = __ldg(A.tubes + 3 * (size_t)k + 1);
        {
            // float32 filter: a ray further from the capsule's AXIS LINE than its radius (+ 1e-3: thirty times what float32
            // loses on coordinates of this size) cannot meet it; the float64 test below is for the few that come close
            const float ux = sb.x - sa.x, uy = sb.y - sa.y, uz = sb.z - sa.z;
            const float nx = dfy * uz - dfz * uy, ny = dfz * ux - dfx * uz, nz = dfx * uy - dfy * ux;
            const float nn = nx * nx + ny * ny + nz * nz;
            const float wx = sa.x - efx, wy = sa.y - efy, wz = sa.z - efz;
            const float wn = wx * nx + wy * ny + wz * nz, lim = sa.w + 1.0e-3f;
            if (nn > 1.0e-12f * (ux * ux + uy * uy + uz * uz) && wn * wn > lim * lim * nn) continue;
        }
        const double t = capsule_hit(cam.eye, d, sa, sb);
        if (t > 0.0 && t < best) { best = t; which = (int)k; }
    }
    if (which < 0) return false;
    const float4 c = __ldg(A.tubes + 3 * (size_t)which + 2);
    col = make_float3(c.x, c.y, c.z);
    if (sm == A.hit_sample && A.hit)
        A.hit[(size_t)y * A.width + x] = make_float4((float)(cam.eye[0] + best * d[0]), (float)(cam.eye[1] + best * d[1]),
                                                     (float)(cam.eye[2] + best * d[2]), (float)best);
    if (A.hit64) A.hit64[(size_t)y * A.width + x] = make_double4(-2.0, 0.0, 0.0, best);      // s = -2: a tube, at distance w
    return true;
}

// bins the segments to the screen tiles their projection (grown by the tube radius and a pixel of jitter) touches
__global__ void __launch_bounds__(128)
tube_bin_kernel(const float4* __restrict__ tubes, unsigned n, unsigned* __restrict__ tiles, int tx, int ty, Camera cam, int width, int height) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 sa = tubes[3 * (size_t)i], sb = tubes[3 * (size_t)i + 1];
    const double aspect = (double)width / (double)height;
    double lo_x = 1e30, hi_x = -1e30, lo_y = 1e30, hi_y = -1e30;
    bool everywhere = false;
    const float4 ends[2] = {sa, sb};
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const double vx = (double)ends[e].x - cam.eye[0], vy = (double)ends[e].y - cam.eye[1], vz = (double)ends[e].z - cam.eye[2];
        const double z = vx * cam.w[0] + vy * cam.w[1] + vz * cam.w[2];
        if (!(z > 4.0 * (double)sa.w + 1e-6)) { everywhere = true; continue; }       // at or behind the eye
        const double px = ((vx * cam.right[0] + vy * cam.right[1] + vz * cam.right[2]) / z / (cam.tan_half_fov * aspect) + 1.0) * 0.5 * width;
        const double py = (1.0 - (vx * cam.up[0] + vy * cam.up[1] + vz * cam.up[2]) / z / cam.tan_half_fov) * 0.5 * height;
        // radius on the screen (x 3: off-axis stretch of the perspective projection, 1 / cos of up to 64 deg at the corners of a
        // 90-degree view, and some)
        const double pr = 3.0 * (double)sa.w / z / cam.tan_half_fov * 0.5 * height + 2.0;
        lo_x = fmin(lo_x, px - pr); hi_x = fmax(hi_x, px + pr); lo_y = fmin(lo_y, py - pr); hi_y = fmax(hi_y, py + pr);
    }
    int t0x = 0, t1x = tx - 1, t0y = 0, t1y = ty - 1;
    if (!everywhere) {
        if (hi_x < 0.0 || hi_y < 0.0 || lo_x >= (double)width || lo_y >= (double)height) return;
        t0x = max((int)floor(lo_x) >> MRTX_TUBE_TILE_LOG2, 0); t1x = min((int)floor(hi_x) >> MRTX_TUBE_TILE_LOG2, tx - 1);
        t0y = max((int)floor(lo_y) >> MRTX_TUBE_TILE_LOG2, 0); t1y = min((int)floor(hi_y) >> MRTX_TUBE_TILE_LOG2, ty - 1);
    }
    for (int y = t0y; y <= t1y; ++y)
        for (int x = t0x; x <= t1x; ++x) {
            unsigned* L = tiles + (size_t)(y * tx + x) * (MRTX_TUBE_TILE_CAP + 2);
            const unsigned k = atomicAdd(L, 1u);
            if (k < MRTX_TUBE_TILE_CAP) L[1 + k] = i;
        }
}

__device__ __forceinline__ void flush_counters(const RenderArgs& A, const RayStats& rs, const Counters& cnt, int lane) {
    const unsigned vals[8] = {rs.primary, rs.inside, rs.hits, rs.shadow, rs.occluded, cnt.nodes, cnt.tests, cnt.overflow};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const unsigned v = __reduce_add_sync(0xffffffffu, vals[i]);
        if (lane == 0 && v) atomicAdd(&A.counters[i], (unsigned long long)v);
    }
}

// ---- pass 1 of the production path: whole-pixel cull + compaction -------------------------------------
// 64 % of a whole-disk frame never touches the Moon.  One thread per pixel (8x4 tiles, so the list
// keeps screen-space coherence) tests the pixel's centre ray against the bounding sphere grown by 1.5
// pixels; pixels that cannot hit are finished here, the rest are appended to the work list that the
// persistent kernel consumes - its lanes then only ever receive pixels with real work.
__device__ __forceinline__ void primary_ray_fast(const RenderArgs& A, int x, int y, uint32_t pixel, unsigned sm, Ray64& R);
__device__ __forceinline__ float3 miss_radiance_body(const RenderArgs& A, const Ray64& R);
__device__ __forceinline__ bool sees_background(const RenderArgs& A);

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
        bool mine = true;
        if (A.tile_ranks > 1) {
            const unsigned ttx = ((unsigned)A.width + (1u << A.tile_log2) - 1u) >> A.tile_log2;
            const unsigned t = ((unsigned)y >> A.tile_log2) * ttx + ((unsigned)x >> A.tile_log2);
            mine = t % (unsigned)A.tile_ranks == (unsigned)A.tile_rank;
        }
        if (x < A.x1 && y < A.y1 && mine) {
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
            if (eye_dist > cull_r && (d2 > cull_r * cull_r || od > 0.0) && !tube_tile_count(A, x, y)) {
                culled = 1;                                     // every sample of this pixel misses
                write_miss(A, x, y, true);
                float4* ap = A.accum + (size_t)y * A.width + x;
                float4 old = *ap;
                old.w += (float)A.nsamples;
                if (sees_background(A)) {
                    // star map / Sun disk: every sample's own direction, summed in sample order
                    const uint32_t pixel = (uint32_t)y * (uint32_t)A.width + (uint32_t)x;
                    for (unsigned k = 0; k < A.nsamples; ++k) {
                        Ray64 R;
                        primary_ray_fast(A, x, y, pixel, A.sample0 + k, R);
                        const float3 m = miss_radiance_body(A, R);
                        old.x += m.x; old.y += m.y; old.z += m.z;
                    }
                }
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

// ray through the point (x + jx, y + jy) of the frame, in the body frame
__device__ __forceinline__ void primary_ray_at(const RenderArgs& A, int x, int y, double jx, double jy, Ray64& R) {
    const SceneParams& sp = A.sp;
    const Camera& cam = A.cam;
    const double aspect = (double)A.width / (double)A.height;
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

__device__ __forceinline__ void primary_ray_fast(const RenderArgs& A, int x, int y, uint32_t pixel, unsigned sm, Ray64& R) {
    const bool j = A.sp.jitter != 0;
    primary_ray_at(A, x, y, j ? rnd(pixel, sm, 0) : 0.5, j ? rnd(pixel, sm, 1) : 0.5, R);
}

// ---- rays that miss the Moon --------------------------------------------------------------------
// The visible Sun disk (a flat-shaded sphere: its colour, no lighting, never an occluder - moon_renderer.py:133-139,
// 643-650) and the star map (environment texture of linear radiance, equirectangular: scene +Z up, -Y at longitude 0,
// bilinear, columns wrap, rows clamp).  d = unit direction from the eye, scene space.
__device__ __forceinline__ bool sees_background(const RenderArgs& A) { return A.env.data != nullptr || A.sp.sun_disk_radius > 0.0; }

#ifndef MRTX_MISS_ATTR
#define MRTX_MISS_ATTR __noinline__     // rare in the walk kernels (rays that graze past the limb), 6 KB of code when inlined
#endif
__device__ MRTX_MISS_ATTR float3 miss_radiance(const RenderArgs& A, double dx, double dy, double dz) {
    const SceneParams& sp = A.sp;
    if (sp.sun_disk_radius > 0.0) {
        const double ox = A.cam.eye[0] - sp.sun_disk_pos[0], oy = A.cam.eye[1] - sp.sun_disk_pos[1], oz = A.cam.eye[2] - sp.sun_disk_pos[2];
        const double b = ox * dx + oy * dy + oz * dz, c = ox * ox + oy * oy + oz * oz - sp.sun_disk_radius * sp.sun_disk_radius;
        const double disc = b * b - c;
        if (disc > 0.0 && -b + sqrt(disc) > 0.0) return make_float3(sp.sun_disk_color[0], sp.sun_disk_color[1], sp.sun_disk_color[2]);
    }
    if (!A.env.data) return make_float3(0.f, 0.f, 0.f);
    const int w = A.env.W, h = A.env.H;
    // (float32 angles: 1e-7 rad is 3e-4 texel of a 16k star map - the lookup runs for every sample of every pixel beside the
    //  Moon, 85 M times per 4K frame)
    const float fx = (float)dx, fy = (float)dy, fz = (float)dz;
    const float lon = atan2f(fx, -fy), lat = atan2f(fz, sqrtf(fx * fx + fy * fy));
    const float u = (lon * (0.5f / PI_F) + 0.5f) * (float)w - 0.5f, v = (0.5f - lat * (1.0f / PI_F)) * (float)h - 0.5f;
    const float fu = floorf(u);
    int c0 = (int)fu;
    const float fc = u - fu;
    c0 = c0 < 0 ? c0 + w : (c0 >= w ? c0 - w : c0);
    const int c1 = c0 + 1 == w ? 0 : c0 + 1;
    const int r0 = min(max((int)floorf(v), 0), h - 2);
    const float fr = fminf(fmaxf(v - (float)r0, 0.0f), 1.0f);
    const uchar4 a = __ldg(A.env.data + (size_t)r0 * w + c0), b = __ldg(A.env.data + (size_t)r0 * w + c1);
    const uchar4 c = __ldg(A.env.data + (size_t)(r0 + 1) * w + c0), e = __ldg(A.env.data + (size_t)(r0 + 1) * w + c1);
    const float s = 1.0f / 255.0f;
    return make_float3(((a.x * (1.0f - fc) + b.x * fc) * (1.0f - fr) + (c.x * (1.0f - fc) + e.x * fc) * fr) * s,
                       ((a.y * (1.0f - fc) + b.y * fc) * (1.0f - fr) + (c.y * (1.0f - fc) + e.y * fc) * fr) * s,
                       ((a.z * (1.0f - fc) + b.z * fc) * (1.0f - fr) + (c.z * (1.0f - fc) + e.z * fc) * fr) * s);
}

// the same for a body-frame ray direction (scene = R^T body)
__device__ __forceinline__ float3 miss_radiance_body(const RenderArgs& A, const Ray64& R) {
    const SceneParams& sp = A.sp;
    return miss_radiance(A, sp.ex[0] * R.dx + sp.ey[0] * R.dy + sp.ez[0] * R.dz,
                            sp.ex[1] * R.dx + sp.ey[1] * R.dy + sp.ez[1] * R.dz,
                            sp.ex[2] * R.dx + sp.ey[2] * R.dy + sp.ez[2] * R.dz);
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
#ifndef MRTX_SHADE_ATTR
#define MRTX_SHADE_ATTR __forceinline__
#endif
// what the interreflection pass needs of a shaded hit: where it is, its normal and its albedo
struct ShadeAux { double px, py, pz; float nx, ny, nz; float3 alb; };

// dim0: first random dimension of the light sample (2 at the camera hit, 2 + 4 d at the d-th bounce); primary: the hit
// of a camera ray (the only one that goes to the hit buffer)
__device__ MRTX_SHADE_ATTR bool shade_fast(const RenderArgs& A, const Ray64& R, const FastHit& h, int x, int y,
                                           uint32_t pixel, unsigned sm, float3& lit, Ray64& S,
                                           unsigned dim0 = 2u, bool primary = true, ShadeAux* aux = nullptr) {
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
        const float rr = (float)sp.light_radius * sqrtf((float)rnd(pixel, sm, dim0));
        float st, ct;
        sincospif(2.0f * (float)rnd(pixel, sm, dim0 + 1u), &st, &ct);
        tx += (double)(rr * (ct * b1x + st * b2x)); ty += (double)(rr * (ct * b1y + st * b2y)); tz += (double)(rr * (ct * b1z + st * b2z));
    }
    const double ln = d_rsqrt(tx * tx + ty * ty + tz * tz);
    const double lx = tx * ln, ly = ty * ln, lz = tz * ln;
    const float cosl = nx * (float)lx + ny * (float)ly + nz * (float)lz;
    if (primary && sm == A.hit_sample && A.hit) {
        // scene = pos + R^T p_body
        const float hx = (float)(sp.pos[0] + sp.ex[0] * px + sp.ey[0] * py + sp.ez[0] * pz);
        const float hy = (float)(sp.pos[1] + sp.ex[1] * px + sp.ey[1] * py + sp.ez[1] * pz);
        const float hz = (float)(sp.pos[2] + sp.ex[2] * px + sp.ey[2] * py + sp.ez[2] * pz);
        A.hit[(size_t)y * A.width + x] = make_float4(hx, hy, hz, (float)h.s);
    }
    if (primary && A.hit64) write_hit64(A, R, h, x, y);
    lit = make_float3(0.f, 0.f, 0.f);
    if (aux) { aux->px = px; aux->py = py; aux->pz = pz; aux->nx = nx; aux->ny = ny; aux->nz = nz; aux->alb = make_float3(0.f, 0.f, 0.f); }
    if (!(cosl > 0.0f) && !aux) return false;
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
    if (aux) aux->alb = alb;
    if (!(cosl > 0.0f)) return false;
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
// max_rounds (any-hit rays only): after that many rounds with intervals still pending the warp gives up and returns -2 with
// the pending intervals in stack[0 .. n_left): the caller hands them to referee_hard_kernel.
__device__ int referee_ray(const RenderArgs& A, RefIv* stack, const Ray64& R, double s_min, int start_level, bool any_hit,
                           bool& fast, FastHit& fh, TraceOut& h, Counters& cnt, bool& entered, int max_rounds = 0x7fffffff,
                           int* n_left = nullptr) {
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
        if (any_hit && rounds >= max_rounds) { if (n_left) *n_left = top; return -2; }
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

// ---- second tier: shadow rays whose chain one warp cannot finish in time -------------------------------------------------------
// Measured on the terminator sweep (config 4, frames 24..62): ONE sun ray that skims the ground 1.5 km from the south pole
// crosses tens of thousands of cells 10 cm wide and holds one warp of trace_kernel_referee for up to 19 ms while 2 367 others
// have long finished (frame 56: 50 ms instead of 31).  A warp that has spent MRTX_HARD_ROUNDS rounds on a shadow ray now
// publishes the ray, what it would add to the pixel and the intervals still pending; referee_hard_kernel then walks each
// such ray with the whole grid - thousands of short pieces, each decided by the same float32 filter + float64 exact test.
// Any crossing occludes, so the pieces need no order.
#ifndef MRTX_REFEREE_CLOCK
#define MRTX_REFEREE_CLOCK 0
#endif
#ifndef MRTX_HARD_ROUNDS
#define MRTX_HARD_ROUNDS 1
#endif
constexpr int HARD_MAX_RAYS = 256;
struct HardRay {
    double ox, oy, oz, dx, dy, dz;
    float lit[3]; uint32_t pixel;
    int n_iv, occluded, pad0, pad1;
    RefIv iv[REFEREE_STACK];
};

// WAVE: entries are (list pixel, sample) items of the wavefront pipeline and the result goes to the item's
// radiance slot; otherwise (pixel, sample mask) entries of trace_kernel_fast and the result is added to the accumulator.
template <bool I16, bool WAVE>
__global__ void __launch_bounds__(64)
trace_kernel_referee(const __grid_constant__ RenderArgs A) {
    __shared__ RefIv stacks[2][REFEREE_STACK];
    RefIv* const stack = stacks[threadIdx.x >> 5];
    // entries [0, n_list) come from the list; if it overflowed, entries n_list + p are the pixels of the frame with the
    // samples defer_push() parked in their mask words (cleared here for the next launch)
    const unsigned n_pushed = A.work_counter[3];
    const unsigned n_list = WAVE ? n_pushed : min(n_pushed, A.list_cap);
    const unsigned total = !WAVE && n_pushed > A.list_cap ? n_list + (unsigned)(A.width * A.height) : n_list;
    const int lane = threadIdx.x & 31;
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    Counters cnt = {0u, 0u, 0u};
    RayStats rs = {0u, 0u, 0u, 0u, 0u};                    // lane 0 counts rays
    const unsigned n_limb = A.work_counter[4];
    for (unsigned e = warp; e < total; e += nwarps) {
        uint2 ent;
        if (WAVE) ent = A.defer_items[e];
        else if (e < n_list) ent = A.defer_list[e];
        else {
            const unsigned p = e - n_list, m = A.defer_mask[p];
            if (!m) continue;
            __syncwarp();
            if (lane == 0) A.defer_mask[p] = 0u;
            ent = make_uint2((p % (unsigned)A.width) | ((p / (unsigned)A.width) << 16), m);
        }
        const unsigned packed = WAVE ? list_pixel(A, ent.x, n_limb) : ent.x;
        const int x = (int)(packed & 0xffffu), y = (int)(packed >> 16);
        const uint32_t pixel = (uint32_t)y * (uint32_t)A.width + (uint32_t)x;
#if MRTX_REFEREE_CLOCK
        const long long t_entry = clock64();
        long long t_primary = 0;
#endif
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
#if MRTX_REFEREE_CLOCK
            t_primary = clock64() - t_entry;
#endif
            if (who < 0) {
                if (lane == 0) {
                    float3 tc;
                    if (tube_tile_count(A, x, y) && tube_nearest(A, x, y, pixel, sm, 1.0e300, tc)) { acc.x += tc.x; acc.y += tc.y; acc.z += tc.z; }
                    else {
                        write_miss(A, x, y, sm == A.hit_sample);
                        if (sees_background(A)) { const float3 m = miss_radiance_body(A, R); acc.x += m.x; acc.y += m.y; acc.z += m.z; }
                    }
                }
                continue;
            }
            float3 lit = make_float3(0.f, 0.f, 0.f);
            bool need_shadow = false;
            if (lane == who) {
                // (a tube in front of the hit: its flat colour is the sample)
                if (tube_tile_count(A, x, y) && tube_nearest(A, x, y, pixel, sm, fast ? fh.s : h.s, lit)) need_shadow = false;
                else need_shadow = fast ? shade_fast(A, R, fh, x, y, pixel, sm, lit, S) : shade_hit(A, R, h, x, y, pixel, sm, lit, S);
            }
            need_shadow = __shfl_sync(0xffffffffu, need_shadow ? 1 : 0, who) != 0;
            lit.x = __shfl_sync(0xffffffffu, lit.x, who); lit.y = __shfl_sync(0xffffffffu, lit.y, who); lit.z = __shfl_sync(0xffffffffu, lit.z, who);
            if (lane == 0) ++rs.hits;
            bool occluded = false;
            if (need_shadow) {
                S.ox = shfl_d(S.ox, who); S.oy = shfl_d(S.oy, who); S.oz = shfl_d(S.oz, who);
                S.dx = shfl_d(S.dx, who); S.dy = shfl_d(S.dy, who); S.dz = shfl_d(S.dz, who);
                S.oo = S.ox * S.ox + S.oy * S.oy + S.oz * S.oz;
                S.od = S.ox * S.dx + S.oy * S.dy + S.oz * S.dz;
                int n_left = 0;
                const int sr = referee_ray<I16>(A, stack, S, 0.0, 2, true, fast, fh, h, cnt, entered,
                                                !WAVE && A.hard ? MRTX_HARD_ROUNDS : 0x7fffffff, &n_left);
                occluded = sr >= 0;
                if (sr == -2) {
                    // too long a chain for one warp: the grid finishes it (referee_hard_kernel adds lit if the sun is visible)
                    unsigned slot = 0;
                    if (lane == 0) slot = atomicAdd(&A.work_counter[12], 1u);
                    slot = __shfl_sync(0xffffffffu, slot, 0);
                    if (slot < (unsigned)HARD_MAX_RAYS) {
                        HardRay* hr = A.hard + slot;
                        if (lane == 0) {
                            hr->ox = S.ox; hr->oy = S.oy; hr->oz = S.oz; hr->dx = S.dx; hr->dy = S.dy; hr->dz = S.dz;
                            hr->lit[0] = lit.x; hr->lit[1] = lit.y; hr->lit[2] = lit.z; hr->pixel = pixel;
                            hr->n_iv = n_left; hr->occluded = 0;
                        }
                        for (int i = lane; i < n_left; i += 32) hr->iv[i] = stack[i];
                        __syncwarp();
                        occluded = true;                    // (nothing is added here)
                        if (lane == 0) ++rs.shadow;
                        continue;
                    }
                    // (list full: finish it here after all)
                    occluded = referee_ray<I16>(A, stack, S, 0.0, 2, true, fast, fh, h, cnt, entered) >= 0;
                }
                if (lane == 0) { ++rs.shadow; if (occluded) ++rs.occluded; }
            }
            if (!occluded) { acc.x += lit.x; acc.y += lit.y; acc.z += lit.z; }
        }
#if MRTX_REFEREE_CLOCK
        if (lane == 0) printf("REF %u %u %lld %lld %lld\n", e, blockIdx.x, t_entry, t_primary, clock64() - t_entry);
#endif
        if (lane == 0) {
            if (WAVE) {
                float* slot = A.rad + ((size_t)(ent.x - A.wave_p0) * A.nsamples + ent.y) * 3;
                slot[0] = acc.x; slot[1] = acc.y; slot[2] = acc.z;
            } else {
                accfix_add(A.accfix, pixel, acc);               // (the filtered kernel has counted the samples; a pixel may
                                                                //  have several entries: order-independent sum)
            }
        }
    }
    __syncwarp();
    flush_counters(A, rs, cnt, lane);
}

// One thread per piece: the pending intervals of hard ray blockIdx.y, each cut into gridDim.x * blockDim.x / n_iv pieces.
template <bool I16>
__global__ void __launch_bounds__(64)
referee_hard_kernel(const __grid_constant__ RenderArgs A) {
    const unsigned n_hard = min(A.work_counter[12], (unsigned)HARD_MAX_RAYS);
    Counters cnt = {0u, 0u, 0u};
    for (unsigned hi = blockIdx.y; hi < n_hard; hi += gridDim.y) {
        HardRay* hr = A.hard + hi;
        const int n_iv = hr->n_iv;
        if (n_iv <= 0) continue;
        const unsigned P = gridDim.x * blockDim.x, t = blockIdx.x * blockDim.x + threadIdx.x;
        const unsigned m = max(P / (unsigned)n_iv, 1u);                   // pieces per interval
        const unsigned i = t / m, j = t - i * m;
        if (i >= (unsigned)n_iv) continue;
        if (*(volatile int*)&hr->occluded) continue;
        Ray64 S;
        S.ox = hr->ox; S.oy = hr->oy; S.oz = hr->oz; S.dx = hr->dx; S.dy = hr->dy; S.dz = hr->dz;
        S.oo = S.ox * S.ox + S.oy * S.oy + S.oz * S.oz; S.od = S.ox * S.dx + S.oy * S.dy + S.oz * S.dz;
        const RefIv iv = hr->iv[i];
        const double step = (iv.b - iv.a) / (double)m;
        const double own = iv.a + j * step, end = j == m - 1 ? iv.b : iv.a + (j + 1) * step;
        // (pieces start a little early, as in referee_ray: the first cell is found from a float32 position inside the piece)
        const float t0_rel = 1.0e-6f;
        const double lap = 3.0 * (double)t0_rel * A.sp.radius;
        const double lo = fmax(0.0, own - lap);
        double s_stop = end;
        bool fast = false;
        FastHit fh;
        TraceOut h;
        const int st = trace_referee<I16>(A, S, lo, own, end, 2, t0_rel, true, 0x7fffffff, s_stop, fast, fh, h, cnt);
        if (st == RS_HIT) atomicExch(&hr->occluded, 1);
    }
    const Counters& c = cnt;
    const RayStats rs = {0u, 0u, 0u, 0u, 0u};
    flush_counters(A, rs, c, threadIdx.x & 31);
}

// what the hard rays add to their pixels, once all their pieces are decided
__global__ void referee_hard_finish_kernel(const __grid_constant__ RenderArgs A) {
    const unsigned n_hard = min(A.work_counter[12], (unsigned)HARD_MAX_RAYS);
    unsigned occluded = 0;
    for (unsigned hi = threadIdx.x; hi < n_hard; hi += blockDim.x) {
        const HardRay* hr = A.hard + hi;
        if (hr->occluded) ++occluded;
        else accfix_add(A.accfix, hr->pixel, make_float3(hr->lit[0], hr->lit[1], hr->lit[2]));
    }
    if (occluded) atomicAdd(&A.counters[4], (unsigned long long)occluded);
}

}  // namespace

static void to_body(const SceneParams& sp, const double* v, double* out) {
    out[0] = sp.ex[0] * v[0] + sp.ex[1] * v[1] + sp.ex[2] * v[2];
    out[1] = sp.ey[0] * v[0] + sp.ey[1] * v[1] + sp.ey[2] * v[2];
    out[2] = sp.ez[0] * v[0] + sp.ez[1] * v[1] + sp.ez[2] * v[2];
}


// launch arguments common to every kernel of the path
static void fill_render_args(mrtx_ctx* ctx, int x0, int y0, int x1, int y1, unsigned s0, unsigned ns, RenderArgs& A) {
    A.hf = ctx->hf; A.tex = ctx->tex[0]; A.env = ctx->tex[2]; A.cam = ctx->cam; A.sp = ctx->sp;
    A.width = ctx->width; A.height = ctx->height;
    A.x0 = x0; A.y0 = y0; A.x1 = x1; A.y1 = y1;
    A.sample0 = s0; A.nsamples = ns; A.hit_sample = 0u;
    A.tile_log2 = ctx->tile_log2; A.tile_ranks = ctx->tile_log2 ? ctx->nranks : 1; A.tile_rank = ctx->rank;
    A.accum = ctx->accum; A.hit = ctx->hit;
    A.hit64 = ctx->sp.debug_hits ? ctx->hit64 : nullptr;
    A.counters = ctx->d_counters;
    A.work_counter = ctx->d_work;
    A.hard = ctx->sp.hard_rays ? (HardRay*)ctx->hard_buf : nullptr;
    A.tubes = ctx->tube_seg; A.n_tubes = ctx->n_tubes; A.tube_tiles = ctx->tube_tiles; A.tube_tx = ctx->tube_tx;
    A.pixel_list = ctx->pixel_list;
    const double er[3] = {A.cam.eye[0] - A.sp.pos[0], A.cam.eye[1] - A.sp.pos[1], A.cam.eye[2] - A.sp.pos[2]};
    const double lr[3] = {A.sp.light_pos[0] - A.sp.pos[0], A.sp.light_pos[1] - A.sp.pos[1], A.sp.light_pos[2] - A.sp.pos[2]};
    to_body(A.sp, er, A.eye_b);
    to_body(A.sp, lr, A.light_b);
    A.defer_list = ctx->defer_list;
    A.defer_mask = ctx->defer_mask;
    A.rad = nullptr; A.rays = nullptr; A.hits = nullptr; A.srays = nullptr; A.sitem = nullptr; A.defer_items = nullptr;
    A.wave_p0 = 0; A.wave_np = 0; A.lvl_primary = 0; A.lvl_shadow = 0;
    A.defer_stats = ctx->d_defer_stats;
    A.K = make_fast_consts(ctx->hf, ctx->sp.radius);
    A.inv_rs = 1.0f / ctx->hf.radius_scale;
    A.g_log2 = 0;
    A.sq_rays = nullptr; A.sq_aux = nullptr; A.sq_cap = 0; A.sq_level = 0; A.accfix = ctx->accfix;
    A.beam_s = nullptr; A.beam_l = nullptr; A.beam_drop = (int)ctx->sp.beam_drop;
    A.list_cap = (unsigned)((size_t)ctx->width * ctx->height);
}

static int launch_cull(mrtx_ctx* ctx, const RenderArgs& A) {
    MRTX_CUDA(cudaMemsetAsync(A.work_counter, 0, 16 * sizeof(unsigned), ctx->stream));
    const unsigned tiles_x = (unsigned)(A.x1 - A.x0 + 7) / 8u, tiles_y = (unsigned)(A.y1 - A.y0 + 3) / 4u;
    const unsigned total = tiles_x * tiles_y * 32u;
    cull_kernel<<<(total + 255u) / 256u, 256, 0, ctx->stream>>>(A);
    return MRTX_OK;
}

// trace_alt.cu
int launch_trace_alt(mrtx_ctx* ctx, int x0, int y0, int x1, int y1, unsigned s0, unsigned ns, unsigned kernel);
