// Core of the ray / height-field intersection (K5, K6): float32 pyramid traversal + float64
// exact patch test.  Everything here is __host__ __device__ so that tools/trace_host.cu can run
// the very same code on the CPU for debugging against the oracle; the product only ever calls
// it from the kernels in trace.cu.
#pragma once

#include "common.cuh"

#define MRTX_HD __host__ __device__

#ifdef __CUDA_ARCH__
#define MRTX_LDG(p) __ldg(p)
__device__ __forceinline__ float mrtx_fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float mrtx_fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float mrtx_fdiv(float a, float b) { return __fdiv_rn(a, b); }
#else
#define MRTX_LDG(p) (*(p))
static inline float mrtx_fmul(float a, float b) { volatile float r = a * b; return r; }
static inline float mrtx_fadd(float a, float b) { volatile float r = a + b; return r; }
static inline float mrtx_fdiv(float a, float b) { volatile float r = a / b; return r; }
#endif

namespace mrtx_core {

// float32 traversal arithmetic is protected by margins, so approximate division is enough there
#ifdef __CUDA_ARCH__
__device__ __forceinline__ float fdiv_fast(float a, float b) { return __fdividef(a, b); }
#else
static inline float fdiv_fast(float a, float b) { return a / b; }
#endif

static int g_debug = 0;   // host-side debugging only (tools/trace_host.cu)
#ifndef __CUDA_ARCH__
#define MRTX_DBG(...) do { if (g_debug) printf(__VA_ARGS__); } while (0)
#else
#define MRTX_DBG(...) do {} while (0)
#endif

constexpr double PI_D = 3.14159265358979323846;
constexpr float  PI_F = 3.14159265358979323846f;
constexpr int    MAX_STEPS = 60000;

struct Counters { unsigned nodes, tests, overflow; };

// ---- height field access ---------------------------------------------------------------
template <bool I16>
MRTX_HD inline float texel_D(const HeightField& hf, int r, int c) {
    if (I16) {
        const float v = (float)MRTX_LDG((const int16_t*)hf.base + (size_t)r * hf.W + c);
        // exactly data_loader.py:219-242: *scale, +1, /radius_scale, one rounding each
        return mrtx_fdiv(mrtx_fadd(mrtx_fmul(v, hf.scale), 1.0f), hf.radius_scale);
    }
    return MRTX_LDG((const float*)hf.base + (size_t)r * hf.W + c);
}

template <bool I16>
MRTX_HD inline float level_D(const HeightField& hf, int L, int J, int I) {
    if (I16) {
        const float v = (float)MRTX_LDG((const int16_t*)hf.level[L] + (size_t)J * hf.nx[L] + I);
        return mrtx_fdiv(mrtx_fadd(mrtx_fmul(v, hf.scale), 1.0f), hf.radius_scale);
    }
    return MRTX_LDG((const float*)hf.level[L] + (size_t)J * hf.nx[L] + I);
}

struct Patch { int r0, c0; float d00, d01, d10, d11; };

template <bool I16>
MRTX_HD inline void load_patch(const HeightField& hf, int r0, int c0, Patch& P) {
    const int c1 = c0 + 1 == hf.W ? 0 : c0 + 1;
    P.r0 = r0; P.c0 = c0;
    P.d00 = texel_D<I16>(hf, r0, c0);     P.d01 = texel_D<I16>(hf, r0, c1);
    P.d10 = texel_D<I16>(hf, r0 + 1, c0); P.d11 = texel_D<I16>(hf, r0 + 1, c1);
}

// ---- float64 exact patch test -------------------------------------------------------------
struct Ray64 { double ox, oy, oz, dx, dy, dz, oo, od; };

struct HitInfo { double s, r, lon, lat, fc, fr; };

// The four walls of a level-0 cell, from the tables: west/east (cos, sin) of the wall longitude,
// north/south (sin, cos) of the wall latitude.  Row index r0 / r0+1 always exists in the table
// (it is the latitude of a texel-centre row) even where a polar cap has no wall.
struct Cell64 { double wcs, wsn, ecs, esn, nk, nc, sk; bool has_n, has_s; };

MRTX_HD inline void load_cell64(const HeightField& hf, const Patch& P, Cell64& C) {
    const double2 w = MRTX_LDG(hf.lon64 + P.c0), e = MRTX_LDG(hf.lon64 + P.c0 + 1);
    const double2 n = MRTX_LDG(hf.lat64 + P.r0), s = MRTX_LDG(hf.lat64 + P.r0 + 1);
    C.wcs = w.x; C.wsn = w.y; C.ecs = e.x; C.esn = e.y;
    C.nk = n.x; C.nc = n.y; C.sk = s.x;
    C.has_n = P.r0 > 0; C.has_s = P.r0 < hf.H - 2;
}

// atan2(y, x) for x > 0; the angles met here are at most one texel wide, so the odd series is
// exact to double rounding (|y/x| < 1/8: next term < 3e-18 relative).  Maps coarser than 64 x 32
// texels take the library call, kept out of line so it does not bloat the hot loop.
#ifdef __CUDA_ARCH__
__device__ __noinline__ double atan_wide(double y, double x) { return atan2(y, x); }
#else
static double atan_wide(double y, double x) { return atan2(y, x); }
#endif

MRTX_HD inline double atan_small(double y, double x) {
    const double q = y / x, q2 = q * q;
    if (q2 < 0.015625) {
        double p = 1.0 / 17.0;
        p = fma(p, q2, -1.0 / 15.0); p = fma(p, q2, 1.0 / 13.0); p = fma(p, q2, -1.0 / 11.0);
        p = fma(p, q2, 1.0 / 9.0);   p = fma(p, q2, -1.0 / 7.0);  p = fma(p, q2, 1.0 / 5.0);
        p = fma(p, q2, -1.0 / 3.0);  p = fma(p, q2, 1.0);
        return q * p;
    }
    return atan_wide(y, x);
}

// f(s) = |p(s)| - R * bilinear(fc, fr): fc, fr = texel fractions measured from the cell's west
// wall and from latitude row r0 as SMALL angles (no full-circle atan2, no cancellation)
MRTX_HD inline double patch_f(const Ray64& R, const Patch& P, const Cell64& C, int W, int H, double radius, double s,
                              HitInfo& info) {
    const double x = fma(s, R.dx, R.ox), y = fma(s, R.dy, R.oy), z = fma(s, R.dz, R.oz);
    const double rho = sqrt(x * x + y * y);
    const double r = sqrt(rho * rho + z * z);
    const double fc = atan_small(x * C.wcs + y * C.wsn, x * C.wsn - y * C.wcs) * (0.5 / PI_D) * W;
    const double fr_raw = -atan_small(z * C.nc - rho * C.nk, rho * C.nc + z * C.nk) * (1.0 / PI_D) * H;
    const double fr = fr_raw < 0.0 ? 0.0 : (fr_raw > 1.0 ? 1.0 : fr_raw);        // rows clamp (renderer_navigation.py:584)
    const double top = fma(fc, (double)P.d01 - (double)P.d00, (double)P.d00);
    const double bot = fma(fc, (double)P.d11 - (double)P.d10, (double)P.d10);
    const double d = fma(fr, bot - top, top);
    info.s = s; info.r = r; info.fc = fc; info.fr = fr_raw;
    return fma(-radius, d, r);
}

MRTX_HD inline void finish_info(const Patch& P, int W, int H, HitInfo& info) {
    info.lon = ((P.c0 + 0.5 + info.fc) / W - 0.5) * (2.0 * PI_D);
    info.lat = (0.5 - (P.r0 + 0.5 + info.fr) / H) * PI_D;
}

// First root of f on [a, b] inside one patch (f is close to a parabola there: a bilinear patch
// along a nearly straight (u, v) path): parabola through the ends and the middle, then Illinois
// regula falsi on the bracket until the step is below 1e-10 R.
// Written as an ITERATOR - root_step() performs exactly one evaluation of f - so that the render
// kernel can run one warp-uniform loop in which every lane that still needs an evaluation takes
// it at the same instruction (patch_f is ~150 float64 instructions; lanes need 4-9 of them).
struct RootIter {
    double a, m, h, fa, fb, fm, c1, c2, tv, lo, flo, hi, fhi, xs;
    int stage, side;
};

MRTX_HD inline void root_begin(RootIter& q, double a, double b) {
    q.a = a; q.m = 0.5 * (a + b); q.h = 0.5 * (b - a);
    q.fa = q.fb = q.fm = q.c1 = q.c2 = q.tv = q.flo = q.fhi = 0.0;
    q.lo = a; q.hi = b; q.xs = a; q.stage = 0; q.side = 0;
}

// 0 = needs another evaluation, 1 = root found (info describes it), 2 = no root on this piece
MRTX_HD inline int root_step(RootIter& q, const Ray64& R, const Patch& P, const Cell64& C, int W, int H, double radius,
                             HitInfo& info) {
    const double tol = 1.0e-10 * radius;
    const double fx = patch_f(R, P, C, W, H, radius, q.xs, info);
    bool guess = false;
    if (q.stage == 0) {                         // f(a)
        q.fa = fx;
        if (fx <= 0.0) return 1;                // entered below the surface: root at a
        q.xs = q.m + q.h; q.stage = 1;
    } else if (q.stage == 1) {                  // f(b)
        q.fb = fx; q.xs = q.m; q.stage = 2;
    } else if (q.stage == 2) {                  // f(m): f ~ fm + c1 t + c2 t^2, t = s - m
        q.fm = fx;
        q.c2 = (q.fa - 2.0 * fx + q.fb) / (2.0 * q.h * q.h); q.c1 = (q.fb - q.fa) / (2.0 * q.h);
        q.flo = q.fa;
        if (fx <= 0.0) { q.hi = q.m; q.fhi = fx; guess = true; }
        else if (q.fb <= 0.0) { q.lo = q.m; q.flo = fx; q.fhi = q.fb; guess = true; }
        else {
            // no sign change at the three samples: a grazing double root shows up as a dip
            if (!(q.c2 > 0.0)) return 2;
            q.tv = -q.c1 / (2.0 * q.c2);
            if (!(q.tv > -q.h && q.tv < q.h)) return 2;
            if (fx - q.c1 * q.c1 / (4.0 * q.c2) > 0.25 * fmin(fx, fmin(q.fa, q.fb))) return 2;   // dip clear of zero
            q.xs = q.m + q.tv; q.stage = 3;
        }
    } else if (q.stage == 3) {                  // f at the parabola's vertex
        if (fx > 0.0) return 2;
        if (q.tv > 0.0) { q.lo = q.m; q.flo = q.fm; }
        q.hi = q.m + q.tv; q.fhi = fx; guess = true;
    } else if (q.stage == 4) {                  // Illinois iterations
        if (fx > 0.0) { q.lo = q.xs; q.flo = fx; if (q.side == 1) q.fhi *= 0.5; q.side = 1; }
        else { q.hi = q.xs; q.fhi = fx; if (q.side == -1) q.flo *= 0.5; q.side = -1; }
        double nx = q.lo + (q.hi - q.lo) * q.flo / (q.flo - q.fhi);
        if (!(nx > q.lo && nx < q.hi)) nx = 0.5 * (q.lo + q.hi);
        if (q.hi - q.lo < tol || (fabs(nx - q.xs) < 0.25 * tol && fx <= 0.0) || ++q.stage > 64 + 4) {
            if (q.xs == q.hi) return 1;         // info already describes the root
            q.xs = q.hi; q.stage = 100;
        } else { q.xs = nx; q.stage = 4; }
    } else if (q.stage >= 100) {                // final evaluation at the root
        return 1;
    } else {                                    // stages 5..68 count Illinois iterations
        q.stage = 4;
    }
    if (guess) {
        // first guess: the parabola's root inside the bracket
        q.xs = 0.5 * (q.lo + q.hi);
        const double disc = q.c1 * q.c1 - 4.0 * q.c2 * q.fm;
        if (disc >= 0.0) {
            const double sq = sqrt(disc);
            const double t = -0.5 * (q.c1 + (q.c1 >= 0.0 ? sq : -sq));
            const double t1 = q.c2 != 0.0 ? t / q.c2 : 2.0 * q.h, t2 = t != 0.0 ? q.fm / t : 2.0 * q.h;
            const double s1 = q.m + t1, s2 = q.m + t2;
            const bool in1 = s1 > q.lo && s1 < q.hi, in2 = s2 > q.lo && s2 < q.hi;
            if (in1) q.xs = s1;
            if (in2 && (!in1 || s2 < s1)) q.xs = s2;
        }
        q.stage = 4;
    }
    return 0;
}

// The exact interval(s) of the ray inside cell (r0, c0): all wall crossings inside the generous
// window [wa, wb] split it into pieces; the pieces inside the cell that touch [oa, ob] (where the
// float32 traversal believes the ray crosses the cell) are the ones searched for a root, in order
// and in full - so the searched intervals of consecutive cells tile the ray exactly even where
// the float32 crossing parameters are ill-conditioned (ray nearly tangent to a wall).
// Cell walls: lon half-planes g = p.t (t = (cos lam, sin lam, 0)), lat cones h = z - k r (k = sin phi).
// next_piece() returns the k-th such piece (k counted from `from`), its start wall id in *cid
// (-1 = window edge) and the index to continue from.
MRTX_HD inline bool next_piece(const Ray64& R, const Cell64& C, double wa, double wb, double oa, double ob,
                               int& from, double& pa, double& pb, int& pcid) {
    double crit[10];
    int cid[10];
    int n = 0;
    crit[n] = wa; cid[n++] = -1;
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        const double cs = side ? C.ecs : C.wcs, sn = side ? C.esn : C.wsn;
        const double g0 = R.ox * cs + R.oy * sn, g1 = R.dx * cs + R.dy * sn;
        if (g1 != 0.0) {
            const double sc = -g0 / g1;
            if (sc > wa && sc < wb) { crit[n] = sc; cid[n++] = side; }
        }
    }
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        if (!(side ? C.has_s : C.has_n)) continue;
        const double k = side ? C.sk : C.nk;
        const double k2 = k * k;
        const double A = R.dz * R.dz - k2, B = R.oz * R.dz - k2 * R.od, Cq = R.oz * R.oz - k2 * R.oo;
        double r1 = wa, r2 = wa;                               // "not inside the window"
        if (fabs(A) < 1e-300) { if (B != 0.0) r1 = -Cq / (2.0 * B); }
        else {
            const double disc = B * B - A * Cq;
            if (disc >= 0.0) {
                const double q = -(B + (B >= 0.0 ? 1.0 : -1.0) * sqrt(disc));
                r1 = q / A;
                if (q != 0.0) r2 = Cq / q;
            }
        }
        if (r1 > wa && r1 < wb) { crit[n] = r1; cid[n++] = 2 + side; }
        if (r2 > wa && r2 < wb) { crit[n] = r2; cid[n++] = 2 + side; }
    }
    crit[n] = wb; cid[n++] = -1;
    for (int i = 1; i < n; ++i) {                                // insertion sort, n <= 8
        const double key = crit[i];
        const int kid = cid[i];
        int j = i - 1;
        while (j >= 0 && crit[j] > key) { crit[j + 1] = crit[j]; cid[j + 1] = cid[j]; --j; }
        crit[j + 1] = key; cid[j + 1] = kid;
    }
    for (int i = from; i + 1 < n; ++i) {
        const double a = crit[i], b = crit[i + 1];
        MRTX_DBG("   piece [%.9f, %.9f] overlap [%.9f, %.9f]\n", a, b, oa, ob);
        if (!(b > a) || b < oa || a > ob) continue;
        const double m = 0.5 * (a + b);
        const double x = R.ox + m * R.dx, y = R.oy + m * R.dy, z = R.oz + m * R.dz;
        if (x * C.wcs + y * C.wsn < 0.0) continue;               // west of the cell
        if (x * C.ecs + y * C.esn > 0.0) continue;               // east of it
        const double r = sqrt(x * x + y * y + z * z);
        if (C.has_n && z - C.nk * r > 0.0) continue;             // north of it
        if (C.has_s && z - C.sk * r < 0.0) continue;             // south of it
        pa = a; pb = b; pcid = cid[i]; from = i + 1;
        return true;
    }
    return false;
}

// ---- float32 pyramid traversal ---------------------------------------------------------------
struct Trav {
    float ox, oy, oz, dx, dy, dz, oo, od;    // re-based ray
    float smax;
};

MRTX_HD inline float ray_r2(const Trav& T, float s) { return fmaf(s, fmaf(2.0f, T.od, s), T.oo); }

// Exit parameter + face (0 lon-lo, 1 lon-hi, 2 north, 3 south, 4 end of ray) of the level-L cell
// (J, I) for a ray currently at s.  Every wall has a signed "outsideness" q(s) (> 0 beyond the
// wall).  A point within `tol` of a wall counts as ON it and the direction of travel decides:
// moving outward -> leave now, moving inward -> stay.  Both cells that share a wall evaluate the
// same q with opposite sign, so a ray handed across a wall can never be handed straight back.
MRTX_HD inline float cell_exit32(const HeightField& hf, const Trav& T, int L, int J, int I, float s, int& face) {
    const int W = hf.W, H = hf.H;
    const int a = I << L, b = min((I + 1) << L, W);
    const int n = J << L, m = min((J + 1) << L, H - 1);
    float best = T.smax;
    face = 4;
    const float x = fmaf(s, T.dx, T.ox), y = fmaf(s, T.dy, T.oy), z = fmaf(s, T.dz, T.oz);
    const float r = sqrtf(fmaxf(ray_r2(T, s), 1e-30f));
    // "on the wall" = within the float32 error of the wall function: 2e-7 r from the re-based ray plus the
    // rounding of the function's two cancelling terms (a uniform 2e-6 r is 3 % of a full-resolution cell and
    // sends rays that run nearly parallel to a wall into the neighbouring row tens of cells early)
    const float tol0 = 2.0e-7f * r;
#pragma unroll
    for (int side = 0; side < 2; ++side) {      // 0 = west wall (index a), 1 = east wall (index b)
        const float2 wl = MRTX_LDG(hf.lon32 + (side ? b : a));
        const float cs = wl.x, sn = wl.y;
        const float sg = side ? 1.0f : -1.0f;                   // outward = +g east, -g west
        const float tol = fmaf(6.0e-7f, fabsf(x * cs) + fabsf(y * sn), tol0);
        const float q = sg * (x * cs + y * sn), dq = sg * (T.dx * cs + T.dy * sn);
        if (!(dq > 0.0f)) continue;                             // not heading out through this wall
        float sc;
        if (q >= -tol) sc = s;                                  // on or beyond the wall, moving outward
        else {
            sc = s - fdiv_fast(q, dq);
            const float px = fmaf(sc, T.dx, T.ox), py = fmaf(sc, T.dy, T.oy);
            if (!(px * sn - py * cs > 0.0f)) continue;          // crosses the opposite half-plane
        }
        if (sc < best) { best = sc; face = side; }
    }
#pragma unroll
    for (int side = 0; side < 2; ++side) {      // 0 = north wall (index n), 1 = south wall (index m)
        if (side == 0 ? (n == 0) : (m == H - 1)) continue;      // polar caps have no wall
        const float2 kk = MRTX_LDG(hf.latsc32 + (side == 0 ? n : m));           // (sin, cos)(lat) of the wall
        const float k = kk.x;
        const bool polar = fabsf(k) > 0.70710678f;
        const float sg = side ? -1.0f : 1.0f;                   // outward = north of the north wall, south of the south wall
        // Signed distance north of the wall.  z - k r cancels two numbers ~R and shows the distance only as
        // r cos(lat) d(lat): beyond 45 deg the same cone is tested as rho = kc r (rho = distance from the axis).
        const float rho = sqrtf(x * x + y * y);
        const float side_n = !polar ? z - k * r : (k > 0.0f ? (z > 0.0f ? kk.y * r - rho : -r) : (z < 0.0f ? rho - kk.y * r : r));
        const float tol = fmaf(6.0e-7f * r, fminf(fabsf(k), kk.y), tol0);
        const float q = sg * side_n, dq = sg * (T.dz * r - k * (T.od + s));  // dq: sign of d(q)/ds (times r > 0)
        if (q >= -tol && dq > 0.0f) {                           // on or beyond the wall, moving outward
            if (s < best) { best = s; face = 2 + side; }
            continue;
        }
        // outward crossings ahead: roots of F(s) = z^2 - k^2 r^2 = A s^2 + 2 B s + C on the wall's nappe
        // (polar walls: -F = rho^2 - kc^2 r^2, the same roots from small numbers).
        // On that nappe h = F / (z + k r) and z + k r has the sign of k, so the crossing direction
        // d(h)/ds has the sign of k * F'(s) = k * 2 (A s + B): no square root or division needed.
        float A, B, Cq;
        if (polar) {
            const float c2 = kk.y * kk.y;
            A = -(T.dx * T.dx + T.dy * T.dy - c2); B = -(T.ox * T.dx + T.oy * T.dy - c2 * T.od); Cq = -(T.ox * T.ox + T.oy * T.oy - c2 * T.oo);
        } else {
            const float k2 = k * k;
            A = T.dz * T.dz - k2; B = T.oz * T.dz - k2 * T.od; Cq = T.oz * T.oz - k2 * T.oo;
        }
        float r1 = -1.0f, r2 = -1.0f;
        if (fabsf(A) < 1e-12f) { if (B != 0.0f) r1 = fdiv_fast(-Cq, 2.0f * B); }
        else {
            const float disc = B * B - A * Cq;
            if (disc >= 0.0f) {
                const float qq = -(B + copysignf(sqrtf(disc), B));
                r1 = fdiv_fast(qq, A);
                if (qq != 0.0f) r2 = fdiv_fast(Cq, qq);
            }
        }
        const float ksg = k >= 0.0f ? sg : -sg;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float sc = i ? r2 : r1;
            if (!(sc > s) || !(sc < best)) continue;
            const float zc = fmaf(sc, T.dz, T.oz);
            if (k != 0.0f && zc * k < 0.0f) continue;            // the cone's other nappe
            const float dF = fmaf(A, sc, B);
            if ((k != 0.0f ? ksg * dF : sg * T.dz) > 0.0f) { best = sc; face = 2 + side; }
        }
    }
    return best;
}

struct TraceOut { bool hit; double s; HitInfo info; Patch patch; };

// Traversal state of one ray (float32, re-based at the bounding-sphere entry s_in).
struct TravState {
    Trav T;
    float s;                 // current parameter (relative to s_in)
    int L, J, I;             // current cell
    int steps;
    double s_in, s_end, s_min;
};

enum { TR_CONTINUE = 0, TR_CANDIDATE = 1, TR_END = 2 };

// Clip the ray to the bounding sphere R * dmax and find its first cell.  false = misses the Moon.
MRTX_HD inline bool trav_begin(const HeightField& hf, double radius, const Ray64& R, double s_min, int start_level,
                               TravState& st) {
    const double Rb = radius * (double)hf.dmax;
    const double disc = R.od * R.od - (R.oo - Rb * Rb);
    if (disc < 0.0) return false;
    const double sq = sqrt(disc);
    const double s_end = -R.od + sq;
    if (s_end <= s_min) return false;
    const double s_in = fmax(s_min, -R.od - sq);
    st.s_in = s_in; st.s_end = s_end; st.s_min = s_min;
    Trav& T = st.T;
    const double bx = R.ox + s_in * R.dx, by = R.oy + s_in * R.dy, bz = R.oz + s_in * R.dz;
    T.ox = (float)bx; T.oy = (float)by; T.oz = (float)bz;
    T.dx = (float)R.dx; T.dy = (float)R.dy; T.dz = (float)R.dz;
    T.oo = T.ox * T.ox + T.oy * T.oy + T.oz * T.oz;
    T.od = T.ox * T.dx + T.oy * T.dy + T.oz * T.dz;
    T.smax = (float)(s_end - s_in);
    const int W = hf.W, H = hf.H;
    const int L = min(max(start_level, 0), hf.top);
    // first cell from the position just inside (a wrong neighbour is corrected by the exit rules)
    const float t0 = fminf(1e-5f * (float)radius, 0.5f * T.smax);
    const float x = fmaf(t0, T.dx, T.ox), y = fmaf(t0, T.dy, T.oy), z = fmaf(t0, T.dz, T.oz);
    const float lon = atan2f(x, -y), lat = atan2f(z, sqrtf(x * x + y * y));
    const float u = (lon * (0.5f / PI_F) + 0.5f) * (float)W - 0.5f, v = (0.5f - lat * (1.0f / PI_F)) * (float)H - 0.5f;
    int c0 = (int)floorf(u);
    c0 = c0 < 0 ? c0 + W : (c0 >= W ? c0 - W : c0);
    const int r0 = min(max((int)floorf(v), 0), H - 2);
    st.L = L; st.J = r0 >> L; st.I = c0 >> L;
    st.s = 0.0f; st.steps = 0;
    return true;
}

// Hand the ray to the neighbour across `face` (ascending one level when a parent wall is crossed).
// false = the ray has left the bounding sphere.
MRTX_HD inline bool trav_advance(const HeightField& hf, TravState& st, float sx, int face) {
    if (face == 4) return false;
    st.s = sx;
    int L = st.L, J = st.J, I = st.I;
    bool up;
    if (face == 1)      { I += 1; if (I >= hf.nx[L]) I = 0; up = (I & 1) == 0; }
    else if (face == 0) { up = (I & 1) == 0; I -= 1; if (I < 0) I = hf.nx[L] - 1; }
    else if (face == 3) { J += 1; up = (J & 1) == 0; }
    else                { up = (J & 1) == 0; J -= 1; }
    if (J < 0 || J >= hf.ny[L]) return false;                // cannot happen (caps have no wall); be safe
    if (up && L < hf.top) { L += 1; I >>= 1; J >>= 1; }
    st.L = L; st.J = J; st.I = I;
    return true;
}

// One node visit.  TR_CANDIDATE: the current cell is a level-0 patch the ray may hit - P, sx, face
// describe it and the caller runs exact_test() then trav_advance().  TR_CONTINUE: skipped, advanced
// or descended.  TR_END: left the sphere (or ran out of steps: counted in cnt.overflow).
template <bool I16>
MRTX_HD inline int trav_step(const HeightField& hf, float Rf, TravState& st, Patch& P, float& sx_out, int& face_out,
                             Counters& cnt) {
    if (++st.steps > MAX_STEPS) { ++cnt.overflow; return TR_END; }
    const Trav& T = st.T;
    const int L = st.L, J = st.J, I = st.I;
    const float s = st.s;
    int face;
    const float sx = fmaxf(cell_exit32(hf, T, L, J, I, s, face), s);
    ++cnt.nodes;
    float dmax;                                              // max of the surface over this cell
    if (L == 0) {
        load_patch<I16>(hf, J, I, P);
        dmax = fmaxf(fmaxf(P.d00, P.d01), fmaxf(P.d10, P.d11));
    } else {
        dmax = level_D<I16>(hf, L, J, I);
    }
    const float marg = 2.0e-6f * Rf;                         // > float32 error of a radius near R
    const float rc = fmaf(Rf, dmax, marg);
    MRTX_DBG("step %d L%d J%d I%d s=%.7f sx=%.7f face=%d dmax=%.7f r(s)=%.7f\n", st.steps, L, J, I, s, sx, face, dmax, sqrtf(ray_r2(T, s)));
    // min radius of the ray over [s, sx] (with a little slack either side)
    const float pad = 4.0e-6f * Rf;
    const float ta = fmaxf(s - pad, 0.0f), tb = fminf(sx + pad, T.smax);
    const float tm = fminf(fmaxf(-T.od, ta), tb);
    if (ray_r2(T, tm) <= rc * rc) {
        if (L == 0) { sx_out = sx; face_out = face; return TR_CANDIDATE; }
        // move up to where the ray enters the cell's shell, then pick the child there
        float sd = s;
        if (ray_r2(T, s) > rc * rc) {
            const float dq = T.od * T.od - (T.oo - rc * rc);
            if (dq > 0.0f) sd = fminf(fmaxf(-T.od - sqrtf(dq), s), sx);
        }
        const float x = fmaf(sd, T.dx, T.ox), y = fmaf(sd, T.dy, T.oy), z = fmaf(sd, T.dz, T.oz);
        const int mi = (2 * I + 1) << (L - 1), mj = (2 * J + 1) << (L - 1);
        int ci = 2 * I, cj = 2 * J;
        if (mi < min((I + 1) << L, hf.W)) {
            const float2 wl = MRTX_LDG(hf.lon32 + mi);
            if (x * wl.x + y * wl.y >= 0.0f) ci += 1;
        }
        if (mj < min((J + 1) << L, hf.H - 1)) {
            const float2 kk = MRTX_LDG(hf.latsc32 + mj);
            const float rho = sqrtf(x * x + y * y), rr = sqrtf(x * x + y * y + z * z);
            const float side_n = fabsf(kk.x) <= 0.70710678f ? z - kk.x * rr
                               : (kk.x > 0.0f ? (z > 0.0f ? kk.y * rr - rho : -rr) : (z < 0.0f ? rho - kk.y * rr : rr));
            if (side_n < 0.0f) cj += 1;                              // south of the mid wall
        }
        st.s = sd; st.L = L - 1; st.I = ci; st.J = cj;
        return TR_CONTINUE;
    }
    return trav_advance(hf, st, sx, face) ? TR_CONTINUE : TR_END;
}

// Exact float64 test of a candidate patch (window [st.s, sx] of the float32 walk), as a state
// machine: exact_begin() sets it up, exact_step() spends ONE evaluation of f and returns true while
// more are needed; afterwards X.found tells whether the ray hits (X.info, X.P describe the hit).
// If the ray turns out to ENTER a cell already below the surface, the first crossing lies in a cell
// the float32 walk skipped (ray within rounding of a wall or a corner): the machine then walks back
// through the neighbours across the entry walls until it is found.
struct ExactState {
    Patch P;
    Cell64 C;
    RootIter q;
    HitInfo info;            // scratch of the evaluation in flight
    HitInfo hit;             // the accepted root
    Patch hitP;
    double wa, wb, oa, ob;
    int from, pcid, back, found;
};

template <bool I16>
MRTX_HD inline bool exact_begin(const HeightField& hf, double radius, const Ray64& R, const TravState& st, const Patch& P,
                                float sx, ExactState& X, Counters& cnt) {
    const double big = 3.0e-3 * radius, small = 2.0e-5 * radius;
    const double sa = st.s_in + (double)st.s, sb = st.s_in + (double)sx;
    X.P = P;
    X.wa = fmax(sa - big, st.s_min); X.wb = fmin(sb + big, st.s_end); X.oa = sa - small; X.ob = sb + small;
    X.from = 0; X.back = 0; X.found = 0; X.pcid = -1;
    load_cell64(hf, X.P, X.C);
    ++cnt.tests;
    double pa, pb;
    if (!next_piece(R, X.C, X.wa, X.wb, X.oa, X.ob, X.from, pa, pb, X.pcid)) return false;
    root_begin(X.q, pa, pb);
    return true;
}

template <bool I16>
MRTX_HD inline bool exact_step(const HeightField& hf, double radius, const Ray64& R, const TravState& st, ExactState& X,
                               Counters& cnt) {
    const int r = root_step(X.q, R, X.P, X.C, hf.W, hf.H, radius, X.info);
    if (r == 0) return true;
    double pa, pb;
    if (r == 1) {
        X.found = 1; X.hit = X.info; X.hitP = X.P;
        if (!(X.info.s == X.q.a && X.pcid >= 0 && X.back < 16)) return false;     // a proper root: done
        // below the surface at the wall the ray came in through: look in the cell behind that wall
        const double big = 3.0e-3 * radius, small = 2.0e-5 * radius, tiny = 1.0e-9 * radius;
        int r0 = X.P.r0, c0 = X.P.c0;
        if (X.pcid == 0) c0 = c0 == 0 ? hf.W - 1 : c0 - 1;
        else if (X.pcid == 1) c0 = c0 + 1 == hf.W ? 0 : c0 + 1;
        else if (X.pcid == 2) r0 -= 1;
        else r0 += 1;
        if (r0 < 0 || r0 > hf.H - 2) return false;
        load_patch<I16>(hf, r0, c0, X.P);
        load_cell64(hf, X.P, X.C);
        ++cnt.tests; ++X.back;
        const double sh = X.info.s;
        X.wa = fmax(sh - big, st.s_min); X.wb = fmin(sh + small, st.s_end); X.oa = sh - small; X.ob = sh - tiny;
        X.from = 0;
    }
    // r == 2 (no root on this piece) or a walk-back: continue with the next piece, if any
    if (!next_piece(R, X.C, X.wa, X.wb, X.oa, X.ob, X.from, pa, pb, X.pcid)) return false;
    root_begin(X.q, pa, pb);
    return true;
}

MRTX_HD inline void exact_result(const HeightField& hf, const ExactState& X, TraceOut& out) {
    out.hit = true; out.info = X.hit; out.patch = X.hitP; out.s = X.hit.s;
    finish_info(out.patch, hf.W, hf.H, out.info);
}

// sequential form (host tool, per-pixel kernel)
template <bool I16>
MRTX_HD inline bool exact_test(const HeightField& hf, double radius, const Ray64& R, const TravState& st, const Patch& P,
                               float sx, TraceOut& out, Counters& cnt) {
    ExactState X;
    bool run = exact_begin<I16>(hf, radius, R, st, P, sx, X, cnt);
    while (run) run = exact_step<I16>(hf, radius, R, st, X, cnt);
    if (!X.found) return false;
    exact_result(hf, X, out);
    return true;
}

// First intersection of the body-frame ray for s >= s_min (sequential form of the three pieces
// above; the render kernel interleaves them across a warp).  any_hit: stop at any intersection.
template <bool I16>
MRTX_HD void trace_ray(const HeightField& hf, double radius, const Ray64& R, double s_min, bool any_hit,
                       int start_level, TraceOut& out, Counters& cnt) {
    out.hit = false;
    TravState st;
    if (!trav_begin(hf, radius, R, s_min, start_level, st)) return;
    const float Rf = (float)radius;
    for (;;) {
        Patch P;
        float sx;
        int face;
        const int r = trav_step<I16>(hf, Rf, st, P, sx, face, cnt);
        if (r == TR_END) return;
        if (r == TR_CANDIDATE) {
            if (exact_test<I16>(hf, radius, R, st, P, sx, out, cnt)) return;
            if (!trav_advance(hf, st, sx, face)) return;
        }
    }
}

}  // namespace mrtx_core
