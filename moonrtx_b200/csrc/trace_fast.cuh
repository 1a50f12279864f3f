// Filtered fast path of the ray / height-field intersection (K5, K6).
//
// trace_core.cuh decides every candidate patch with ~2 500 float64 instructions.  Almost all
// of that precision is spent on cancellation: f(s) = |p(s)| - R*D(u, v) subtracts two numbers
// near R = 10 to find a root to 1e-10 R.  Here the same function is evaluated in a frame in which
// nothing is large:
//
//   * origin on the sphere of radius R*D00 at the patch's north-west corner (lambda_w, phi_n),
//     axes east / north / up there.  The ray is moved into that frame ONCE per candidate in
//     float64 (about 25 DFMA, walls from the same float64 tables the exact test uses) and is
//     then a float32 line  (a, b, c)(t) = (a0, b0, c0) + t (da, db, dc)  whose coordinates are of
//     the size of one cell;
//   * longitude / latitude offsets from the corner are small-angle arctangents of ratios of
//     those coordinates, the height above the corner sphere is c + (a^2 + b^2) / (r + R + c):
//     no cancellation anywhere, so float32 keeps ~1e-10 R in f where the global form keeps 1e-7 R.
//
// The result is a FILTER, not a replacement: every decision carries a margin (FastConsts::mg), and
// whatever falls inside a margin - grazing double roots, a ray that enters a cell already below
// the surface, polar-cap rows, ill-conditioned roots - returns FT_DEFER.  Deferred samples are
// re-traced from scratch by the exact kernel, so the rendered frame never depends on a float32
// decision that float64 could have made differently.
#pragma once

#include "trace_core.cuh"

namespace mrtx_core {

enum { FT_MISS = 0, FT_HIT = 1, FT_DEFER = 2 };
#if defined(MRTX_LEVEL_HIST) && !defined(__CUDA_ARCH__)
extern unsigned long long g_level_hist[32];
#endif
// A walk longer than SceneParams::long_walk nodes (default 2048) is a grazing ray among polar slivers (cells
// centimetres wide): tens of thousands of nodes, one dependent fetch after the other, in ONE lane.  It is handed to
// the referee (reason 15), whose warp cuts the ray into pieces and walks them side by side.
// fast_test() reports WHY it defers in the bits above the status (statistics only: mrtx_defer_stats)
#define FT_DEFER_R(reason) (FT_DEFER | ((reason) << 2))

#ifdef __CUDA_ARCH__
__device__ __forceinline__ float f_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float f_rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#else
static inline float f_rcp(float x) { return 1.0f / x; }
static inline float f_rsqrt(float x) { return 1.0f / sqrtf(x); }
#endif

#ifdef __CUDA_ARCH__
__device__ __forceinline__ float f_sqrt_fast(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float f_rcp_fast(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#else
static inline float f_sqrt_fast(float x) { return sqrtf(x); }
static inline float f_rcp_fast(float x) { return 1.0f / x; }
#endif

// 1/sqrt(x) in float64 from the float32 seed (two Newton steps: 1e-7 -> 1e-14 -> rounding)
MRTX_HD inline double d_rsqrt(double x) {
    double y = (double)f_rsqrt((float)x);
    y = y * fma(-0.5 * x, y * y, 1.5);
    y = y * fma(-0.5 * x, y * y, 1.5);
    return y;
}

#ifndef MRTX_PREFETCH
#define MRTX_PREFETCH 0
#endif
#ifndef MRTX_DESCENT2
#define MRTX_DESCENT2 0          // > 0: descend two levels at a time from cells of this level upwards (walk_step); measured at
                                 // config 3: node visits 12.2 -> 10.2 per ray, both walk kernels 2-10 % slower: off
#endif
// MRTX_TILED (common.cuh): walk the copy of the levels laid out in 8 x 8-cell tiles.  Measured at config 3: 14.88 / 12.06 ms
// (trace_kernel_fast / shadow_kernel) against 14.77 / 12.30 ms with rows - no difference, the walk is not waiting for cache
// lines; off by default (the tiled copy is another 2.8 GB).
// 32-bit read-only load of two neighbouring int16 cells, issued where it stands (volatile: the compiler must not sink it
// to its use, which would put the latency back on the critical path)
#ifdef __CUDA_ARCH__
__device__ __forceinline__ unsigned mrtx_ldg_u32(const int16_t* p) {
    unsigned v;
    asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
#else
static inline unsigned mrtx_ldg_u32(const int16_t* p) { unsigned v; memcpy(&v, p, 4); return v; }
#endif

struct FastConsts {
    float R;            // sphere radius
    float Kw, Kh;       // texels per radian: W / 2 pi, H / pi
    float mg;           // |f| below this is "on the surface": undecidable in float32
    float pad;          // float32 traversal parameter slack
    float epsc;         // a root this far outside the cell (in cells) still belongs to it
    float ds_tol;       // a root must be located to this (ray parameter) or the sample is deferred
    float sD, cD, tD;   // sin, cos of one cell in longitude (2 pi / W); tan of one cell in latitude (pi / H)
    float cell;         // size of a cell at the equator, scene units
    int   enabled;      // 0: map too coarse for the small-angle series (W < 360) -> always defer
};

MRTX_HD inline FastConsts make_fast_consts(const HeightField& hf, double radius) {
    FastConsts K;
    K.R = (float)radius;
    K.Kw = (float)(hf.W / (2.0 * PI_D));
    K.Kh = (float)(hf.H / PI_D);
    K.mg = (float)(2.0e-9 * radius);
    K.pad = (float)(4.0e-6 * radius);
    K.epsc = 2.0e-6f;
    K.sD = (float)sin(2.0 * PI_D / hf.W); K.cD = (float)cos(2.0 * PI_D / hf.W); K.tD = (float)tan(PI_D / hf.H);
    K.cell = (float)(2.0 * PI_D * radius / hf.W);
    K.ds_tol = (float)(5.0e-4 * 2.0 * PI_D * radius / hf.W);     // 5e-4 texel (the parity budget is 1e-3)
    K.enabled = hf.W >= 360 && hf.H >= 180;
    return K;
}

// Raw patch as the traversal loaded it: texel values before decoding (int16 counts as floats, or D).
struct RawPatch { int r0, c0; float v00, v01, v10, v11; };

template <bool I16>
MRTX_HD inline float decode_exact(const HeightField& hf, float v) {
    // exactly data_loader.py:219-242: *scale, +1, /radius_scale, one rounding each
    return I16 ? mrtx_fdiv(mrtx_fadd(mrtx_fmul(v, hf.scale), 1.0f), hf.radius_scale) : v;
}
// for the traversal's conservative bounds two roundings fewer are fine (covered by its margin)
template <bool I16>
MRTX_HD inline float decode_bound(const HeightField& hf, float v, float inv_rs) {
    return I16 ? fmaf(v, hf.scale, 1.0f) * inv_rs : v;
}

struct FastHit {
    double s;               // ray parameter of the hit
    float fc, fr;           // position in the patch, cells from the west wall / north row
    float d00, d01, d10, d11;
    int r0, c0;
};

// One evaluation of f in the local frame; also yields the in-cell coordinates.
struct LocalRay {
    float a0, b0, c0, da, db, dc;       // east, north, up (above the corner sphere R*D00)
    float nk, nc;                       // sin, cos of the north row's latitude
    float Rc0;                          // R * D00: radius of the corner sphere
    float e01, e10, exx;                // R * (D01-D00), R * (D10-D00), R * (D11-D10-D01+D00)
    float fr_lo, fr_hi;                 // polar-cap rows: the rows clamp (renderer_navigation.py:581-587), D does not depend
                                        // on latitude beyond the last texel centre; elsewhere -inf, +inf
};

MRTX_HD inline float local_f(const LocalRay& Q, const FastConsts& K, float t, float& fc, float& fr) {
    const float a = fmaf(t, Q.da, Q.a0), b = fmaf(t, Q.db, Q.b0), c = fmaf(t, Q.dc, Q.c0);
    const float Rc = Q.Rc0 + c;
    const float hh = fmaf(Q.nc, Rc, -Q.nk * b);                 // horizontal distance from the polar axis (towards lambda_w)
    const float q = a * f_rcp(hh), q2 = q * q;
    fc = K.Kw * q * fmaf(q2, fmaf(q2, 0.2f, -0.33333334f), 1.0f);           // atan(q) * W / 2 pi
    const float u = 0.5f * a * q * fmaf(-0.25f, q2, 1.0f);      // rho - hh = a^2 / (rho + hh)
    const float v = fmaf(-Q.nk, u, b) * f_rcp(fmaf(Q.nc, u, Rc)), v2 = v * v;
    fr = -K.Kh * v * fmaf(v2, fmaf(v2, 0.2f, -0.33333334f), 1.0f);          // -atan(v) * H / pi
    fr = fminf(fmaxf(fr, Q.fr_lo), Q.fr_hi);
    const float rr2 = fmaf(a, a, b * b);
    const float r = f_sqrt_fast(fmaf(Rc, Rc, rr2));
    const float hr = fmaf(rr2, f_rcp(r + Rc), c);               // |p| - R*D00
    const float dd = fmaf(fc, Q.e01, fr * fmaf(fc, Q.exx, Q.e10));          // R*(D(fc, fr) - D00)
    return hr - dd;
}

// Decide one candidate patch.  [ws, we]: the float32 traversal's window (relative to s_in);
// s_lo, s_in + smax: the extent of the ray.
//
// The traversal's idea of where the ray is inside this cell is only good to its wall tolerance (2e-6 R:
// 3 % of a cell of the full-resolution map), and its on-wall rule hands a ray to the next cell that early.
// The test therefore clips the ray to the cell ITSELF, in the local frame where the four walls are (to
// 1e-6 cell) linear in t, and searches from max(traversal start, true entry) to the TRUE exit: the intervals
// of consecutive cells tile the ray whatever the float32 walk believed, as next_piece() does for the exact
// test.
template <bool I16>
MRTX_HD inline void load_raw_patch(const HeightField& hf, int r0, int c0, RawPatch& P) {
    const int W = hf.W, c1 = c0 + 1 == W ? 0 : c0 + 1;
    P.r0 = r0; P.c0 = c0;
    if (I16) {
        const int16_t* b = (const int16_t*)hf.base + (size_t)r0 * W;
        P.v00 = (float)MRTX_LDG(b + c0); P.v01 = (float)MRTX_LDG(b + c1);
        P.v10 = (float)MRTX_LDG(b + W + c0); P.v11 = (float)MRTX_LDG(b + W + c1);
    } else {
        const float* b = (const float*)hf.base + (size_t)r0 * W;
        P.v00 = MRTX_LDG(b + c0); P.v01 = MRTX_LDG(b + c1);
        P.v10 = MRTX_LDG(b + W + c0); P.v11 = MRTX_LDG(b + W + c1);
    }
}

#ifndef MRTX_TEST_ATTR
#define MRTX_TEST_ATTR inline
#endif
template <bool I16>
MRTX_HD MRTX_TEST_ATTR int fast_test(const HeightField& hf, const FastConsts& K, const Ray64& R, double s_in, double s_lo,
                             float ws, float we, float smax, RawPatch P, bool any_hit, FastHit& out) {
    if (!K.enabled) return FT_DEFER_R(1);
    const double s_c = s_in + (double)ws;
    const float t_lo = (float)(s_lo - s_c), t_hi = smax - ws;
    float t_cap = INFINITY;                                     // walk-back: the neighbour is searched up to our entry
    bool from_entry = false;                                    // search from the cell's own entry, not from where the walk stands
#pragma unroll 1
    for (int back = 0; ; ++back) {
        // Polar-cap rows (r0 = 0, H - 2): the cell runs on to the pole, with the row coordinate clamped at the last
        // texel centre and no wall on that side.  Its longitude walls still are planes through the axis and its
        // latitude offsets still are small angles, so the local frame holds; only next to the axis itself
        // (a / hh -> 0 / 0) it does not, and that is deferred below.
        if (P.r0 < 0 || P.r0 > hf.H - 2) return FT_DEFER_R(1);
        const bool cap_n = P.r0 == 0, cap_s = P.r0 == hf.H - 2;
        const float d00 = decode_exact<I16>(hf, P.v00), d01 = decode_exact<I16>(hf, P.v01);
        const float d10 = decode_exact<I16>(hf, P.v10), d11 = decode_exact<I16>(hf, P.v11);
        const double2 w = MRTX_LDG(hf.lon64 + P.c0), n = MRTX_LDG(hf.lat64 + P.r0);   // (cos, sin) lon_w; (sin, cos) lat_n
        LocalRay Q;
        {
            const double px = fma(s_c, R.dx, R.ox), py = fma(s_c, R.dy, R.oy), pz = fma(s_c, R.dz, R.oz);
            const double a0 = px * w.x + py * w.y, m0 = px * w.y - py * w.x;
            const double da = R.dx * w.x + R.dy * w.y, dm = R.dx * w.y - R.dy * w.x;
            const double Rc0 = (double)K.R * (double)d00;
            Q.a0 = (float)a0; Q.da = (float)da;
            Q.b0 = (float)(n.y * pz - n.x * m0); Q.db = (float)(n.y * R.dz - n.x * dm);
            Q.c0 = (float)(fma(n.y, m0, n.x * pz) - Rc0); Q.dc = (float)fma(n.y, dm, n.x * R.dz);
            Q.nk = (float)n.x; Q.nc = (float)n.y; Q.Rc0 = (float)Rc0;
        }
        Q.e01 = K.R * (d01 - d00); Q.e10 = K.R * (d10 - d00); Q.exx = K.R * ((d11 - d10) - (d01 - d00));
        Q.fr_lo = cap_n ? 0.0f : -INFINITY; Q.fr_hi = cap_s ? 1.0f : INFINITY;

        // the ray inside the cell: four walls (west, east, north, south), each g(t) = g0 + t g1 >= 0 inside
        float t_in = -INFINITY, t_out = INFINITY;
        int wall_in = -1;
        {
            const float Rc = Q.Rc0 + Q.c0;
            const float hh0 = fmaf(Q.nc, Rc, -Q.nk * Q.b0), dhh = fmaf(Q.nc, Q.dc, -Q.nk * Q.db);
            const float am = fmaf(0.5f * (we - ws), Q.da, Q.a0);
            const float u0 = 0.5f * am * am * f_rcp(hh0);       // rho - hh at mid-window (second order in the cell size)
            const float g0[4] = {Q.a0, fmaf(hh0, K.sD, -Q.a0 * K.cD), fmaf(Q.nk, u0, -Q.b0),
                                 fmaf(fmaf(Q.nc, u0, Rc), K.tD, fmaf(-Q.nk, u0, Q.b0))};
            const float g1[4] = {Q.da, fmaf(dhh, K.sD, -Q.da * K.cD), -Q.db, fmaf(Q.dc, K.tD, Q.db)};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if ((i == 2 && cap_n) || (i == 3 && cap_s)) continue;            // no wall towards the pole
                const float tc = -g0[i] * f_rcp(g1[i]);
                if (g1[i] > 0.0f) { if (tc > t_in) { t_in = tc; wall_in = i; } }
                else if (g1[i] < 0.0f) t_out = fminf(t_out, tc);
                else if (g0[i] < 0.0f) return back ? FT_DEFER_R(14) : FT_MISS;    // parallel to the wall and outside it
            }
        }
        const float dl = 1.0e-4f * fminf(t_out - t_in, 16.0f * K.cell) + 1.0e-8f * K.R;
        // first visit: from where the walk stands (it may have entered the cell's shell in mid-cell)
        const float t_walk = from_entry ? -INFINITY : -K.pad;
        const float ta = fmaxf(fmaxf(t_walk, t_in - dl), t_lo), tb = fminf(fminf(t_out + dl, t_hi), t_cap);
        if (!(tb > ta)) return back ? FT_DEFER_R(14) : FT_MISS;  // the ray does not cross this cell (walk tolerance)
        if (cap_n || cap_s) {
            // distance from the polar axis is linear in t: its minimum over the interval is at an end
            const float Rc = Q.Rc0 + Q.c0;
            const float hh0 = fmaf(Q.nc, Rc, -Q.nk * Q.b0), dhh = fmaf(Q.nc, Q.dc, -Q.nk * Q.db);
            if (!(fminf(fmaf(ta, dhh, hh0), fmaf(tb, dhh, hh0)) > 1.0e-3f * K.cell)) return FT_DEFER_R(1);
        }
        float fc, fr;
        // start of the window: the ray must be clear of the surface there
        const float fa = local_f(Q, K, ta, fc, fr);
        MRTX_DBG("  fast_test back=%d cell (%d,%d) t_in=%.4e t_out=%.4e ta=%.4e tb=%.4e dl=%.3e wall_in=%d fa=%.4e (fc=%.5f fr=%.5f) from_entry=%d\n",
                 back, P.r0, P.c0, t_in, t_out, ta, tb, dl, wall_in, fa, fc, fr, (int)from_entry);
        if (!(fa > K.mg)) {
            // Below the surface where the ray comes in through a wall: the crossing lies in the cell behind that
            // wall, which the float32 walk did not propose (the ray clips it within the walk's wall tolerance).
            if (!(fa < -K.mg) || back >= 4) return FT_DEFER_R(fa < -K.mg ? 2 : 3);
            if (!(ta == t_in - dl)) {
                // ... or in this very cell, before the point the walk arrived at: search it from its entry
                if (from_entry || !(t_in - dl < ta)) return FT_DEFER_R(2);
                from_entry = true;
                continue;
            }
            from_entry = true;
            int r0 = P.r0, c0 = P.c0;
            if (wall_in == 0) c0 = c0 == 0 ? hf.W - 1 : c0 - 1;
            else if (wall_in == 1) c0 = c0 + 1 == hf.W ? 0 : c0 + 1;
            else if (wall_in == 2) r0 -= 1;
            else r0 += 1;
            load_raw_patch<I16>(hf, r0, c0, P);
            t_cap = t_in + dl;
            continue;
        }
        const float fca = fc, fra = fr;
        const float fb = local_f(Q, K, tb, fc, fr);
        const float dfc = fc - fca, dfr = fr - fra;             // direction of travel through the cell
        const float tm = 0.5f * (ta + tb), h = 0.5f * (tb - ta);
        const float fm = local_f(Q, K, tm, fc, fr);
        if (any_hit && fm < -K.mg) {
            // a shadow ray only asks whether the surface is crossed: above it at the start, below it in mid-cell
            out.s = s_c + (double)tm; out.fc = fc; out.fr = fr; out.r0 = P.r0; out.c0 = P.c0;
            out.d00 = d00; out.d01 = d01; out.d10 = d10; out.d11 = d11;
            return FT_HIT;
        }
        // f ~ fm + c1 tau + c2 tau^2 (a bilinear patch along a nearly straight track)
        const float ih = f_rcp(h);
        MRTX_DBG("    fb=%.4e fm=%.4e h=%.4e dfc=%.4f dfr=%.4f\n", fb, fm, h, dfc, dfr);
        // (curvature below the evaluation noise is no curvature: slivers of cells near the poles are 1e-7 R wide)
        const float curv = fa - 2.0f * fm + fb;
        const float c2 = fabsf(curv) > 4.0f * K.mg ? 0.5f * curv * ih * ih : 0.0f, c1 = 0.5f * (fb - fa) * ih;
        float lo, hi, flo, fhi;
        // (a sample within the margin of zero is a root there if f goes through it steeply: the conditioning
        //  check after the polish decides, a tangent ray ends in FT_DEFER)
        if (!(fm > K.mg)) { lo = -h; hi = 0.0f; flo = fa; fhi = fm; }
        else if (!(fb > K.mg)) { lo = 0.0f; hi = h; flo = fm; fhi = fb; }
        else {
            // no sign change at the three samples: a grazing double root shows up as a dip
            bool miss = !(c2 > 0.0f);
            float tv = 0.0f;
            if (!miss) {
                tv = -0.5f * c1 / c2;
                miss = !(tv > -h && tv < h) || fm - 0.25f * c1 * c1 / c2 > 0.25f * fminf(fm, fminf(fa, fb));    // dip clear of zero
            }
            float fv = 0.0f;
            if (!miss) {
                fv = local_f(Q, K, tm + tv, fc, fr);
                miss = fv > K.mg + 0.05f * c2 * h * h;
                if (!miss && !(fv < -K.mg)) return FT_DEFER_R(7);
            }
            if (miss) return back ? FT_DEFER_R(14) : FT_MISS;
            lo = tv > 0.0f ? 0.0f : -h; hi = tv; flo = tv > 0.0f ? fm : fa; fhi = fv;
        }
        // the parabola's root inside the bracket, polished on f itself
        float tau = 0.5f * (lo + hi);
        if (fabsf(fhi) <= K.mg) tau = hi;                       // the sample at the bracket's end IS the root (to the margin)
        else {
            const float disc = c1 * c1 - 4.0f * c2 * fm;
            if (disc >= 0.0f) {
                const float sq = f_sqrt_fast(disc);
                const float tq = -0.5f * (c1 + (c1 >= 0.0f ? sq : -sq));
                const float t1 = c2 != 0.0f ? tq / c2 : 2.0f * h, t2 = tq != 0.0f ? fm / tq : 2.0f * h;
                const bool in1 = t1 > lo && t1 < hi, in2 = t2 > lo && t2 < hi;
                if (in1) tau = t1;
                if (in2 && (!in1 || t2 < t1)) tau = t2;
            }
        }
        // f(lo) > 0 >= f(hi) brackets the crossing: Newton on the parabola's slope, regula falsi when that leaves the bracket
        float fx = 0.0f, slope = 0.0f;
        bool conv = false;
#pragma unroll 1
        for (int it = 0; it < 6; ++it) {
            fx = local_f(Q, K, tm + tau, fc, fr);
            slope = fmaf(2.0f * c2, tau, c1);
            if (!(slope < 0.0f)) slope = (fhi - flo) * f_rcp(hi - lo);
            conv = slope < 0.0f && fabsf(fx) <= -slope * K.ds_tol && fabsf(fx) <= 64.0f * K.mg;
            MRTX_DBG("    it=%d tau=%.6e fx=%.4e slope=%.4e lo=%.6e hi=%.6e flo=%.3e fhi=%.3e conv=%d need|fx|<=%.3e\n", it, tau, fx, slope, lo, hi, flo, fhi, (int)conv, -slope * K.ds_tol);
            // stop at 2 % of the tolerance; what only reaches the tolerance itself (grazing roots at the float32
            // noise floor) is accepted after the last iteration
            if (conv && it >= 1 && fabsf(fx) <= -slope * (0.02f * K.ds_tol)) break;
            if (fx > 0.0f) { lo = tau; flo = fx; } else { hi = tau; fhi = fx; }
            float tn = slope < 0.0f ? tau - fx / slope : INFINITY;
            if (!(tn >= lo && tn <= hi)) tn = lo - flo * (hi - lo) * f_rcp(fhi - flo);
            if (!(tn >= lo && tn <= hi)) tn = 0.5f * (lo + hi);
            tau = tn;
        }
        if (!conv) return FT_DEFER_R(slope < 0.0f ? 9 : 8);
        // whose root is it?  Outside the cell and moving further out: the ray left the cell before it
        // reached the surface (the next cell decides on its own patch).  Outside and moving in: the root
        // lies before the entry, where this patch is only an extrapolation -> exact path.
        const float e = K.epsc;
        bool outside = false, leaving = false;
        if (fc < -e) { outside = true; leaving = dfc < 0.0f; }
        else if (fc > 1.0f + e) { outside = true; leaving = dfc > 0.0f; }
        else if (fr < -e) { outside = true; leaving = dfr < 0.0f; }
        else if (fr > 1.0f + e) { outside = true; leaving = dfr > 0.0f; }
        if (outside) return leaving && !back ? FT_MISS : FT_DEFER_R(12);
        out.s = s_c + (double)(tm + tau);
        out.fc = fminf(fmaxf(fc, 0.0f), 1.0f); out.fr = fminf(fmaxf(fr, 0.0f), 1.0f);
        out.d00 = d00; out.d01 = d01; out.d10 = d10; out.d11 = d11;
        out.r0 = P.r0; out.c0 = P.c0;
        return FT_HIT;
    }
}

// ---- directional walk ----------------------------------------------------------------------------
// trav_step() solves all four walls of a cell at every node (two planes, two cones, each with its on-wall
// rules): ~600 instructions, most of them predicated off for most lanes.  A straight ray, however,
//   * turns about the polar axis in ONE sense for its whole length (x dy - y dx is constant), so of the
//     two longitude walls only the one ahead can ever be crossed, and
//   * has a single turning point in latitude (d(z/r)/ds has the sign of n0 + s n1, linear in s), so only
//     the latitude wall ahead matters, except in the one cell that contains the turning point.
// One plane and one cone per node, no wall can be crossed backwards, and the rule "within tol of the wall
// ahead = on it, leave now" is all the rounding protection that is needed.  Margins, pads and the
// conservative overlap test are those of trav_step().
struct Walk {
    float ox, oy, oz, dx, dy, dz, oo, od, smax;   // ray re-based at the bounding-sphere entry s_in
    float n0, n1;           // heading in latitude: northwards where n0 + s n1 > 0
    float s;                // current parameter (relative to s_in)
    int L, J, I;            // current cell
    int steps;
    bool east;              // longitude increases along the ray
    float vnext;            // the current cell's stored max if the parent's visit has fetched it already (NaN: not known)
    double s_in;
};

// ---- beam pre-pass ---------------------------------------------------------------------------------
// The samples of a pixel leave the same eye within a cone of half-angle Delta (the pixel's half diagonal): at the same
// ray parameter s a sample ray q lies within rho = s * Delta of the CENTRE ray c, and since q - c is (to second order)
// perpendicular to the ray, |q(s)| >= |c(s)| - rho * b / |c(s)| with b the centre ray's distance from the Moon's centre.
// The directions of q(s) and c(s) differ by at most rho / |c(s)|.  So wherever the centre ray, lowered by
// lift >= rho * b / r, stays above the max of a cell DILATED by that angle (dil[k]: one cell either side, valid as long
// as the level's cells are wider than the angle in both directions), no sample of the pixel can be below the surface.
// walk_step<.., BEAM> walks the centre ray through dil[] with that lift and stops at the first cell it cannot clear at the
// lowest level the dilation still covers; what it returns is a ray parameter before which every sample of the pixel is
// above the surface - the samples start their own walks there, a few levels from the bottom, instead of at the
// bounding sphere at level top - 3.
struct BeamCtl {
    float lift;         // scene units the centre ray is lowered by
    float rho_tex;      // angular radius of the beam in texels of the map (with a safety factor)
    int   lmin;         // lowest level the beam descends to: 2^lmin >= rho_tex
};

MRTX_HD inline float walk_r2(const Walk& w, float s) { return fmaf(s, fmaf(2.0f, w.od, s), w.oo); }

// float32 constants of the walk for the ray re-based at s_in (the same arithmetic whoever calls it: a walk that is
// suspended and resumed from (s_in, s, L, J, I) continues exactly as it would have)
MRTX_HD inline void walk_setup(const Ray64& R, double s_in, float smax, Walk& w) {
    w.s_in = s_in;
    w.ox = (float)fma(s_in, R.dx, R.ox); w.oy = (float)fma(s_in, R.dy, R.oy); w.oz = (float)fma(s_in, R.dz, R.oz);
    w.dx = (float)R.dx; w.dy = (float)R.dy; w.dz = (float)R.dz;
    w.oo = w.ox * w.ox + w.oy * w.oy + w.oz * w.oz;
    w.od = w.ox * w.dx + w.oy * w.dy + w.oz * w.dz;
    w.smax = smax;
    w.east = w.ox * w.dy - w.oy * w.dx > 0.0f;
    w.n0 = w.dz * w.oo - w.oz * w.od; w.n1 = w.dz * w.od - w.oz;
}

// where the ray leaves the bounding sphere R * dmax (s_end) - false if it misses it
MRTX_HD inline bool walk_extent(const HeightField& hf, double radius, const Ray64& R, double& s_first, double& s_end, double lift = 0.0) {
    const double Rb = radius * (double)hf.dmax + lift;
    const double disc = R.od * R.od - (R.oo - Rb * Rb);
    if (!(disc > 0.0)) return false;
    const double sq = disc * d_rsqrt(disc);
    s_first = -R.od - sq; s_end = -R.od + sq;
    return true;
}

// 0: the ray misses the bounding sphere; 1: it enters it, but not beyond s_min; 2: walk set up
MRTX_HD inline int walk_begin2(const HeightField& hf, double radius, const Ray64& R, double s_min, int start_level, Walk& w,
                               float t0_rel = 1e-5f, double lift = 0.0) {
    double s_first, s_end;
    if (!walk_extent(hf, radius, R, s_first, s_end, lift)) return 0;
    if (s_end <= s_min) return s_end > 0.0 ? 1 : 0;
    const double s_in = fmax(s_min, s_first);
    walk_setup(R, s_in, (float)(s_end - s_in), w);
    const int W = hf.W, H = hf.H;
    const int L = min(max(start_level, 0), hf.top);
    // first cell from the position just inside (a wrong neighbour is corrected by the on-wall rule)
    const float t0 = fminf(t0_rel * (float)radius, 0.5f * w.smax);
    const float x = fmaf(t0, w.dx, w.ox), y = fmaf(t0, w.dy, w.oy), z = fmaf(t0, w.dz, w.oz);
    const float lon = atan2f(x, -y), lat = atan2f(z, sqrtf(x * x + y * y));
    const float u = (lon * (0.5f / PI_F) + 0.5f) * (float)W - 0.5f, v = (0.5f - lat * (1.0f / PI_F)) * (float)H - 0.5f;
    int c0 = (int)floorf(u);
    c0 = c0 < 0 ? c0 + W : (c0 >= W ? c0 - W : c0);
    const int r0 = min(max((int)floorf(v), 0), H - 2);
    w.L = L; w.J = r0 >> L; w.I = c0 >> L;
    w.s = 0.0f; w.steps = 0; w.vnext = NAN;
    return 2;
}

MRTX_HD inline bool walk_begin(const HeightField& hf, double radius, const Ray64& R, double s_min, int start_level, Walk& w,
                               float t0_rel = 1e-5f, double lift = 0.0) {
    return walk_begin2(hf, radius, R, s_min, start_level, w, t0_rel, lift) == 2;
}

// Latitude walls.  A wall is the cone z = k r (k = sin phi).  Towards the poles that form loses the wall in
// float32: z and k r are both ~R while the distance to the wall only shows as r cos(phi) d(phi) in their
// difference - at 88.5 deg half a cell of the full-resolution map is 8e-7 R, the rounding error of z.  Walls
// beyond 45 deg are therefore tested as rho = kc r (kc = cos phi, rho = distance from the polar axis): the
// same cone, with small numbers where the other form has large ones.
// lat_side(): signed "northness" of a point relative to the wall, in units in which one radian of latitude is
// at least 0.7 r (so one tolerance serves both forms).
MRTX_HD inline float lat_side(float2 kk, float z, float rho, float r) {
    if (kk.x > 0.70710678f) return z > 0.0f ? kk.y * r - rho : -r;
    if (kk.x < -0.70710678f) return z < 0.0f ? rho - kk.y * r : r;
    return z - kk.x * r;
}

// smallest crossing of the wall in (lo, hi], or +inf
MRTX_HD inline float lat_cross(const Walk& w, float2 kk, float lo, float hi) {
    const float k = kk.x;
    float A, B, C;
    if (fabsf(k) > 0.70710678f) {
        const float c2 = kk.y * kk.y;
        A = fmaf(w.dx, w.dx, fmaf(w.dy, w.dy, -c2));
        B = fmaf(w.ox, w.dx, fmaf(w.oy, w.dy, -c2 * w.od));
        C = fmaf(w.ox, w.ox, fmaf(w.oy, w.oy, -c2 * w.oo));
    } else {
        const float k2 = k * k;
        A = fmaf(w.dz, w.dz, -k2); B = fmaf(w.oz, w.dz, -k2 * w.od); C = fmaf(w.oz, w.oz, -k2 * w.oo);
    }
    const float disc = fmaf(B, B, -A * C);
    float best = INFINITY;
    if (disc >= 0.0f) {
        const float q = -(B + copysignf(f_sqrt_fast(disc), B));
        const float r1 = q * f_rcp_fast(A), r2 = C * f_rcp_fast(q);     // A or q zero: inf / nan, rejected below
        if (r1 > lo && r1 <= hi && fmaf(r1, w.dz, w.oz) * k >= 0.0f) best = r1;
        if (r2 > lo && r2 <= hi && r2 < best && fmaf(r2, w.dz, w.oz) * k >= 0.0f) best = r2;
    }
    return best;
}

// cells per level, by arithmetic (= hf.nx[L], hf.ny[L]: indexing those kernel-parameter arrays with a per-lane level was
// 11 % of the kernel's instructions)
MRTX_HD inline int lvl_nx(const HeightField& hf, int L) { return (hf.W + (1 << L) - 1) >> L; }
MRTX_HD inline int lvl_ny(const HeightField& hf, int L) { return (hf.H - 2 + (1 << L)) >> L; }

// faces: 0 longitude wall ahead, 1 north wall, 2 south wall, 4 end of the ray
MRTX_HD inline bool walk_advance(const HeightField& hf, Walk& w, float sx, int face) {
    if (face == 4) return false;
    w.s = sx;
    // the index that moves, its step, and the boundary it crosses (the higher of the two cells' indices): the walk goes
    // up a level where that boundary is also a boundary of the level above.  Written without branches: the lanes that leave
    // through a longitude wall and those that leave through a latitude wall run these lines together (profiles/r08: the
    // branched form ran at 3.8 - 9.8 of 32 lanes and was 13 % of shadow_kernel_pool's instructions).
    const int L = w.L;
    const bool lon = face == 0;
    const int step = lon ? (w.east ? 1 : -1) : (face == 2 ? 1 : -1);
    const int from = lon ? w.I : w.J, to = from + step;
    const int nx = lvl_nx(hf, L);
    int I = lon ? to : w.I;
    const int J = lon ? w.J : to;
    I = I >= nx ? 0 : (I < 0 ? nx - 1 : I);
    if (J < 0 || J >= lvl_ny(hf, L)) return false;           // cannot happen (caps have no wall); be safe
    const int sh = (max(from, to) & 1) == 0 && L < hf.top ? 1 : 0;
    w.L = L + sh; w.J = J >> sh; w.I = I >> sh; w.vnext = NAN;
    return true;
}

// may the beam descend to row J of level L?  Its dilation must still cover rho_tex texels of longitude at the most
// poleward latitude of the rows J - 1 .. J + 1 (the rows next to a pole never qualify: cos -> 0)
MRTX_HD inline bool beam_level_ok(const HeightField& hf, int L, int J, float rho_tex) {
    const int r_n = max((J - 1) << L, 0), r_s = min((J + 2) << L, hf.H - 1);
    const float cn = MRTX_LDG(hf.latsc32 + r_n).y, cs = MRTX_LDG(hf.latsc32 + r_s).y;
    return (float)(1 << L) * fminf(cn, cs) >= rho_tex;
}

// BEAM: the pre-pass form (see BeamCtl): nodes are read from dil[], the shell is raised by bc->lift, and the cell
// at which the walk stops (TR_CANDIDATE, level >= bc->lmin >= MRTX_DIL_MIN_LEVEL) comes back with sx_out = the parameter at
// which the lowered ray enters that cell's shell.
// ASCEND (rays that leave the surface: shadow and bounce rays): the max of this cell's PARENT is fetched together with the
// cell's own.  When the ray leaves the cell sideways into a sibling (same parent) while rising above the parent's max, the
// walk continues at the parent: a rising ray stays above it, so the parent is skipped whole on the next step instead of
// sibling by sibling - the walk climbs a level per step on its way out instead of waiting for an aligned boundary.
// Measured at config 3 (MRTX_SQ_ASCEND): shadow-ray node visits -14 %, shadow_kernel 12.39 against 12.42 ms - the second
// fetch and its address arithmetic cost what the saved nodes gave; off.
template <bool I16, bool BEAM = false, bool ASCEND = false>
MRTX_HD inline int walk_step(const HeightField& hf, float Rf, float inv_rs, Walk& w, RawPatch& P,
                             float& sx_out, int& face_out, Counters& cnt, const BeamCtl* bc = nullptr,
                             const unsigned* loff = nullptr) {
    if (!loff) loff = hf.off;                                   // (hot kernels pass their shared-memory copy)
    if (++w.steps > MAX_STEPS) { ++cnt.overflow; return TR_END; }
    const int L = w.L, J = w.J, I = w.I;
    const int W = hf.W, H = hf.H;
    const float s = w.s;
    ++cnt.nodes;
#if defined(MRTX_LEVEL_HIST) && !defined(__CUDA_ARCH__)
    __atomic_fetch_add(&g_level_hist[L], 1ull, __ATOMIC_RELAXED);      // host tool: node visits by level (tools/trace_host.cu)
#endif
    // The walls ahead depend on (L, J, I) and the ray's headings only: their table entries are requested here,
    // together with the node itself, so that one memory latency covers all three (the loop is latency-bound).
    const bool north = fmaf(s, w.n1, w.n0) > 0.0f;
    const int jn = J << L, js = min((J + 1) << L, H - 1);
    const float2 wl = MRTX_LDG(hf.lon32 + (w.east ? min((I + 1) << L, W) : (I << L)));
    const float2 kk = MRTX_LDG(hf.latsc32 + (north ? jn : js));
    // The four children of this cell are requested NOW, with the cell itself: if the walk descends, the child's max is in
    // a register when the child is visited and its visit starts without a memory round trip (the kernel is bound by
    // the latency of the dependent chain node -> decision -> next node, not by bandwidth).
#if MRTX_PREFETCH
    unsigned ch0 = 0u, ch1 = 0u;                                // int16 maps: children (2J, 2I..2I+1) and (2J+1, 2I..2I+1), packed
    const bool pre = !BEAM && I16 && L >= 2;
    if (pre) {
        const int cnx = lvl_nx(hf, L - 1), cny = lvl_ny(hf, L - 1);
        const unsigned e0 = loff[L - 1] + (unsigned)(2 * J) * (unsigned)cnx + (unsigned)(2 * I);
        const unsigned e1 = 2 * J + 1 < cny ? e0 + (unsigned)cnx : e0;
        ch0 = mrtx_ldg_u32((const int16_t*)hf.lvl_base + e0);
        ch1 = mrtx_ldg_u32((const int16_t*)hf.lvl_base + e1);
    }
#endif
    float vpar = 0.0f;
    const bool par_ok = ASCEND && L < hf.top;
    if (par_ok) {
        const unsigned ep = loff[L + 1] + (unsigned)(J >> 1) * (unsigned)lvl_nx(hf, L + 1) + (unsigned)(I >> 1);
        vpar = I16 ? (float)MRTX_LDG((const int16_t*)hf.lvl_base + ep) : MRTX_LDG((const float*)hf.lvl_base + ep);
    }
    float vmax;
    if (BEAM) {
        // a row whose cells are narrower than the beam (towards the poles) cannot bound it: the pre-pass ends here and
        // the samples walk on from this point on their own
        if (!beam_level_ok(hf, L, J, bc->rho_tex)) { sx_out = s; return TR_CANDIDATE; }
        const unsigned e = loff[MRTX_MAX_LEVELS + L] + (unsigned)J * (unsigned)lvl_nx(hf, L) + (unsigned)I;
        vmax = I16 ? (float)MRTX_LDG((const int16_t*)hf.lvl_base + e) : MRTX_LDG((const float*)hf.lvl_base + e);
    } else if (L == 0) {
        const int c1 = I + 1 == W ? 0 : I + 1;
        P.r0 = J; P.c0 = I;
        if (I16) {
            const int16_t* b = (const int16_t*)hf.base + (size_t)J * W;
            P.v00 = (float)MRTX_LDG(b + I); P.v01 = (float)MRTX_LDG(b + c1);
            P.v10 = (float)MRTX_LDG(b + W + I); P.v11 = (float)MRTX_LDG(b + W + c1);
        } else {
            const float* b = (const float*)hf.base + (size_t)J * W;
            P.v00 = MRTX_LDG(b + I); P.v01 = MRTX_LDG(b + c1);
            P.v10 = MRTX_LDG(b + W + I); P.v11 = MRTX_LDG(b + W + c1);
        }
        vmax = fmaxf(fmaxf(P.v00, P.v01), fmaxf(P.v10, P.v11));
    } else {
        if (w.vnext == w.vnext) vmax = w.vnext;
        else {
#if MRTX_TILED
            // (the level in 8 x 8-cell tiles, pyramid.cu)
            const unsigned tx = (unsigned)(lvl_nx(hf, L) + 7) >> 3;
            const unsigned e = loff[2 * MRTX_MAX_LEVELS + L] + ((((unsigned)J >> 3) * tx + ((unsigned)I >> 3)) << 6) + (((unsigned)J & 7u) << 3) + ((unsigned)I & 7u);
#else
            const unsigned e = loff[L] + (unsigned)J * (unsigned)lvl_nx(hf, L) + (unsigned)I;
#endif
            vmax = I16 ? (float)MRTX_LDG((const int16_t*)hf.lvl_base + e) : MRTX_LDG((const float*)hf.lvl_base + e);
        }
    }
    const float dmax = decode_bound<I16>(hf, vmax, inv_rs);
    const float marg = 3.0e-6f * Rf + (BEAM ? bc->lift : 0.0f);      // float32 error of a radius near R + the cheap decode
    const float rc = fmaf(Rf, dmax, marg), rc2 = rc * rc;
    const float r2s = walk_r2(w, s);
    float sd = s;
    if (!(r2s <= rc2 && L > 0)) {       // (inside the cell's shell already: descend where we stand)
        const float x = fmaf(s, w.dx, w.ox), y = fmaf(s, w.dy, w.oy), z = fmaf(s, w.dz, w.oz);
        const float rs = f_sqrt_fast(fmaxf(r2s, 1e-30f));
        // "On the wall" = within the float32 error of the wall function, which is NOT uniform: the re-based ray
        // carries 1e-7 r everywhere, the wall functions add the rounding of their two cancelling terms.  A uniform
        // 2e-6 r (3 % of a full-resolution cell) would hand a ray that runs nearly parallel to a wall to the next
        // row tens of cells early, and a directional walk never comes back.
        const float tol0 = 2.0e-7f * rs;
        float sx = w.smax;
        int face = 4;
        {   // the longitude wall ahead
            const float sg = w.east ? 1.0f : -1.0f;
            const float g = sg * fmaf(x, wl.x, y * wl.y), dg = sg * fmaf(w.dx, wl.x, w.dy * wl.y);      // outwards positive
            const float tol = fmaf(6.0e-7f, fabsf(x * wl.x) + fabsf(y * wl.y), tol0);
            if (dg > 0.0f) {
                if (g >= -tol) { sx = s; face = 0; }            // on (or just beyond) it, moving out: leave now
                else {
                    const float sc = s - g * f_rcp_fast(dg);
                    const float px = fmaf(sc, w.dx, w.ox), py = fmaf(sc, w.dy, w.oy);
                    if (px * wl.y - py * wl.x > 0.0f && sc < sx) { sx = sc; face = 0; }     // not the opposite half-plane
                }
            }
        }
        if (sx > s) {   // the latitude wall ahead
            float s_turn = INFINITY;                            // where the heading reverses, if that is still ahead
            if (w.n1 != 0.0f) { const float t = -w.n0 * f_rcp_fast(w.n1); if (t > s) s_turn = t; }
            float sl = INFINITY;
            int fl = north ? 1 : 2;
            if (north ? jn > 0 : js < H - 1) {                  // polar caps have no wall
                const float side = lat_side(kk, z, f_sqrt_fast(fmaf(x, x, y * y)), rs);
                const float G = north ? side : -side;                   // outwards positive
                const float tol = fmaf(6.0e-7f * rs, fminf(fabsf(kk.x), kk.y), tol0);
                sl = G >= -tol ? s : lat_cross(w, kk, s, s_turn);
            }
            if (!(sl <= s_turn) && s_turn < sx) {
                // the ray turns inside this cell: from there on it heads for the other wall
                sl = INFINITY; fl = north ? 2 : 1;
                if (north ? js < H - 1 : jn > 0) sl = lat_cross(w, MRTX_LDG(hf.latsc32 + (north ? js : jn)), s_turn, INFINITY);
            }
            if (sl < sx) { sx = sl; face = fl; }
        }
        const float pad = 4.0e-6f * Rf;
        const float ta = fmaxf(s - pad, 0.0f), tb = fminf(sx + pad, w.smax);
        const float tm = fminf(fmaxf(-w.od, ta), tb);
        MRTX_DBG("walk %d L%d J%d I%d s=%.7f sx=%.7f face=%d rc=%.7f r(s)=%.7f r(tm)=%.7f east=%d north=%d\n", w.steps, L, J, I, s, sx, face, rc,
                 sqrtf(r2s), sqrtf(walk_r2(w, tm)), (int)w.east, (int)(fmaf(s, w.n1, w.n0) > 0.0f));
        if (!(walk_r2(w, tm) <= rc2)) {
            if (!walk_advance(hf, w, sx, face)) return TR_END;
            if (par_ok && w.L == L && (w.J >> 1) == (J >> 1) && (w.I >> 1) == (I >> 1)) {
                const float sa = fmaxf(sx - pad, 0.0f);
                const float rcp = fmaf(Rf, decode_bound<I16>(hf, vpar, inv_rs), marg);
                if (w.od + sa > 0.0f && walk_r2(w, sa) > rcp * rcp) { w.L = L + 1; w.J >>= 1; w.I >>= 1; w.vnext = vpar; }
            }
            return TR_CONTINUE;
        }
        if (!BEAM && L == 0) { sx_out = sx; face_out = face; return TR_CANDIDATE; }
        if (r2s > rc2) {
            const float dq = fmaf(w.od, w.od, rc2 - w.oo);
            if (dq > 0.0f) sd = fminf(fmaxf(-w.od - f_sqrt_fast(dq), s), sx);
        }
    }
    if (BEAM && L <= bc->lmin) { sx_out = sd; return TR_CANDIDATE; }
    // pick the child at sd
    const float x = fmaf(sd, w.dx, w.ox), y = fmaf(sd, w.dy, w.oy), z = fmaf(sd, w.dz, w.oz);
    const int mi = (2 * I + 1) << (L - 1), mj = (2 * J + 1) << (L - 1);
    int ci = 2 * I, cj = 2 * J;
    if (mi < min((I + 1) << L, W)) {
        const float2 wl = MRTX_LDG(hf.lon32 + mi);
        if (fmaf(x, wl.x, y * wl.y) >= 0.0f) ci += 1;
    }
    const float rho2 = fmaf(x, x, y * y);
    const float rho = f_sqrt_fast(rho2), rr = f_sqrt_fast(fmaf(z, z, rho2));
    if (mj < min((J + 1) << L, H - 1)) {
        if (lat_side(MRTX_LDG(hf.latsc32 + mj), z, rho, rr) < 0.0f) cj += 1;    // south of the mid wall
    }
    if (BEAM && !beam_level_ok(hf, L - 1, cj, bc->rho_tex)) { sx_out = sd; return TR_CANDIDATE; }
#if MRTX_DESCENT2
    // Two levels at a time: measured on config 3, a camera ray visits 1.15 nodes per level on its way down - the levels in
    // between cull next to nothing, but every visit is a dependent fetch and two wall solves.  The grandchild that holds
    // p(sd) is found with two more side tests and the walk goes on there.
    if (!BEAM && L >= MRTX_DESCENT2) {
        const int L1 = L - 1;
        const int mi2 = (2 * ci + 1) << (L1 - 1), mj2 = (2 * cj + 1) << (L1 - 1);
        int gi = 2 * ci, gj = 2 * cj;
        if (mi2 < min((ci + 1) << L1, W)) {
            const float2 wl2 = MRTX_LDG(hf.lon32 + mi2);
            if (fmaf(x, wl2.x, y * wl2.y) >= 0.0f) gi += 1;
        }
        if (mj2 < min((cj + 1) << L1, H - 1)) {
            if (lat_side(MRTX_LDG(hf.latsc32 + mj2), z, rho, rr) < 0.0f) gj += 1;
        }
        w.s = sd; w.L = L - 2; w.I = gi; w.J = gj; w.vnext = NAN;
        return TR_CONTINUE;
    }
#endif
    w.s = sd; w.L = L - 1; w.I = ci; w.J = cj;
#if MRTX_PREFETCH
    if (pre) {
        const unsigned pair = (cj & 1) ? ch1 : ch0;
        w.vnext = (float)(int16_t)((ci & 1) ? pair >> 16 : pair & 0xffffu);
    } else w.vnext = NAN;
#else
    w.vnext = NAN;
#endif
    return TR_CONTINUE;
}

// ---- ceiling test (shadow rays) ---------------------------------------------------------------------
// A ray that is RISING (p . d >= 0: its distance from the centre only grows from here on) and stands at the point
// p(s) of cell (L, J, I) has nothing left to meet once it is above everything it can still pass over before it leaves
// the bounding sphere.  dil[K] of the level-K cell that contains p bounds the surface for every direction within one
// level-K cell of it, and the ray cannot leave that neighbourhood before it has travelled D_K = one cell's width at the
// neighbourhood's most poleward latitude; by then its radius has grown to r(s + D_K), which is compared with dil[K + 2],
// and so on up to the top level - a handful of independent loads instead of the ~8 dependent node visits the walk
// needs to climb from level 3 to the top of the pyramid.  "Clear" is exact (conservative bounds only); "not clear"
// says nothing, the walk just goes on.
template <bool I16>
MRTX_HD inline bool ceiling_clear(const HeightField& hf, const FastConsts& K, float inv_rs, const Walk& w, float dmin,
                                  const unsigned* loff = nullptr) {
    if (!loff) loff = hf.off;
    const float s = w.s;
    const float pd = w.od + s;                                  // p(s) . d  (|d| = 1)
    if (!(pd >= 0.0f)) return false;
    int L = w.L, J = w.J, I = w.I;
    if (L < MRTX_DIL_MIN_LEVEL) return false;
    const float r2 = walk_r2(w, s);
    const float x = fmaf(s, w.dx, w.ox), y = fmaf(s, w.dy, w.oy);
    const float cosphi = f_sqrt_fast(fmaf(x, x, y * y)) * f_rsqrt(r2);
    const float ang_lat = 1.0f / K.Kh, ang_lon = 1.0f / K.Kw;  // one texel of latitude / of longitude at the equator, radians
    const float Rmin = K.R * dmin;
    const float Rb = K.R * hf.dmax, Rb2 = Rb * Rb;
    const float marg = 3.0e-6f * K.R;
    float t = 0.0f;                                             // the ray has surely travelled this far when stage K is looked at
    for (;;) {
        const float rt2 = fmaf(t, fmaf(2.0f, pd, t), r2);       // r^2 after t, a lower bound of r^2 from there on
        if (rt2 >= Rb2) return true;                            // beyond the bounding sphere
        const unsigned e = loff[MRTX_MAX_LEVELS + L] + (unsigned)J * (unsigned)lvl_nx(hf, L) + (unsigned)I;
        const float vmax = I16 ? (float)MRTX_LDG((const int16_t*)hf.lvl_base + e) : MRTX_LDG((const float*)hf.lvl_base + e);
        const float rc = fmaf(K.R, decode_bound<I16>(hf, vmax, inv_rs), marg);
        if (!(rt2 > rc * rc)) return false;
        // distance before which the ray cannot have left the 3 x 3 neighbourhood of its level-L cell
        const float cells = (float)(1 << L);
        const float cosmin = cosphi - 2.0f * cells * ang_lat;   // cos is 1-Lipschitz: the neighbourhood's most poleward latitude
        if (!(cosmin > 0.0f)) return false;
        const float D = 0.9f * cells * fminf(ang_lat, ang_lon * cosmin) * Rmin;
        if (L >= hf.top) {
            // the top level's neighbourhood must see the ray out of the bounding sphere
            const float re2 = fmaf(D, fmaf(2.0f, pd, D), r2);
            return re2 >= Rb2;
        }
        t = D;
        const int up = min(2, hf.top - L);
        L += up; J >>= up; I >>= up;
    }
}

// Beam pre-pass of one pixel: centre ray C, angular radius delta (radians) of the pixel.  Returns false if no sample
// of the pixel can hit at all; else s_start = a parameter before which every sample is above the surface and level =
// the level the walk stopped at (the samples start a little below it).
template <bool I16>
MRTX_HD inline bool beam_walk(const HeightField& hf, const FastConsts& K, double radius, const Ray64& C, double delta,
                              double& s_start, int& level, Counters& cnt) {
    const double Rmin = radius * (double)hf.dmin;
    // rho: the beam's radius at the far end of anything it can meet (its centre ray is nearest to the Moon's centre at
    // s = -od; nothing it meets lies further than a bounding-sphere radius beyond that)
    const double Rb = radius * (double)hf.dmax;
    const double b2 = fmax(C.oo - C.od * C.od, 0.0), b = sqrt(b2);
    const double rho = (fmax(-C.od, 0.0) + 1.1 * Rb) * delta;
    if (b - 1.05 * rho > Rb) return false;                      // no ray of the pixel comes within the bounding sphere
    BeamCtl bc;
    bc.lift = (float)(1.02 * rho * b / Rmin + rho * rho / Rmin + 1.0e-6 * radius);
    bc.rho_tex = (float)(1.1 * (rho / Rmin) * (double)K.Kh + 0.25);
    int lmin = MRTX_DIL_MIN_LEVEL;
    while ((float)(1 << lmin) < bc.rho_tex && lmin < hf.top) ++lmin;
    bc.lmin = lmin;
    s_start = 0.0; level = hf.top;
    if ((float)(1 << lmin) < bc.rho_tex || (double)bc.lift > 0.01 * radius) return true;     // beam too wide to bound: no information
    Walk w;
    if (!walk_begin(hf, radius, C, 0.0, hf.top, w, 1e-5f, (double)bc.lift)) return false;
    const float Rf = (float)radius, inv_rs = 1.0f / hf.radius_scale;
    for (;;) {
        RawPatch P;
        float sx;
        int face;
        const int r = walk_step<I16, true>(hf, Rf, inv_rs, w, P, sx, face, cnt, &bc);
        if (r == TR_END) return false;
        if (r == TR_CANDIDATE) {
            s_start = fmax(w.s_in + (double)sx - 2.0 * (double)K.pad, 0.0);
            level = w.L;
            return true;
        }
    }
}

// Sequential form (host tool): FT_HIT / FT_MISS / FT_DEFER for the whole ray.
template <bool I16>
MRTX_HD inline int trace_ray_fast(const HeightField& hf, const FastConsts& K, double radius, const Ray64& R, double s_min,
                                  int start_level, FastHit& out, Counters& cnt, int ceil_level = 0) {
    Walk w;
    if (!walk_begin(hf, radius, R, s_min, start_level, w)) return FT_MISS;
    const float Rf = (float)radius, inv_rs = 1.0f / hf.radius_scale;
    int ceil_next = ceil_level > 0 ? ceil_level : 0x7fffffff;
    for (;;) {
        RawPatch P;
        float sx;
        int face;
        if (w.L >= ceil_next) {
            ceil_next = w.L + 2;
            if (ceiling_clear<I16>(hf, K, inv_rs, w, hf.dmin)) return FT_MISS;
        }
        const int r = walk_step<I16>(hf, Rf, inv_rs, w, P, sx, face, cnt);
        if (r == TR_END) return FT_MISS;
        if (r == TR_CANDIDATE) {
            ++cnt.tests;
            const int t = fast_test<I16>(hf, K, R, w.s_in, s_min, w.s, sx, w.smax, P, false, out);
            if (t != FT_MISS) return t;                      // FT_DEFER carries its reason in the upper bits
            if (!walk_advance(hf, w, sx, face)) return FT_MISS;
        }
    }
}

}  // namespace mrtx_core
