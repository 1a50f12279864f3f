/*
 * TEST INFRASTRUCTURE - float64 CPU oracle of the render hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this; the product (moonrtx_b200) never does.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in the closed PlotOptiX engine
 * (plotoptix>=0.19.2, requirements.txt:3), which is not under /root/reference and cannot
 * run here.  The reference has no tests, golden images or known-answer vectors for it.
 * This file therefore restates the scene CONTRACT the reference's own code fixes:
 *   - geometry: sphere radius 10 displaced radially by the float32 map D, surface radius
 *     = 10 * D(u, v)                                   moon_renderer.py:37, 620-624
 *   - texel convention: bilinear between texel centres (-0.5 offset), columns wrap, rows
 *     clamp; row 0 = +90 deg, u = (lon+180)/360, v = (90-lat)/180
 *                                                      renderer_navigation.py:558-599
 *   - body frame: +Z north pole, -Y lon 0, +X lon +90 E; lat = asin(z), lon = atan2(x,-y)
 *                                                      renderer_navigation.py:47-53, 452-492
 *   - orientation: u = R[:,2], v = -R[:,1]             moon_renderer.py:844-845
 *   - pinhole camera eye/target/up/vertical fov        moon_renderer.py:507-520, 627-635
 *   - spherical sun light, radiance * solid angle      moon_renderer.py:65-83, 640, 859
 *   - shadow-ray origin lifted by scene_epsilon        moon_renderer.py:85-93
 *   - Gamma post-process (exposure * L)^(1/gamma)      moon_renderer.py:597-600
 * and is pinned only on what the reference itself can state: get_elevation_m and
 * hit_to_selenographic fixtures (tests/golden/convention.npz), the camera / light vectors
 * (tests/golden/scene_vectors.json) and analytic known answers (flat sphere, shadow
 * length h/tan(alt), terminator at N.L = 0).
 *
 * Algorithm (deliberately NOT the product's): no pyramid.  The ray is walked through
 * every (lon, lat) texel cell it crosses, cell exits are the exact plane / cone
 * crossings, and inside a cell the first root of f(s) = |p(s)| - R*bilinear(u(s), v(s))
 * is bracketed on a fine subdivision (with a local-minimum probe for grazing double
 * roots) and polished by bisection-safeguarded secant steps, all in float64.
 * Around the Moon: rays that miss it see the Sun disk / the star map (N1); overlay tubes are capsules tested
 * one by one against every camera sample (N4); with n_bounce > 0 a path continues from every hit in a cosine-distributed
 * direction, by Russian roulette on the albedo, and collects the direct light of the hits it finds (N2) - the same random
 * dimensions as the product, so that paths can be compared sample for sample.
 */

#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PI 3.14159265358979323846

typedef struct {
    /* height field: exactly one of map_f32 / map_i16 is non-null */
    int W, H;
    const float* map_f32;
    const int16_t* map_i16;
    float scale, radius_scale;
    double dmax;                /* global max of D (bounding sphere = radius * dmax) */
    /* body placement: scene -> body rows, centre, radius */
    double ex[3], ey[3], ez[3], pos[3], radius;
    /* camera */
    double eye[3], w[3], right[3], up[3], tan_half_fov;
    int img_w, img_h;
    /* light */
    double light_pos[3], light_radius, light_radiance, scene_epsilon;
    int jitter, shadows;
    /* albedo texture (RGBA8, may be null) */
    const uint8_t* tex;
    int tex_w, tex_h;
    float exposure, inv_gamma;
    /* what a ray that misses the Moon sees (moon_renderer.py:602-609, 643-650): the visible Sun disk, a flat-shaded
     * sphere in scene space (radius <= 0: none), else the star map, an equirectangular RGBA8 environment texture of
     * linear radiance (null: black) */
    double sun_disk_pos[3], sun_disk_radius, sun_disk_color[3];
    const uint8_t* env;
    int env_w, env_h;
    /* overlay tubes (rt.set_graph: grid lines, labels, pins - renderer_labels.py:263-305, renderer_pins.py:18-55): n_tubes
     * capsules, 12 floats each = (a.xyz, r, b.xyz, -, colour.rgb, -), scene space.  Flat-shaded (the colour is the radiance
     * of a camera sample that meets the tube before the surface), never an occluder of the sun (renderer_labels.py:133-139) */
    const float* tubes;
    int n_tubes;
    /* diffuse interreflection bounces after the camera hit (rt.set_uint("path_seg_range", 2, 4), moon_renderer.py:583:
     * max segments - 2; 0 = direct light only) */
    int n_bounce;
} orc_scene;

/* ------------------------------------------------------------------------------ */
static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

static double texel(const orc_scene* S, int r, int c) {
    if (S->map_f32) return (double)S->map_f32[(size_t)r * S->W + c];
    /* D exactly as data_loader.py:219-242 would hold it in float32 */
    volatile float v = (float)S->map_i16[(size_t)r * S->W + c];
    v = v * S->scale;
    v = v + 1.0f;
    v = v / S->radius_scale;
    return (double)v;
}

/* continuous texel coordinates of a body-frame point (renderer_navigation.py:577-582) */
static void point_uv(const orc_scene* S, const double* p, double* u, double* v, double* lon, double* lat) {
    *lon = atan2(p[0], -p[1]);
    *lat = atan2(p[2], sqrt(p[0] * p[0] + p[1] * p[1]));
    *u = (*lon / (2.0 * PI) + 0.5) * S->W - 0.5;
    *v = (0.5 - *lat / PI) * S->H - 0.5;
}

static void cell_of(const orc_scene* S, double u, double v, int* r0, int* c0) {
    int c = (int)floor(u);
    if (c < 0) c += S->W;
    if (c >= S->W) c -= S->W;
    int r = (int)floor(v);
    if (r < 0) r = 0;
    if (r > S->H - 2) r = S->H - 2;
    *r0 = r; *c0 = c;
}

typedef struct { int r0, c0; double d00, d01, d10, d11; } patch_t;

static void load_patch(const orc_scene* S, int r0, int c0, patch_t* P) {
    const int c1 = (c0 + 1) % S->W;                       /* wraps at the +/-180 seam */
    P->r0 = r0; P->c0 = c0;
    P->d00 = texel(S, r0, c0);     P->d01 = texel(S, r0, c1);
    P->d10 = texel(S, r0 + 1, c0); P->d11 = texel(S, r0 + 1, c1);
}

/* fractional position inside a patch; fc may leave [0,1] slightly (smooth extension),
 * fr is clamped exactly like renderer_navigation.py:584-585 */
static void patch_frac(const orc_scene* S, const patch_t* P, double u, double v, double* fc, double* fr) {
    double a = u - P->c0;
    if (a < -0.5 * S->W) a += S->W;
    if (a > 0.5 * S->W) a -= S->W;
    double b = v - P->r0;
    if (b < 0.0) b = 0.0;
    if (b > 1.0) b = 1.0;
    *fc = a; *fr = b;
}

static double patch_value(const patch_t* P, double fc, double fr) {
    return P->d00 * (1.0 - fr) * (1.0 - fc) + P->d10 * fr * (1.0 - fc) + P->d01 * (1.0 - fr) * fc + P->d11 * fr * fc;
}

/* displacement factor at (lon, lat) in degrees - the reference's get_elevation_m without
 * the unit conversion; exported so the tests can pin it on the reference fixture */
double orc_displacement(const orc_scene* S, double lat_deg, double lon_deg) {
    const int h = S->H, w = S->W;
    double row = (90.0 - lat_deg) / 180.0 * h - 0.5;
    double col = fmod((lon_deg + 180.0) / 360.0 * w - 0.5, (double)w);
    if (col < 0) col += w;
    int r0 = (int)floor(row);
    if (r0 < 0) r0 = 0;
    if (r0 > h - 2) r0 = h - 2;
    double fr = row - r0;
    if (fr < 0) fr = 0;
    if (fr > 1) fr = 1;
    int c0 = (int)floor(col);
    if (c0 >= w) c0 = w - 1;
    patch_t P;
    load_patch(S, r0, c0, &P);
    return patch_value(&P, col - c0, fr);
}

typedef struct { const orc_scene* S; const double* o; const double* d; const patch_t* P; } fctx_t;

static double f_eval(const fctx_t* F, double s) {
    double p[3] = {F->o[0] + s * F->d[0], F->o[1] + s * F->d[1], F->o[2] + s * F->d[2]};
    double u, v, lon, lat, fc, fr;
    point_uv(F->S, p, &u, &v, &lon, &lat);
    patch_frac(F->S, F->P, u, v, &fc, &fr);
    return sqrt(dot3(p, p)) - F->S->radius * patch_value(F->P, fc, fr);
}

/* polish a bracket lo (f > 0) .. hi (f <= 0) */
static double refine(const fctx_t* F, double lo, double flo, double hi, double fhi) {
    for (int it = 0; it < 200 && hi - lo > 1e-14 * (1.0 + fabs(hi)); ++it) {
        double m = (it & 1) ? 0.5 * (lo + hi) : lo + (hi - lo) * flo / (flo - fhi);
        if (!(m > lo && m < hi)) m = 0.5 * (lo + hi);
        const double fm = f_eval(F, m);
        if (fm > 0.0) { lo = m; flo = fm; } else { hi = m; fhi = fm; }
    }
    return hi;
}

/* first root of f in [a, b] (a patch crossing); returns 1 and *s_hit, or 0 */
static int patch_first_root(const fctx_t* F, double a, double b, double* s_hit) {
    enum { N = 16 };
    double s[N + 1], f[N + 1];
    for (int i = 0; i <= N; ++i) {
        s[i] = a + (b - a) * i / N;
        f[i] = f_eval(F, s[i]);
    }
    if (f[0] <= 0.0) { *s_hit = a; return 1; }          /* entered below the surface */
    for (int i = 1; i <= N; ++i) {
        if (f[i] <= 0.0) { *s_hit = refine(F, s[i - 1], f[i - 1], s[i], f[i]); return 1; }
        /* grazing: a local minimum between samples may dip below zero */
        if (i < N && f[i] < f[i - 1] && f[i] <= f[i + 1]) {
            double lo = s[i - 1], hi = s[i + 1];
            const double g = 0.6180339887498949;
            double x1 = hi - g * (hi - lo), x2 = lo + g * (hi - lo);
            double f1 = f_eval(F, x1), f2 = f_eval(F, x2);
            for (int it = 0; it < 80; ++it) {
                if (f1 <= 0.0) { *s_hit = refine(F, s[i - 1], f[i - 1], x1, f1); return 1; }
                if (f2 <= 0.0 ) {
                    /* is there an earlier crossing before x2?  x1 > 0 here, so bracket (x1, x2) */
                    *s_hit = refine(F, x1, f1, x2, f2); return 1;
                }
                if (f1 < f2) { hi = x2; x2 = x1; f2 = f1; x1 = hi - g * (hi - lo); f1 = f_eval(F, x1); }
                else { lo = x1; x1 = x2; f1 = f2; x2 = lo + g * (hi - lo); f2 = f_eval(F, x2); }
                if (hi - lo < 1e-13) break;
            }
        }
    }
    return 0;
}

/* parameter at which the ray leaves the (lon, lat) cell (r0, c0), given it is inside at s */
/* *face: 0 west, 1 east, 2 north, 3 south, 4 = end of the ray */
static double cell_exit(const orc_scene* S, const double* o, const double* d, int r0, int c0, double s, double s_end, int* face) {
    double best = s_end;
    *face = 4;
    const double oo = dot3(o, o), od = dot3(o, d);
    /* longitude half-planes through the polar axis */
    for (int side = 0; side < 2; ++side) {
        const double lam = ((c0 + side + 0.5) / S->W - 0.5) * 2.0 * PI;
        const double t[3] = {cos(lam), sin(lam), 0.0}, e[3] = {sin(lam), -cos(lam), 0.0};
        const double g0 = dot3(o, t), g1 = dot3(d, t);
        if ((side == 1 && g1 > 0.0) || (side == 0 && g1 < 0.0)) {
            const double sc = -g0 / g1;
            const double p[3] = {o[0] + sc * d[0], o[1] + sc * d[1], o[2] + sc * d[2]};
            if (sc > s && sc < best && dot3(p, e) > 0.0) { best = sc; *face = side; }
        }
    }
    /* latitude cones about the polar axis (none at the polar caps) */
    for (int side = 0; side < 2; ++side) {
        if (side == 0 && r0 == 0) continue;
        if (side == 1 && r0 == S->H - 2) continue;
        const double phi = (0.5 - (r0 + side + 0.5) / S->H) * PI;
        const double k = sin(phi), k2 = k * k;
        const double A = d[2] * d[2] - k2, B = o[2] * d[2] - k2 * od, Cq = o[2] * o[2] - k2 * oo;
        double roots[2];
        int nr = 0;
        if (fabs(A) < 1e-300) {
            if (B != 0.0) roots[nr++] = -Cq / (2.0 * B);
        } else {
            const double disc = B * B - A * Cq;
            if (disc >= 0.0) {
                const double q = -(B + (B >= 0 ? 1.0 : -1.0) * sqrt(disc));
                roots[nr++] = q / A;
                if (q != 0.0) roots[nr++] = Cq / q;
            }
        }
        for (int i = 0; i < nr; ++i) {
            const double sc = roots[i];
            if (!(sc > s && sc < best)) continue;
            const double z = o[2] + sc * d[2];
            const double r = sqrt(oo + 2.0 * od * sc + sc * sc);
            if (k != 0.0 && z * k < 0.0) continue;                    /* other nappe */
            const double dh = d[2] - k * (od + sc) / r;               /* d/ds (z - k r) */
            if ((side == 0 && dh > 0.0) || (side == 1 && dh < 0.0)) { best = sc; *face = 2 + side; }
        }
    }
    return best;
}

typedef struct { double s, r, lon, lat; double n[3]; double p[3]; int hit; long cells; } orc_hit;

/* first intersection of the body-frame ray o + s d (|d| = 1) for s in [s0, inf) */
static void trace(const orc_scene* S, const double* o, const double* d, double s0, orc_hit* out) {
    out->hit = 0; out->cells = 0; out->s = -1.0;
    const double Rb = S->radius * S->dmax;
    const double b = dot3(o, d), c = dot3(o, o) - Rb * Rb;
    const double disc = b * b - c;
    if (disc < 0.0) return;
    const double sq = sqrt(disc);
    double s_in = -b - sq, s_end = -b + sq;
    if (s_end <= s0) return;
    double s = s_in > s0 ? s_in : s0;
    /* The first cell is found from the position just inside; after that the walk is handed from
     * cell to cell across the wall it leaves through (a ray nearly tangent to a wall changes its
     * texel coordinate by less than atan2's rounding per step, so re-locating by position can
     * stall in the cell just left). */
    int r0, c0;
    {
        const double sp = s + 1e-9;
        const double p[3] = {o[0] + sp * d[0], o[1] + sp * d[1], o[2] + sp * d[2]};
        double u, v, lon, lat;
        point_uv(S, p, &u, &v, &lon, &lat);
        cell_of(S, u, v, &r0, &c0);
    }
    int face = 4;
    for (long guard = 0; guard < 4000000 && s < s_end; ++guard) {
        double u, v, lon, lat;
        if (guard > 0) {
            if (face == 0) c0 = c0 == 0 ? S->W - 1 : c0 - 1;
            else if (face == 1) c0 = c0 + 1 == S->W ? 0 : c0 + 1;
            else if (face == 2) r0 -= 1;
            else if (face == 3) r0 += 1;
            else break;
            if (r0 < 0 || r0 > S->H - 2) break;            /* cannot happen: polar caps have no wall */
        }
        double sx = cell_exit(S, o, d, r0, c0, s, s_end, &face);
        if (!(sx > s)) sx = s;
        if (sx - s > 1e-7) {
            /* self-check: the middle of the piece must lie in this cell */
            const double sm = 0.5 * (s + sx);
            const double pm[3] = {o[0] + sm * d[0], o[1] + sm * d[1], o[2] + sm * d[2]};
            int rr, cc;
            point_uv(S, pm, &u, &v, &lon, &lat);
            cell_of(S, u, v, &rr, &cc);
            if (rr != r0 || cc != c0) { out->cells = -1000000000L; }
        }
        patch_t P;
        load_patch(S, r0, c0, &P);
        out->cells++;
        if (getenv("ORC_DEBUG")) fprintf(stderr, "orc cell r%d c%d s=%.9f sx=%.9f\n", r0, c0, s, sx);
        /* cheap cull: the ray stays above the patch's highest corner over [s, sx] */
        double dm = P.d00 > P.d01 ? P.d00 : P.d01;
        if (P.d10 > dm) dm = P.d10;
        if (P.d11 > dm) dm = P.d11;
        double rmin2;
        {   /* min |p|^2 over [s, sx] */
            double sm = -b;
            if (sm < s) sm = s;
            if (sm > sx) sm = sx;
            rmin2 = dot3(o, o) + 2.0 * b * sm + sm * sm;
        }
        const double rc = S->radius * dm;
        if (rmin2 <= rc * rc * (1.0 + 1e-12)) {
            fctx_t F = {S, o, d, &P};
            double sh;
            if (patch_first_root(&F, s, sx, &sh)) {
                out->hit = 1; out->s = sh;
                const double q[3] = {o[0] + sh * d[0], o[1] + sh * d[1], o[2] + sh * d[2]};
                double fc, fr;
                point_uv(S, q, &u, &v, &lon, &lat);
                patch_frac(S, &P, u, v, &fc, &fr);
                out->r = sqrt(dot3(q, q)); out->lon = lon; out->lat = lat;
                memcpy(out->p, q, sizeof(q));
                /* normal of r(lon, lat) = R * D: n ~ e_r - (r_lon / (r cos lat)) e_lon - (r_lat / r) e_lat */
                const double dD_dfc = (P.d01 - P.d00) * (1.0 - fr) + (P.d11 - P.d10) * fr;
                double dD_dfr = (P.d10 - P.d00) * (1.0 - fc) + (P.d11 - P.d01) * fc;
                const double vv = v - P.r0;
                if (vv <= 0.0 || vv >= 1.0) dD_dfr = 0.0;           /* clamped rows: flat in latitude */
                const double r_lon = S->radius * dD_dfc * S->W / (2.0 * PI);
                const double r_lat = -S->radius * dD_dfr * S->H / PI;
                const double cl = cos(lat), sl = sin(lat), co = cos(lon), so = sin(lon);
                const double er[3] = {cl * so, -cl * co, sl};
                const double el[3] = {co, so, 0.0};
                const double ep[3] = {-sl * so, sl * co, cl};
                const double clc = cl > 1e-12 ? cl : 1e-12;
                double n[3];
                for (int i = 0; i < 3; ++i) n[i] = er[i] - (r_lon / (out->r * clc)) * el[i] - (r_lat / out->r) * ep[i];
                const double nn = sqrt(dot3(n, n));
                for (int i = 0; i < 3; ++i) out->n[i] = n[i] / nn;
                return;
            }
        }
        s = sx;
    }
}

/* ---- sampling ---------------------------------------------------------------------- */
static uint32_t hash_u32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
static double rnd(uint32_t pixel, uint32_t sample, uint32_t dim) {
    const uint32_t h = hash_u32(pixel ^ hash_u32(sample * 4u + dim + 0x9e3779b9u));
    return (double)(h >> 8) * (1.0 / 16777216.0);
}

static void to_body(const orc_scene* S, const double* v, double* out) {
    out[0] = dot3(S->ex, v); out[1] = dot3(S->ey, v); out[2] = dot3(S->ez, v);
}

static void sample_albedo(const orc_scene* S, double lon, double lat, double* rgb) {
    if (!S->tex) { rgb[0] = rgb[1] = rgb[2] = 1.0; return; }
    const int w = S->tex_w, h = S->tex_h;
    double u = (lon / (2.0 * PI) + 0.5) * w - 0.5, v = (0.5 - lat / PI) * h - 0.5;
    double fu = floor(u);
    int c0 = (int)fu; double fc = u - fu;
    if (c0 < 0) c0 += w;
    if (c0 >= w) c0 -= w;
    const int c1 = (c0 + 1) % w;
    int r0 = (int)floor(v);
    if (r0 < 0) r0 = 0;
    if (r0 > h - 2) r0 = h - 2;
    double fr = v - r0;
    if (fr < 0) fr = 0;
    if (fr > 1) fr = 1;
    for (int ch = 0; ch < 3; ++ch) {
        const double a = S->tex[((size_t)r0 * w + c0) * 4 + ch], bq = S->tex[((size_t)r0 * w + c1) * 4 + ch];
        const double cq = S->tex[((size_t)(r0 + 1) * w + c0) * 4 + ch], dq = S->tex[((size_t)(r0 + 1) * w + c1) * 4 + ch];
        rgb[ch] = ((a * (1 - fc) + bq * fc) * (1 - fr) + (cq * (1 - fc) + dq * fc) * fr) / 255.0;
    }
}

/* Radiance along a scene-space ray (origin o, unit direction d) that does not meet the Moon.
 * Sun disk: nearest intersection with the sphere (flat material: its colour, no lighting, moon_renderer.py:133-139).
 * Environment: direction -> (lon, lat) with the scene's +Z up and -Y at longitude 0 (the convention the reference uses
 * for the Moon itself, renderer_navigation.py:47-53; PlotOptiX's own TextureEnvironment mapping is closed - inferred),
 * u = (lon / 2 pi + 0.5) w - 0.5, v = (0.5 - lat / pi) h - 0.5, bilinear, columns wrap, rows clamp. */
static void miss_radiance(const orc_scene* S, const double* o, const double* d, double* rgb) {
    rgb[0] = rgb[1] = rgb[2] = 0.0;
    if (S->sun_disk_radius > 0.0) {
        const double oc[3] = {o[0] - S->sun_disk_pos[0], o[1] - S->sun_disk_pos[1], o[2] - S->sun_disk_pos[2]};
        const double b = dot3(oc, d), c = dot3(oc, oc) - S->sun_disk_radius * S->sun_disk_radius;
        const double disc = b * b - c;
        if (disc > 0.0 && -b + sqrt(disc) > 0.0) { for (int q = 0; q < 3; ++q) rgb[q] = S->sun_disk_color[q]; return; }
    }
    if (!S->env) return;
    const double lon = atan2(d[0], -d[1]), lat = asin(d[2] > 1.0 ? 1.0 : (d[2] < -1.0 ? -1.0 : d[2]));
    const int w = S->env_w, h = S->env_h;
    const double u = (lon / (2.0 * PI) + 0.5) * w - 0.5, v = (0.5 - lat / PI) * h - 0.5;
    const double fu = floor(u);
    int c0 = (int)fu;
    const double fc = u - fu;
    c0 = ((c0 % w) + w) % w;
    const int c1 = (c0 + 1) % w;
    int r0 = (int)floor(v);
    if (r0 < 0) r0 = 0;
    if (r0 > h - 2) r0 = h - 2;
    double fr = v - r0;
    if (fr < 0.0) fr = 0.0;
    if (fr > 1.0) fr = 1.0;
    for (int q = 0; q < 3; ++q) {
        const double a = S->env[((size_t)r0 * w + c0) * 4 + q], b = S->env[((size_t)r0 * w + c1) * 4 + q];
        const double c = S->env[((size_t)(r0 + 1) * w + c0) * 4 + q], e = S->env[((size_t)(r0 + 1) * w + c1) * 4 + q];
        rgb[q] = ((a * (1.0 - fc) + b * fc) * (1.0 - fr) + (c * (1.0 - fc) + e * fc) * fr) / 255.0;
    }
}

/* direct light of the sun at a hit (Lambert, one sample of the sun's disk, one shadow ray): rgb = albedo * E * visibility */
static void direct_light(const orc_scene* S, const orc_hit* h, const double* Lb, uint32_t pixel, unsigned sm, unsigned dim0,
                         double* rgb, double* alb, long* shadow_cells) {
    rgb[0] = rgb[1] = rgb[2] = 0.0;
    sample_albedo(S, h->lon, h->lat, alb);
    double tp[3] = {Lb[0] - h->p[0], Lb[1] - h->p[1], Lb[2] - h->p[2]};
    const double dist = sqrt(dot3(tp, tp));
    double lc[3] = {tp[0] / dist, tp[1] / dist, tp[2] / dist};
    double target[3] = {Lb[0], Lb[1], Lb[2]};
    if (S->jitter && S->light_radius > 0.0) {
        /* uniform point on the disk facing the hit (branchless ONB, Duff et al. 2017) */
        const double sg = lc[2] >= 0.0 ? 1.0 : -1.0, a = -1.0 / (sg + lc[2]), bq = lc[0] * lc[1] * a;
        const double b1[3] = {1.0 + sg * lc[0] * lc[0] * a, sg * bq, -sg * lc[0]};
        const double b2[3] = {bq, sg + lc[1] * lc[1] * a, -lc[1]};
        const double rr = S->light_radius * sqrt(rnd(pixel, sm, dim0)), th = 2.0 * PI * rnd(pixel, sm, dim0 + 1u);
        for (int q = 0; q < 3; ++q) target[q] += rr * (cos(th) * b1[q] + sin(th) * b2[q]);
    }
    double l[3] = {target[0] - h->p[0], target[1] - h->p[1], target[2] - h->p[2]};
    const double ln = sqrt(dot3(l, l));
    for (int q = 0; q < 3; ++q) l[q] /= ln;
    const double cosl = dot3(h->n, l);
    if (!(cosl > 0.0)) return;
    double vis = 1.0;
    if (S->shadows) {
        double so[3] = {h->p[0] + S->scene_epsilon * h->n[0], h->p[1] + S->scene_epsilon * h->n[1], h->p[2] + S->scene_epsilon * h->n[2]};
        orc_hit sh;
        trace(S, so, l, 0.0, &sh);
        *shadow_cells = sh.cells;
        if (sh.hit) vis = 0.0;
    }
    const double E = S->light_radiance * (S->light_radius / dist) * (S->light_radius / dist) * cosl * vis;
    for (int q = 0; q < 3; ++q) rgb[q] = alb[q] * E;
}

/* first intersection of the ray o + t d (|d| = 1) with the capsule (a, b, r): negative = none.  Brute force over all tubes. */
static double capsule_hit(const double* o, const double* d, const float* seg) {
    const double a[3] = {seg[0], seg[1], seg[2]}, r = seg[3];
    const double ba[3] = {seg[4] - a[0], seg[5] - a[1], seg[6] - a[2]}, oa[3] = {o[0] - a[0], o[1] - a[1], o[2] - a[2]};
    const double baba = dot3(ba, ba), bard = dot3(ba, d), baoa = dot3(ba, oa), rdoa = dot3(d, oa), oaoa = dot3(oa, oa);
    const double qa = baba - bard * bard, qb = baba * rdoa - baoa * bard, qc = baba * oaoa - baoa * baoa - r * r * baba;
    double best = -1.0;
    if (qa > 1.0e-12 * baba) {
        const double h = qb * qb - qa * qc;
        if (h >= 0.0) {
            const double t = (-qb - sqrt(h)) / qa, y = baoa + t * bard;
            if (y > 0.0 && y < baba) return t;
        }
    }
    {
        const double h = rdoa * rdoa - (oaoa - r * r);
        if (h > 0.0) best = -rdoa - sqrt(h);
    }
    {
        const double ob[3] = {oa[0] - ba[0], oa[1] - ba[1], oa[2] - ba[2]};
        const double B = dot3(d, ob), h = B * B - (dot3(ob, ob) - r * r);
        if (h > 0.0) { const double t = -B - sqrt(h); if (t > 0.0 && (best <= 0.0 || t < best)) best = t; }
    }
    return best;
}

static int nearest_tube(const orc_scene* S, const double* o, const double* d, double s_max, double* s_hit) {
    int which = -1;
    double best = s_max;
    for (int i = 0; i < S->n_tubes; ++i) {
        const double t = capsule_hit(o, d, S->tubes + (size_t)i * 12);
        if (t > 0.0 && t < best) { best = t; which = i; }
    }
    *s_hit = best;
    return which;
}

/*
 * Render samples sample0 .. sample0+nsamples-1 of the pixels (x0 + i*stride, y0 + j*stride)
 * inside [x0,x1) x [y0,y1).  Outputs are compact arrays over that sub-grid, row-major:
 *   accum  float64 [n][4]  sum r,g,b,count
 *   hit64  float64 [n][4]  (s_hit, radius, lon, lat) of the LAST sample, s < 0 = miss
 *   hit32  float32 [n][4]  scene-space hit x,y,z and distance of sample0 (0 = miss)
 *   stats  int64   [n][2]  cells walked by the primary / shadow ray of the last sample
 * Any output pointer may be null.  Returns the number of pixels rendered.
 */
long orc_render(const orc_scene* S, int x0, int y0, int x1, int y1, int stride,
                unsigned sample0, unsigned nsamples, double* accum, double* hit64, float* hit32, int64_t* stats) {
    const int nx = (x1 - x0 + stride - 1) / stride, ny = (y1 - y0 + stride - 1) / stride;
    const double aspect = (double)S->img_w / (double)S->img_h;
    double eye_rel[3] = {S->eye[0] - S->pos[0], S->eye[1] - S->pos[1], S->eye[2] - S->pos[2]};
    double ob[3], lrel[3] = {S->light_pos[0] - S->pos[0], S->light_pos[1] - S->pos[1], S->light_pos[2] - S->pos[2]}, Lb[3];
    to_body(S, eye_rel, ob);
    to_body(S, lrel, Lb);
#pragma omp parallel for schedule(dynamic, 4)
    for (int j = 0; j < ny; ++j) {
        for (int i = 0; i < nx; ++i) {
            const int x = x0 + i * stride, y = y0 + j * stride;
            const size_t k = (size_t)j * nx + i;
            const uint32_t pixel = (uint32_t)y * (uint32_t)S->img_w + (uint32_t)x;
            double acc[4] = {0, 0, 0, 0};
            for (unsigned sm = sample0; sm < sample0 + nsamples; ++sm) {
                const double jx = S->jitter ? rnd(pixel, sm, 0) : 0.5, jy = S->jitter ? rnd(pixel, sm, 1) : 0.5;
                const double sx = ((x + jx) / S->img_w * 2.0 - 1.0) * S->tan_half_fov * aspect;
                const double sy = (1.0 - (y + jy) / S->img_h * 2.0) * S->tan_half_fov;
                double dir[3], db[3];
                for (int a = 0; a < 3; ++a) dir[a] = S->w[a] + sx * S->right[a] + sy * S->up[a];
                const double dn = sqrt(dot3(dir, dir));
                for (int a = 0; a < 3; ++a) dir[a] /= dn;
                to_body(S, dir, db);
                orc_hit h;
                trace(S, ob, db, 0.0, &h);
                long shadow_cells = 0;
                double rgb[3] = {0, 0, 0};
                double s_tube = 0.0;
                const int tube = S->n_tubes ? nearest_tube(S, S->eye, dir, h.hit ? h.s : 1.0e300, &s_tube) : -1;
                if (tube >= 0) {
                    for (int q = 0; q < 3; ++q) rgb[q] = S->tubes[(size_t)tube * 12 + 8 + q];
                } else if (h.hit) {
                    /* the path: direct light at every hit, weighted with the product of the albedos before it; the next
                     * hit lies in a cosine-distributed direction about the normal (density and Lambert term cancel) */
                    double thr[3] = {1.0, 1.0, 1.0};
                    orc_hit cur = h;
                    for (int depth = 0; ; ++depth) {
                        double direct[3], alb[3];
                        long sc = 0;
                        const unsigned dim0 = 2u + 5u * (unsigned)depth;      /* light 2, bounce 2, roulette 1 */
                        direct_light(S, &cur, Lb, pixel, sm, dim0, direct, alb, &sc);
                        if (depth == 0) shadow_cells = sc;
                        for (int q = 0; q < 3; ++q) rgb[q] += thr[q] * direct[q];
                        if (depth >= S->n_bounce) break;
                        /* Russian roulette: go on with probability p = largest albedo component, carry albedo / p */
                        double p = alb[0] > alb[1] ? alb[0] : alb[1];
                        if (alb[2] > p) p = alb[2];
                        if (p > 1.0) p = 1.0;
                        if (!(p > 0.0) || !((double)(float)rnd(pixel, sm, dim0 + 4u) < (double)(float)p)) break;
                        for (int q = 0; q < 3; ++q) thr[q] *= alb[q] / p;
                        if (!(thr[0] > 0.0 || thr[1] > 0.0 || thr[2] > 0.0)) break;
                        const double u1 = rnd(pixel, sm, dim0 + 2u), u2 = rnd(pixel, sm, dim0 + 3u);
                        const double rr = sqrt(u1), cz = sqrt(1.0 - u1 > 0.0 ? 1.0 - u1 : 0.0), th = 2.0 * PI * u2;
                        const double* n = cur.n;
                        const double sg = n[2] >= 0.0 ? 1.0 : -1.0, a = -1.0 / (sg + n[2]), bq = n[0] * n[1] * a;
                        const double b1[3] = {1.0 + sg * n[0] * n[0] * a, sg * bq, -sg * n[0]};
                        const double b2[3] = {bq, sg + n[1] * n[1] * a, -n[1]};
                        double bd[3], bo[3];
                        for (int q = 0; q < 3; ++q) bd[q] = rr * cos(th) * b1[q] + rr * sin(th) * b2[q] + cz * n[q];
                        const double bn = sqrt(dot3(bd, bd));
                        for (int q = 0; q < 3; ++q) { bd[q] /= bn; bo[q] = cur.p[q] + S->scene_epsilon * n[q]; }
                        orc_hit nx;
                        trace(S, bo, bd, 0.0, &nx);
                        if (!nx.hit) break;
                        cur = nx;
                    }
                } else {
                    miss_radiance(S, S->eye, dir, rgb);
                }
                acc[0] += rgb[0]; acc[1] += rgb[1]; acc[2] += rgb[2]; acc[3] += 1.0;
                if (hit64) {
                    hit64[k * 4 + 0] = h.hit ? h.s : -1.0; hit64[k * 4 + 1] = h.hit ? h.r : 0.0;
                    hit64[k * 4 + 2] = h.hit ? h.lon : 0.0; hit64[k * 4 + 3] = h.hit ? h.lat : 0.0;
                    if (tube >= 0) { hit64[k * 4 + 0] = -2.0; hit64[k * 4 + 1] = 0.0; hit64[k * 4 + 2] = 0.0; hit64[k * 4 + 3] = s_tube; }
                }
                if (hit32 && sm == sample0) {
                    for (int q = 0; q < 3; ++q) {
                        /* scene = pos + R^T p_body */
                        double sc = h.hit ? S->pos[q] + S->ex[q] * h.p[0] + S->ey[q] * h.p[1] + S->ez[q] * h.p[2] : 0.0;
                        if (tube >= 0) sc = S->eye[q] + s_tube * dir[q];
                        hit32[k * 4 + q] = (float)sc;
                    }
                    hit32[k * 4 + 3] = tube >= 0 ? (float)s_tube : (h.hit ? (float)h.s : 0.0f);
                }
                if (stats) { stats[k * 2] = h.cells; stats[k * 2 + 1] = shadow_cells; }
            }
            if (accum) for (int q = 0; q < 4; ++q) accum[k * 4 + q] = acc[q];
        }
    }
    return (long)nx * ny;
}

/* single body-frame ray, for the analytic known-answer tests: out = s, r, lon, lat, nx, ny, nz, cells */
int orc_trace_ray(const orc_scene* S, const double* o, const double* d, double* out8) {
    double dn = sqrt(dot3(d, d)), du[3] = {d[0] / dn, d[1] / dn, d[2] / dn};
    orc_hit h;
    trace(S, o, du, 0.0, &h);
    out8[0] = h.hit ? h.s : -1.0; out8[1] = h.r; out8[2] = h.lon; out8[3] = h.lat;
    out8[4] = h.n[0]; out8[5] = h.n[1]; out8[6] = h.n[2]; out8[7] = (double)h.cells;
    return h.hit;
}

/* Gamma post-process to 8 bit (moon_renderer.py:597-600): in = accum [n][4], out = rgba [n][4] */
void orc_tonemap(const orc_scene* S, const double* accum, long n, uint8_t* rgba) {
    for (long k = 0; k < n; ++k) {
        const double wgt = accum[k * 4 + 3] > 0 ? accum[k * 4 + 3] : 1.0;
        for (int q = 0; q < 3; ++q) {
            double c = (double)S->exposure * accum[k * 4 + q] / wgt;
            c = c > 0.0 ? pow(c, (double)S->inv_gamma) : 0.0;
            double v = floor(c * 255.0 + 0.5);
            rgba[k * 4 + q] = (uint8_t)(v > 255.0 ? 255.0 : v);
        }
        rgba[k * 4 + 3] = 255;
    }
}

int orc_scene_size(void) { return (int)sizeof(orc_scene); }
