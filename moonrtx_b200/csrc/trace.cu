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

#include "common.cuh"

namespace {

constexpr double PI_D = 3.14159265358979323846;
constexpr float  PI_F = 3.14159265358979323846f;
constexpr int    MAX_STEPS = 60000;

struct Counters { unsigned nodes, tests, overflow; };

// ---- height field access ---------------------------------------------------------------
template <bool I16>
__device__ __forceinline__ float texel_D(const HeightField& hf, int r, int c) {
    if (I16) {
        const float v = (float)__ldg((const int16_t*)hf.base + (size_t)r * hf.W + c);
        // exactly data_loader.py:219-242: *scale, +1, /radius_scale, one rounding each
        return __fdiv_rn(__fadd_rn(__fmul_rn(v, hf.scale), 1.0f), hf.radius_scale);
    }
    return __ldg((const float*)hf.base + (size_t)r * hf.W + c);
}

template <bool I16>
__device__ __forceinline__ float level_D(const HeightField& hf, int L, int J, int I) {
    if (I16) {
        const float v = (float)__ldg((const int16_t*)hf.level[L] + (size_t)J * hf.nx[L] + I);
        return __fdiv_rn(__fadd_rn(__fmul_rn(v, hf.scale), 1.0f), hf.radius_scale);
    }
    return __ldg((const float*)hf.level[L] + (size_t)J * hf.nx[L] + I);
}

struct Patch { int r0, c0; float d00, d01, d10, d11; };

template <bool I16>
__device__ __forceinline__ void load_patch(const HeightField& hf, int r0, int c0, Patch& P) {
    const int c1 = c0 + 1 == hf.W ? 0 : c0 + 1;
    P.r0 = r0; P.c0 = c0;
    P.d00 = texel_D<I16>(hf, r0, c0);     P.d01 = texel_D<I16>(hf, r0, c1);
    P.d10 = texel_D<I16>(hf, r0 + 1, c0); P.d11 = texel_D<I16>(hf, r0 + 1, c1);
}

// ---- float64 exact patch test -------------------------------------------------------------
struct Ray64 { double ox, oy, oz, dx, dy, dz, oo, od; };

struct HitInfo { double s, r, lon, lat, fc, fr; };

__device__ __forceinline__ double patch_f(const Ray64& R, const Patch& P, int W, int H, double radius, double s,
                                          HitInfo* info) {
    const double x = R.ox + s * R.dx, y = R.oy + s * R.dy, z = R.oz + s * R.dz;
    const double rho2 = x * x + y * y;
    const double r = sqrt(rho2 + z * z);
    const double lon = atan2(x, -y), lat = atan2(z, sqrt(rho2));
    const double u = (lon * (0.5 / PI_D) + 0.5) * W - 0.5, v = (0.5 - lat * (1.0 / PI_D)) * H - 0.5;
    double fc = u - P.c0;
    if (fc < -0.5 * W) fc += W;
    if (fc > 0.5 * W) fc -= W;
    double fr = v - P.r0;
    fr = fr < 0.0 ? 0.0 : (fr > 1.0 ? 1.0 : fr);
    const double d = (double)P.d00 * (1.0 - fr) * (1.0 - fc) + (double)P.d10 * fr * (1.0 - fc) +
                     (double)P.d01 * (1.0 - fr) * fc + (double)P.d11 * fr * fc;
    if (info) { info->s = s; info->r = r; info->lon = lon; info->lat = lat; info->fc = fc; info->fr = v - P.r0; }
    return r - radius * d;
}

// first root of f on [a, b] inside one patch; lo/hi bracket polished to ~1e-13
__device__ bool patch_root(const Ray64& R, const Patch& P, int W, int H, double radius, double a, double b, double* s_hit) {
    double fa = patch_f(R, P, W, H, radius, a, nullptr);
    if (fa <= 0.0) { *s_hit = a; return true; }
    double fb = patch_f(R, P, W, H, radius, b, nullptr);
    double lo = a, flo = fa, hi = b, fhi = fb;
    if (fb > 0.0) {
        // no sign change at the ends: a grazing double root shows up as a dip in between
        const double m = 0.5 * (a + b);
        const double fm = patch_f(R, P, W, H, radius, m, nullptr);
        if (fm <= 0.0) { hi = m; fhi = fm; }
        else {
            // vertex of the parabola through (a, fa), (m, fm), (b, fb)
            const double h = 0.5 * (b - a);
            const double c2 = (fa - 2.0 * fm + fb) / (2.0 * h * h), c1 = (fb - fa) / (2.0 * h);
            if (!(c2 > 0.0)) return false;
            const double tv = m - c1 / (2.0 * c2);
            if (!(tv > a && tv < b)) return false;
            if (fm - c1 * c1 / (4.0 * c2) > 0.25 * fm + 1e-9) return false;      // the dip stays clear of zero
            const double fv = patch_f(R, P, W, H, radius, tv, nullptr);
            if (fv > 0.0) return false;
            hi = tv; fhi = fv;
        }
    }
    for (int it = 0; it < 100 && hi - lo > 1e-14 * (1.0 + fabs(hi)); ++it) {
        double m = (it & 1) ? 0.5 * (lo + hi) : lo + (hi - lo) * flo / (flo - fhi);
        if (!(m > lo && m < hi)) m = 0.5 * (lo + hi);
        const double fm = patch_f(R, P, W, H, radius, m, nullptr);
        if (fm > 0.0) { lo = m; flo = fm; } else { hi = m; fhi = fm; }
    }
    *s_hit = hi;
    return true;
}

// The exact interval(s) of the ray inside cell (r0, c0) within the window [wa, wb], each
// searched for a root in order.  Cell walls: lon half-planes g = p.t (t = (cos lam, sin lam, 0)),
// lat cones h = z - k r (k = sin phi).
__device__ bool cell_test64(const Ray64& R, const Patch& P, int W, int H, double radius, double wa, double wb,
                            double* s_hit) {
    double crit[10];
    int n = 0;
    crit[n++] = wa;
    double tx[2], ty[2], kk[2];
    bool has_lat[2];
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        double sn, cs;
        sincospi((2.0 * (P.c0 + side) + 1.0) / W - 1.0, &sn, &cs);
        tx[side] = cs; ty[side] = sn;
        const double g0 = R.ox * cs + R.oy * sn, g1 = R.dx * cs + R.dy * sn;
        if (g1 != 0.0) {
            const double sc = -g0 / g1;
            if (sc > wa && sc < wb) crit[n++] = sc;
        }
    }
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        has_lat[side] = side == 0 ? (P.r0 > 0) : (P.r0 < H - 2);
        kk[side] = 0.0;
        if (!has_lat[side]) continue;
        const double k = cospi((P.r0 + side + 0.5) / H);        // sin(phi) of the wall
        kk[side] = k;
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
        if (r1 > wa && r1 < wb) crit[n++] = r1;
        if (r2 > wa && r2 < wb) crit[n++] = r2;
    }
    crit[n++] = wb;
    for (int i = 1; i < n; ++i) {                                // insertion sort, n <= 8
        const double key = crit[i];
        int j = i - 1;
        while (j >= 0 && crit[j] > key) { crit[j + 1] = crit[j]; --j; }
        crit[j + 1] = key;
    }
    for (int i = 0; i + 1 < n; ++i) {
        const double a = crit[i], b = crit[i + 1];
        if (!(b > a)) continue;
        const double m = 0.5 * (a + b);
        const double x = R.ox + m * R.dx, y = R.oy + m * R.dy, z = R.oz + m * R.dz;
        if (x * tx[0] + y * ty[0] < 0.0) continue;               // west of the cell
        if (x * tx[1] + y * ty[1] > 0.0) continue;               // east of it
        const double r = sqrt(x * x + y * y + z * z);
        if (has_lat[0] && z - kk[0] * r > 0.0) continue;         // north of it
        if (has_lat[1] && z - kk[1] * r < 0.0) continue;         // south of it
        if (patch_root(R, P, W, H, radius, a, b, s_hit)) return true;
    }
    return false;
}

// ---- float32 pyramid traversal ---------------------------------------------------------------
struct Trav {
    float ox, oy, oz, dx, dy, dz, oo, od;    // re-based ray
    float smax;
};

__device__ __forceinline__ float ray_r2(const Trav& T, float s) { return fmaf(s, fmaf(2.0f, T.od, s), T.oo); }

// exit parameter + face (0 lon-lo, 1 lon-hi, 2 north, 3 south, 4 end of ray) of the level-L cell (J, I)
__device__ float cell_exit32(const HeightField& hf, const Trav& T, int L, int J, int I, float s, int& face) {
    const int W = hf.W, H = hf.H;
    const int a = I << L, b = min((I + 1) << L, W);
    const int n = J << L, m = min((J + 1) << L, H - 1);
    float best = T.smax;
    face = 4;
    const float invW = 1.0f / (float)W, invH = 1.0f / (float)H;
    {   // east wall: leaving when g = p.t goes positive
        float sn, cs;
        sincospif((float)(2 * b + 1) * invW - 1.0f, &sn, &cs);
        const float g1 = T.dx * cs + T.dy * sn;
        if (g1 > 0.0f) {
            const float sc = -(T.ox * cs + T.oy * sn) / g1;
            const float px = fmaf(sc, T.dx, T.ox), py = fmaf(sc, T.dy, T.oy);
            if (sc < best && px * sn - py * cs > 0.0f) { best = sc; face = 1; }
        }
    }
    {   // west wall: leaving when g goes negative
        float sn, cs;
        sincospif((float)(2 * a + 1) * invW - 1.0f, &sn, &cs);
        const float g1 = T.dx * cs + T.dy * sn;
        if (g1 < 0.0f) {
            const float sc = -(T.ox * cs + T.oy * sn) / g1;
            const float px = fmaf(sc, T.dx, T.ox), py = fmaf(sc, T.dy, T.oy);
            if (sc < best && px * sn - py * cs > 0.0f) { best = sc; face = 0; }
        }
    }
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        if (side == 0 ? (n == 0) : (m == H - 1)) continue;      // polar caps have no wall
        const float k = cospif(((float)(side == 0 ? n : m) + 0.5f) * invH);
        // already beyond the wall?  (h = z - k r; north wall: outside when h > 0)
        const float zs = fmaf(s, T.dz, T.oz), rs = sqrtf(fmaxf(ray_r2(T, s), 0.0f));
        const float hs = zs - k * rs;
        if (side == 0 ? hs > 0.0f : hs < 0.0f) { if (s < best) { best = s; face = 2 + side; } continue; }
        const float k2 = k * k;
        const float A = T.dz * T.dz - k2, B = T.oz * T.dz - k2 * T.od, Cq = T.oz * T.oz - k2 * T.oo;
        float r1 = -1.0f, r2 = -1.0f;
        if (fabsf(A) < 1e-12f) { if (B != 0.0f) r1 = -Cq / (2.0f * B); }
        else {
            const float disc = B * B - A * Cq;
            if (disc >= 0.0f) {
                const float q = -(B + copysignf(sqrtf(disc), B));
                r1 = q / A;
                if (q != 0.0f) r2 = Cq / q;
            }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float sc = i ? r2 : r1;
            if (!(sc > s) || !(sc < best)) continue;
            const float z = fmaf(sc, T.dz, T.oz);
            if (k != 0.0f && z * k < 0.0f) continue;             // the cone's other nappe
            const float r = sqrtf(fmaxf(ray_r2(T, sc), 1e-30f));
            const float dh = T.dz - k * (T.od + sc) / r;
            if (side == 0 ? dh > 0.0f : dh < 0.0f) { best = sc; face = 2 + side; }
        }
    }
    return best;
}

struct TraceOut { bool hit; double s; HitInfo info; Patch patch; };

// First intersection of the body-frame ray for s >= s_min.  any_hit: stop at any intersection.
template <bool I16>
__device__ void trace_ray(const HeightField& hf, double radius, const Ray64& R, double s_min, bool any_hit,
                          TraceOut& out, Counters& cnt) {
    out.hit = false;
    const double Rb = radius * (double)hf.dmax;
    const double disc = R.od * R.od - (R.oo - Rb * Rb);
    if (disc < 0.0) return;
    const double sq = sqrt(disc);
    const double s_end = -R.od + sq;
    if (s_end <= s_min) return;
    const double s_in = fmax(s_min, -R.od - sq);

    Trav T;
    {
        const double bx = R.ox + s_in * R.dx, by = R.oy + s_in * R.dy, bz = R.oz + s_in * R.dz;
        T.ox = (float)bx; T.oy = (float)by; T.oz = (float)bz;
        T.dx = (float)R.dx; T.dy = (float)R.dy; T.dz = (float)R.dz;
        T.oo = T.ox * T.ox + T.oy * T.oy + T.oz * T.oz;
        T.od = T.ox * T.dx + T.oy * T.dy + T.oz * T.dz;
        T.smax = (float)(s_end - s_in);
    }
    const float Rf = (float)radius;
    const float marg = 2.0e-6f * Rf;                             // > float32 error of a radius near R
    const int W = hf.W, H = hf.H;

    // start cell at the top level, from the position just inside
    int L = hf.top, J, I;
    {
        const float t0 = fminf(1e-5f * Rf, 0.5f * T.smax);
        const float x = fmaf(t0, T.dx, T.ox), y = fmaf(t0, T.dy, T.oy), z = fmaf(t0, T.dz, T.oz);
        const float lon = atan2f(x, -y), lat = atan2f(z, sqrtf(x * x + y * y));
        const float u = (lon * (0.5f / PI_F) + 0.5f) * (float)W - 0.5f, v = (0.5f - lat * (1.0f / PI_F)) * (float)H - 0.5f;
        int c0 = (int)floorf(u);
        c0 = c0 < 0 ? c0 + W : (c0 >= W ? c0 - W : c0);
        const int r0 = min(max((int)floorf(v), 0), H - 2);
        J = r0 >> L; I = c0 >> L;
    }

    float s = 0.0f;
    for (int step = 0; step < MAX_STEPS; ++step) {
        int face;
        const float sx_raw = cell_exit32(hf, T, L, J, I, s, face);
        const float sx = fmaxf(sx_raw, s);
        ++cnt.nodes;

        // max radius of the surface over this cell
        float dmax;
        Patch P;
        if (L == 0) {
            load_patch<I16>(hf, J, I, P);
            dmax = fmaxf(fmaxf(P.d00, P.d01), fmaxf(P.d10, P.d11));
        } else {
            dmax = level_D<I16>(hf, L, J, I);
        }
        const float rc = fmaf(Rf, dmax, marg);
        // min radius of the ray over [s, sx] (with a little slack either side)
        const float pad = 4.0e-6f * Rf;
        const float ta = fmaxf(s - pad, 0.0f), tb = fminf(sx + pad, T.smax);
        const float tm = fminf(fmaxf(-T.od, ta), tb);
        const float rmin2 = ray_r2(T, tm);

        bool advance = true;
        if (rmin2 <= rc * rc) {
            if (L > 0) {
                // move up to where the ray enters the cell's shell, then pick the child there
                float sd = s;
                if (ray_r2(T, s) > rc * rc) {
                    const float dq = T.od * T.od - (T.oo - rc * rc);
                    if (dq > 0.0f) sd = fminf(fmaxf(-T.od - sqrtf(dq), s), sx);
                }
                const float x = fmaf(sd, T.dx, T.ox), y = fmaf(sd, T.dy, T.oy), z = fmaf(sd, T.dz, T.oz);
                const int mi = (2 * I + 1) << (L - 1), mj = (2 * J + 1) << (L - 1);
                int ci = 2 * I, cj = 2 * J;
                if (mi < min((I + 1) << L, W)) {
                    float sn, cs;
                    sincospif((float)(2 * mi + 1) / (float)W - 1.0f, &sn, &cs);
                    if (x * cs + y * sn >= 0.0f) ci += 1;
                }
                if (mj < min((J + 1) << L, H - 1)) {
                    const float k = cospif(((float)mj + 0.5f) / (float)H);
                    if (z - k * sqrtf(x * x + y * y + z * z) < 0.0f) cj += 1;     // south of the mid wall
                }
                s = sd; L -= 1; I = ci; J = cj;
                advance = false;
            } else {
                ++cnt.tests;
                const double wpad = 2.0e-5 * radius;
                const double wa = fmax(s_in + (double)s - wpad, s_min), wb = fmin(s_in + (double)sx + wpad, s_end);
                double sh;
                if (cell_test64(R, P, W, H, radius, wa, wb, &sh)) {
                    out.hit = true; out.s = sh; out.patch = P;
                    if (!any_hit) patch_f(R, P, W, H, radius, sh, &out.info);
                    return;
                }
            }
        }
        if (advance) {
            if (face == 4) return;                               // left the bounding sphere
            s = sx;
            bool up;
            if (face == 1)      { I += 1; if (I >= hf.nx[L]) I = 0; up = (I & 1) == 0; }
            else if (face == 0) { up = (I & 1) == 0; I -= 1; if (I < 0) I = hf.nx[L] - 1; }
            else if (face == 3) { J += 1; up = (J & 1) == 0; }
            else                { up = (J & 1) == 0; J -= 1; }
            if (J < 0 || J >= hf.ny[L]) return;                  // cannot happen (caps have no wall); be safe
            if (up && L < hf.top) { L += 1; I >>= 1; J >>= 1; }
        }
    }
    ++cnt.overflow;                                              // step budget exhausted: reported as a miss
}

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
};

template <bool I16>
__global__ void __launch_bounds__(128)
trace_kernel(const __grid_constant__ RenderArgs A) {
    const int x = A.x0 + blockIdx.x * blockDim.x + threadIdx.x;
    const int y = A.y0 + blockIdx.y * blockDim.y + threadIdx.y;
    const bool active = x < A.x1 && y < A.y1;
    Counters cnt = {0u, 0u, 0u};
    unsigned n_primary = 0, n_inside = 0, n_hit = 0, n_shadow = 0, n_occl = 0;
    if (active) {
        const SceneParams& sp = A.sp;
        const Camera& cam = A.cam;
        const uint32_t pixel = (uint32_t)y * (uint32_t)A.width + (uint32_t)x;
        const double aspect = (double)A.width / (double)A.height;
        // eye and light in the body frame
        const double er[3] = {cam.eye[0] - sp.pos[0], cam.eye[1] - sp.pos[1], cam.eye[2] - sp.pos[2]};
        const double lr[3] = {sp.light_pos[0] - sp.pos[0], sp.light_pos[1] - sp.pos[1], sp.light_pos[2] - sp.pos[2]};
        Ray64 R;
        R.ox = sp.ex[0] * er[0] + sp.ex[1] * er[1] + sp.ex[2] * er[2];
        R.oy = sp.ey[0] * er[0] + sp.ey[1] * er[1] + sp.ey[2] * er[2];
        R.oz = sp.ez[0] * er[0] + sp.ez[1] * er[1] + sp.ez[2] * er[2];
        R.oo = R.ox * R.ox + R.oy * R.oy + R.oz * R.oz;
        const double Lx = sp.ex[0] * lr[0] + sp.ex[1] * lr[1] + sp.ex[2] * lr[2];
        const double Ly = sp.ey[0] * lr[0] + sp.ey[1] * lr[1] + sp.ey[2] * lr[2];
        const double Lz = sp.ez[0] * lr[0] + sp.ez[1] * lr[1] + sp.ez[2] * lr[2];

        float3 acc = make_float3(0.f, 0.f, 0.f);
        for (unsigned sm = A.sample0; sm < A.sample0 + A.nsamples; ++sm) {
            const double jx = sp.jitter ? rnd(pixel, sm, 0) : 0.5, jy = sp.jitter ? rnd(pixel, sm, 1) : 0.5;
            const double sx = ((x + jx) / A.width * 2.0 - 1.0) * cam.tan_half_fov * aspect;
            const double sy = (1.0 - (y + jy) / A.height * 2.0) * cam.tan_half_fov;
            double d[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) d[a] = cam.w[a] + sx * cam.right[a] + sy * cam.up[a];
            const double dn = 1.0 / sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
#pragma unroll
            for (int a = 0; a < 3; ++a) d[a] *= dn;
            R.dx = sp.ex[0] * d[0] + sp.ex[1] * d[1] + sp.ex[2] * d[2];
            R.dy = sp.ey[0] * d[0] + sp.ey[1] * d[1] + sp.ey[2] * d[2];
            R.dz = sp.ez[0] * d[0] + sp.ez[1] * d[1] + sp.ez[2] * d[2];
            R.od = R.ox * R.dx + R.oy * R.dy + R.oz * R.dz;

            ++n_primary;
            TraceOut h;
            const unsigned nodes_before = cnt.nodes;
            trace_ray<I16>(A.hf, sp.radius, R, 0.0, false, h, cnt);
            if (cnt.nodes != nodes_before) ++n_inside;
            float3 rgb = make_float3(0.f, 0.f, 0.f);
            if (h.hit) {
                ++n_hit;
                const HitInfo& hi = h.info;
                const double px = R.ox + h.s * R.dx, py = R.oy + h.s * R.dy, pz = R.oz + h.s * R.dz;
                // normal of r(lon, lat) = R * D: n ~ e_r - (r_lon / (r cos lat)) e_lon - (r_lat / r) e_lat
                const Patch& P = h.patch;
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
                double tx = Lx - px, ty = Ly - py, tz = Lz - pz;
                const double dist = sqrt(tx * tx + ty * ty + tz * tz);
                double gx = Lx, gy = Ly, gz = Lz;
                if (sp.jitter && sp.light_radius > 0.0) {
                    const double cx = tx / dist, cy = ty / dist, cz = tz / dist;
                    const double sg = cz >= 0.0 ? 1.0 : -1.0, a = -1.0 / (sg + cz), b = cx * cy * a;
                    const double b1x = 1.0 + sg * cx * cx * a, b1y = sg * b, b1z = -sg * cx;
                    const double b2x = b, b2y = sg + cy * cy * a, b2z = -cy;
                    const double rr = sp.light_radius * sqrt(rnd(pixel, sm, 2)), th = 2.0 * PI_D * rnd(pixel, sm, 3);
                    double st, ct;
                    sincos(th, &st, &ct);
                    gx += rr * (ct * b1x + st * b2x); gy += rr * (ct * b1y + st * b2y); gz += rr * (ct * b1z + st * b2z);
                }
                double lx = gx - px, ly = gy - py, lz = gz - pz;
                const double ln = 1.0 / sqrt(lx * lx + ly * ly + lz * lz);
                lx *= ln; ly *= ln; lz *= ln;
                const double cosl = nx * lx + ny * ly + nz * lz;
                if (cosl > 0.0) {
                    double vis = 1.0;
                    if (sp.shadows) {
                        Ray64 S;
                        S.ox = px + sp.scene_epsilon * nx; S.oy = py + sp.scene_epsilon * ny; S.oz = pz + sp.scene_epsilon * nz;
                        S.dx = lx; S.dy = ly; S.dz = lz;
                        S.oo = S.ox * S.ox + S.oy * S.oy + S.oz * S.oz;
                        S.od = S.ox * S.dx + S.oy * S.dy + S.oz * S.dz;
                        TraceOut sh;
                        ++n_shadow;
                        trace_ray<I16>(A.hf, sp.radius, S, 0.0, true, sh, cnt);
                        if (sh.hit) { vis = 0.0; ++n_occl; }
                    }
                    const float3 alb = sample_albedo(A.tex, hi.lon, hi.lat);
                    const double q = sp.light_radius / dist;
                    const float E = (float)(sp.light_radiance * q * q * cosl * vis);
                    rgb = make_float3(alb.x * E, alb.y * E, alb.z * E);
                }
                if (sm == A.sample0 && A.hit) {
                    // scene = pos + R^T p_body
                    const float hx = (float)(sp.pos[0] + sp.ex[0] * px + sp.ey[0] * py + sp.ez[0] * pz);
                    const float hy = (float)(sp.pos[1] + sp.ex[1] * px + sp.ey[1] * py + sp.ez[1] * pz);
                    const float hz = (float)(sp.pos[2] + sp.ex[2] * px + sp.ey[2] * py + sp.ez[2] * pz);
                    A.hit[(size_t)y * A.width + x] = make_float4(hx, hy, hz, (float)h.s);
                }
                if (A.hit64) A.hit64[(size_t)y * A.width + x] = make_double4(h.s, hi.r, hi.lon, hi.lat);
            } else {
                if (sm == A.sample0 && A.hit) A.hit[(size_t)y * A.width + x] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (A.hit64) A.hit64[(size_t)y * A.width + x] = make_double4(-1.0, 0.0, 0.0, 0.0);
            }
            acc.x += rgb.x; acc.y += rgb.y; acc.z += rgb.z;
        }
        float4* ap = A.accum + (size_t)y * A.width + x;
        float4 old = *ap;
        old.x += acc.x; old.y += acc.y; old.z += acc.z; old.w += (float)A.nsamples;
        *ap = old;
    }
    // counters: one atomic per warp and counter
    unsigned vals[8] = {n_primary, n_inside, n_hit, n_shadow, n_occl, cnt.nodes, cnt.tests, cnt.overflow};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const unsigned v = __reduce_add_sync(0xffffffffu, vals[i]);
        if (((threadIdx.y * blockDim.x + threadIdx.x) & 31) == 0 && v) atomicAdd(&A.counters[i], (unsigned long long)v);
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

int launch_trace(mrtx_ctx* ctx, int x0, int y0, int x1, int y1, unsigned s0, unsigned ns) {
    RenderArgs A;
    A.hf = ctx->hf; A.tex = ctx->tex[0]; A.cam = ctx->cam; A.sp = ctx->sp;
    A.width = ctx->width; A.height = ctx->height;
    A.x0 = x0; A.y0 = y0; A.x1 = x1; A.y1 = y1;
    A.sample0 = s0; A.nsamples = ns;
    A.accum = ctx->accum; A.hit = ctx->hit;
    A.hit64 = ctx->sp.debug_hits ? ctx->hit64 : nullptr;
    A.counters = ctx->d_counters;
    const dim3 block(8, 16);
    const dim3 grid((x1 - x0 + block.x - 1) / block.x, (y1 - y0 + block.y - 1) / block.y);
    if (ctx->hf.is_i16) trace_kernel<true><<<grid, block, 0, ctx->stream>>>(A);
    else                trace_kernel<false><<<grid, block, 0, ctx->stream>>>(A);
    MRTX_CUDA(cudaGetLastError());
    return MRTX_OK;
}

int launch_resolve(mrtx_ctx* ctx) {
    const size_t n = (size_t)ctx->width * ctx->height;
    resolve_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ctx->accum, ctx->tex[1].data, ctx->rgba8, n,
                                                                ctx->sp.exposure, ctx->sp.inv_gamma);
    MRTX_CUDA(cudaGetLastError());
    return MRTX_OK;
}
