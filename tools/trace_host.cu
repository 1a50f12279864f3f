// DEVELOPMENT TOOL (not part of the product, never loaded by moonrtx_b200): runs the
// __host__ __device__ traversal core of csrc/trace_core.cuh on the CPU so that parity
// problems can be investigated against the oracle without a GPU.
#define MRTX_LEVEL_HIST 1
#include "../moonrtx_b200/csrc/trace_fast.cuh"
#include <vector>
#include <algorithm>

void mrtx_set_error(const char*, ...) {}

namespace mrtx_core { unsigned long long g_level_hist[32]; }
using namespace mrtx_core;

// node visits of the filtered walk by pyramid level since the last reset (walk_step, MRTX_LEVEL_HIST)
extern "C" void dbg_level_hist(unsigned long long* out32, int reset) {
    for (int i = 0; i < 32; ++i) { out32[i] = g_level_hist[i]; if (reset) g_level_hist[i] = 0; }
}

// all levels and their dilated copies in ONE buffer (the walk addresses them as element offsets from hf.lvl_base)
template <typename T>
static void build_host_pyramid(HeightField& hf, std::vector<T>& all) {
    const int W = hf.W, H = hf.H;
    int top = 0;
    while ((W >> (top + 1)) >= 64 && top + 1 < MRTX_MAX_LEVELS) ++top;
    hf.top = top; hf.nx[0] = W; hf.ny[0] = H - 1;
    size_t total = 0;
    for (int k = 1; k <= top; ++k) {
        hf.nx[k] = (W + (1 << k) - 1) >> k; hf.ny[k] = (H - 1 + (1 << k) - 1) >> k;
        hf.off[k] = (unsigned)total; total += (size_t)hf.nx[k] * hf.ny[k];
        hf.off[MRTX_MAX_LEVELS + k] = 0;
        if (k >= MRTX_DIL_MIN_LEVEL) { hf.off[MRTX_MAX_LEVELS + k] = (unsigned)total; total += (size_t)hf.nx[k] * hf.ny[k]; }
    }
    all.assign(total, T(0));
    hf.lvl_base = all.data();
    const T* base = (const T*)hf.base;
    for (int k = 1; k <= top; ++k) {
        T* out = all.data() + hf.off[k];
        hf.level[k] = out;
#pragma omp parallel for schedule(static)
        for (int J = 0; J < hf.ny[k]; ++J) for (int I = 0; I < hf.nx[k]; ++I) {
            T m;
            if (k == 1) {
                const int r_lo = 2 * J, r_hi = std::min(2 * J + 2, H - 1), c_lo = 2 * I, c_hi = std::min(2 * I + 2, W);
                m = base[(size_t)r_lo * W + c_lo];
                for (int r = r_lo; r <= r_hi; ++r) for (int c = c_lo; c <= c_hi; ++c) m = std::max(m, base[(size_t)r * W + (c >= W ? c - W : c)]);
            } else {
                const int inx = hf.nx[k - 1], iny = hf.ny[k - 1];
                const int r1 = std::min(2 * J + 1, iny - 1), c1 = std::min(2 * I + 1, inx - 1);
                const T* in = (const T*)hf.level[k - 1];
                m = std::max(std::max(in[(size_t)(2 * J) * inx + 2 * I], in[(size_t)(2 * J) * inx + c1]),
                             std::max(in[(size_t)r1 * inx + 2 * I], in[(size_t)r1 * inx + c1]));
            }
            out[(size_t)J * hf.nx[k] + I] = m;
        }
    }
    for (int k = MRTX_DIL_MIN_LEVEL; k <= top; ++k) {
        const int nx = hf.nx[k], ny = hf.ny[k];
        int reach = ((long long)nx << k) != W ? 2 : 1;
        if (reach > nx / 2) reach = nx / 2;
        const T* in = (const T*)hf.level[k];
        T* out = all.data() + hf.off[MRTX_MAX_LEVELS + k];
        hf.dil[k] = out;
#pragma omp parallel for schedule(static)
        for (int J = 0; J < ny; ++J) for (int I = 0; I < nx; ++I) {
            T m = in[(size_t)J * nx + I];
            for (int j = std::max(J - 1, 0); j <= std::min(J + 1, ny - 1); ++j) for (int d = -reach; d <= reach; ++d) {
                int i = I + d; i = i < 0 ? i + nx : (i >= nx ? i - nx : i);
                m = std::max(m, in[(size_t)j * nx + i]);
            }
            out[(size_t)J * nx + I] = m;
        }
    }
}

extern "C" void dbg_set(int v) { mrtx_core::g_debug = v; }

// pyramid + wall tables are cached across calls (keyed on the map pointer): they take minutes at 92160 x 46080
struct HostScene {
    const void* key = nullptr; int W = 0, H = 0;
    HeightField hf;
    std::vector<int16_t> s16; std::vector<float> s32;
    std::vector<float2> lon32, latsc32; std::vector<double2> lon64, lat64; std::vector<float> lat32;
};
static HostScene g_scene;

static const HeightField& host_scene(const void* map, int is_i16, int W, int H, float scale, float rs, float dmax) {
    HostScene& S = g_scene;
    if (S.key != map || S.W != W || S.H != H) {
        S = HostScene();
        S.key = map; S.W = W; S.H = H;
        HeightField& hf = S.hf;
        memset(&hf, 0, sizeof(hf));
        hf.base = map; hf.is_i16 = is_i16; hf.W = W; hf.H = H;
        if (is_i16) build_host_pyramid<int16_t>(hf, S.s16); else build_host_pyramid<float>(hf, S.s32);
        S.lon32.resize(W + 1); S.lon64.resize(W + 1); S.lat64.resize(H); S.lat32.resize(H); S.latsc32.resize(H);
        for (int i = 0; i <= W; ++i) { double sn, cs; sincospi((2.0 * i + 1.0) / W - 1.0, &sn, &cs); S.lon64[i] = make_double2(cs, sn); S.lon32[i] = make_float2((float)cs, (float)sn); }
        for (int i = 0; i < H; ++i) { double sn, cs; sincospi((i + 0.5) / H, &sn, &cs); S.lat64[i] = make_double2(cs, sn); S.lat32[i] = (float)cs; S.latsc32[i] = make_float2((float)cs, (float)sn); }
        hf.lon32 = S.lon32.data(); hf.lon64 = S.lon64.data(); hf.lat32 = S.lat32.data(); hf.lat64 = S.lat64.data(); hf.latsc32 = S.latsc32.data();

    }
    S.hf.scale = scale; S.hf.radius_scale = rs; S.hf.dmax = dmax;
    return S.hf;
}

extern "C" int dbg_trace(const void* map, int is_i16, int W, int H, float scale, float rs, float dmax,
                         const double* rays, int n, double s_min, double radius, int any_hit, int start_level, double* out) {
    const HeightField& hf = host_scene(map, is_i16, W, H, scale, rs, dmax);
    if (start_level < 0) start_level = hf.top + start_level;
#pragma omp parallel for schedule(dynamic, 16)
    for (int i = 0; i < n; ++i) {
        const double* q = rays + (size_t)i * 6;
        Ray64 R;
        R.ox = q[0]; R.oy = q[1]; R.oz = q[2]; R.dx = q[3]; R.dy = q[4]; R.dz = q[5];
        R.oo = R.ox * R.ox + R.oy * R.oy + R.oz * R.oz; R.od = R.ox * R.dx + R.oy * R.dy + R.oz * R.dz;
        TraceOut t; Counters c = {0, 0, 0};
        if (is_i16) trace_ray<true>(hf, radius, R, s_min, any_hit != 0, start_level, t, c);
        else        trace_ray<false>(hf, radius, R, s_min, any_hit != 0, start_level, t, c);
        double* o = out + (size_t)i * 8;
        o[0] = t.hit; o[1] = t.hit ? t.s : -1; o[2] = t.hit && !any_hit ? t.info.r : 0; o[3] = t.hit && !any_hit ? t.info.lon : 0;
        o[4] = t.hit && !any_hit ? t.info.lat : 0; o[5] = c.nodes; o[6] = c.tests; o[7] = c.overflow;
    }
    return 0;
}

// fast (filtered float32) path: out[i] = {status (0 miss, 1 hit, 2 defer), s, fc, fr, r0, c0, nodes, tests}
extern "C" int dbg_trace_fast(const void* map, int is_i16, int W, int H, float scale, float rs, float dmax,
                              const double* rays, int n, double s_min, double radius, int start_level, double* out) {
    const HeightField& hf = host_scene(map, is_i16, W, H, scale, rs, dmax);
    if (start_level < 0) start_level = hf.top + start_level;
    const FastConsts K = make_fast_consts(hf, radius);
#pragma omp parallel for schedule(dynamic, 16)
    for (int i = 0; i < n; ++i) {
        const double* q = rays + (size_t)i * 6;
        Ray64 R;
        R.ox = q[0]; R.oy = q[1]; R.oz = q[2]; R.dx = q[3]; R.dy = q[4]; R.dz = q[5];
        R.oo = R.ox * R.ox + R.oy * R.oy + R.oz * R.oz; R.od = R.ox * R.dx + R.oy * R.dy + R.oz * R.dz;
        FastHit fh; Counters c = {0, 0, 0};
        memset(&fh, 0, sizeof(fh));
        int st = is_i16 ? trace_ray_fast<true>(hf, K, radius, R, s_min, start_level, fh, c)
                        : trace_ray_fast<false>(hf, K, radius, R, s_min, start_level, fh, c);
        double* o = out + (size_t)i * 8;
        o[0] = st; o[1] = st == FT_HIT ? fh.s : -1; o[2] = fh.fc; o[3] = fh.fr; o[4] = fh.r0; o[5] = fh.c0; o[6] = c.nodes; o[7] = c.tests;
    }
    return 0;
}

// beam pre-pass of n pixels: rays = centre rays, delta = angular radius of the pixel; out[i] = {alive, s_start, level, nodes}
extern "C" int dbg_beam(const void* map, int is_i16, int W, int H, float scale, float rs, float dmax, float dmin,
                        const double* rays, int n, double delta, double radius, double* out) {
    HeightField hf = host_scene(map, is_i16, W, H, scale, rs, dmax);
    hf.dmin = dmin;
    const FastConsts K = make_fast_consts(hf, radius);
#pragma omp parallel for schedule(dynamic, 16)
    for (int i = 0; i < n; ++i) {
        const double* q = rays + (size_t)i * 6;
        Ray64 R;
        R.ox = q[0]; R.oy = q[1]; R.oz = q[2]; R.dx = q[3]; R.dy = q[4]; R.dz = q[5];
        R.oo = R.ox * R.ox + R.oy * R.oy + R.oz * R.oz; R.od = R.ox * R.dx + R.oy * R.dy + R.oz * R.dz;
        Counters c = {0, 0, 0};
        double s_start = 0.0; int level = 0;
        const bool alive = is_i16 ? beam_walk<true>(hf, K, radius, R, delta, s_start, level, c)
                                  : beam_walk<false>(hf, K, radius, R, delta, s_start, level, c);
        double* o = out + (size_t)i * 4;
        o[0] = alive; o[1] = s_start; o[2] = level; o[3] = c.nodes;
    }
    return 0;
}

// as dbg_trace_fast, every ray with its own s_min and start level (what the beam pre-pass hands to the samples)
static int g_ceil_level = 0; static float g_dmin = 0.9f;
extern "C" void dbg_set_ceiling(int level, float dmin) { g_ceil_level = level; g_dmin = dmin; }

extern "C" int dbg_trace_fast_from(const void* map, int is_i16, int W, int H, float scale, float rs, float dmax,
                                   const double* rays, int n, const double* s_min, const int* start_level, double radius, double* out) {
    HeightField hf = host_scene(map, is_i16, W, H, scale, rs, dmax);
    hf.dmin = g_dmin;
    const FastConsts K = make_fast_consts(hf, radius);
#pragma omp parallel for schedule(dynamic, 16)
    for (int i = 0; i < n; ++i) {
        const double* q = rays + (size_t)i * 6;
        Ray64 R;
        R.ox = q[0]; R.oy = q[1]; R.oz = q[2]; R.dx = q[3]; R.dy = q[4]; R.dz = q[5];
        R.oo = R.ox * R.ox + R.oy * R.oy + R.oz * R.oz; R.od = R.ox * R.dx + R.oy * R.dy + R.oz * R.dz;
        FastHit fh; Counters c = {0, 0, 0};
        memset(&fh, 0, sizeof(fh));
        int sl = start_level[i] < 0 ? hf.top + start_level[i] : start_level[i];
        int st = is_i16 ? trace_ray_fast<true>(hf, K, radius, R, s_min[i], sl, fh, c, g_ceil_level)
                        : trace_ray_fast<false>(hf, K, radius, R, s_min[i], sl, fh, c, g_ceil_level);
        double* o = out + (size_t)i * 8;
        o[0] = st; o[1] = st == FT_HIT ? fh.s : -1; o[2] = fh.fc; o[3] = fh.fr; o[4] = fh.r0; o[5] = fh.c0; o[6] = c.nodes; o[7] = c.tests;
    }
    return 0;
}

// ---- the GPU generator of csrc/synth.cu, on the host (same recipe; rounding of the library calls may differ in the last bit) ----
namespace hsynth {
static inline uint32_t hash3(int x, int y, int z, uint32_t seed) {
    uint32_t h = seed ^ (uint32_t)x * 0x8da6b343u ^ (uint32_t)y * 0xd8163841u ^ (uint32_t)z * 0xcb1ab31fu;
    h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
    return h;
}
static inline float u01(uint32_t h) { return (float)(h >> 8) * (1.0f / 16777216.0f); }
static float value_noise(float x, float y, float z, uint32_t seed) {
    const float fx = floorf(x), fy = floorf(y), fz = floorf(z);
    const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
    float tx = x - fx, ty = y - fy, tz = z - fz;
    tx = tx * tx * (3.0f - 2.0f * tx); ty = ty * ty * (3.0f - 2.0f * ty); tz = tz * tz * (3.0f - 2.0f * tz);
    float c[2][2][2];
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int d = 0; d < 2; ++d) c[a][b][d] = u01(hash3(ix + a, iy + b, iz + d, seed)) * 2.0f - 1.0f;
    const float x00 = c[0][0][0] + tx * (c[1][0][0] - c[0][0][0]), x10 = c[0][1][0] + tx * (c[1][1][0] - c[0][1][0]);
    const float x01 = c[0][0][1] + tx * (c[1][0][1] - c[0][0][1]), x11 = c[0][1][1] + tx * (c[1][1][1] - c[0][1][1]);
    const float y0 = x00 + ty * (x10 - x00), y1 = x01 + ty * (x11 - x01);
    return y0 + tz * (y1 - y0);
}
static float crater_octave(float x, float y, float z, float cell, uint32_t seed) {
    const float inv = 1.0f / cell;
    const float gx = x * inv - 0.5f, gy = y * inv - 0.5f, gz = z * inv - 0.5f;
    const int ix = (int)floorf(gx), iy = (int)floorf(gy), iz = (int)floorf(gz);
    float sum = 0.0f;
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int d = 0; d < 2; ++d) {
        const int cx = ix + a, cy = iy + b, cz = iz + d;
        const uint32_t h = hash3(cx, cy, cz, seed);
        if ((h & 3u) != 0u) continue;
        float px = (cx + 0.25f + 0.5f * u01(hash3(cx, cy, cz, seed + 1))) * cell;
        float py = (cy + 0.25f + 0.5f * u01(hash3(cx, cy, cz, seed + 2))) * cell;
        float pz = (cz + 0.25f + 0.5f * u01(hash3(cx, cy, cz, seed + 3))) * cell;
        const float n = 1.0f / sqrtf(px * px + py * py + pz * pz);
        if (fabsf(1.0f / n - 1.0f) > 0.5f * cell) continue;
        px *= n; py *= n; pz *= n;
        const float rad = cell * (0.12f + 0.33f * u01(hash3(cx, cy, cz, seed + 4)));
        const float dx = x - px, dy = y - py, dz = z - pz;
        const float t = sqrtf(dx * dx + dy * dy + dz * dz) / rad;
        if (t >= 1.6f) continue;
        const float depth = rad * 0.18f;
        if (t < 1.0f) sum -= depth * (1.0f - t * t);
        else { const float q = (t - 1.0f) * 4.0f; sum += 0.35f * depth * expf(-q * q); }
    }
    return sum;
}
}  // namespace hsynth

extern "C" void dbg_synth_ldem(int16_t* out, int W, int H, uint32_t seed) {
    using namespace hsynth;
    int octaves = 1;
    while ((1.5f * (float)(1 << octaves)) * 2.0f < (float)W / 3.14159f && octaves < 16) ++octaves;
#pragma omp parallel for schedule(dynamic, 8)
    for (int row = 0; row < H; ++row) {
        const float lat = (0.5f - (row + 0.5f) / H) * 3.141592653589793f;
        const float sl = sinf(lat), cl = cosf(lat);
        for (int col = 0; col < W; ++col) {
            const float lon = ((col + 0.5f) / W - 0.5f) * 6.283185307179586f;
            const float x = cl * sinf(lon), y = -cl * cosf(lon), z = sl;
            float f = 1.5f, a = 1.0f, s = 0.0f, norm = 0.0f;
            for (int o = 0; o < octaves; ++o) {
                s += a * value_noise(x * f + 17.0f, y * f - 5.0f, z * f + 3.0f, seed + 101u * o);
                norm += (o < 3) ? a : 0.0f;
                f *= 2.0f; a *= 0.56f;
            }
            s /= norm;
            float cr = 0.0f;
            cr += crater_octave(x, y, z, 0.30f, seed + 7001u);
            cr += crater_octave(x, y, z, 0.09f, seed + 7002u);
            cr += crater_octave(x, y, z, 0.027f, seed + 7003u);
            cr += crater_octave(x, y, z, 0.008f, seed + 7004u);
            float counts = 1700.0f + 11000.0f * s + cr * 3474800.0f;
            counts = fminf(fmaxf(counts, -18200.0f), 21600.0f);
            out[(size_t)row * W + col] = (int16_t)lrintf(counts);
        }
    }
}
