// DEVELOPMENT TOOL (not part of the product, never loaded by moonrtx_b200): runs the
// __host__ __device__ traversal core of csrc/trace_core.cuh on the CPU so that parity
// problems can be investigated against the oracle without a GPU.
#include "../moonrtx_b200/csrc/trace_fast.cuh"
#include <vector>
#include <algorithm>

void mrtx_set_error(const char*, ...) {}

using namespace mrtx_core;

template <typename T>
static void build_host_pyramid(HeightField& hf, std::vector<std::vector<T>>& store) {
    const int W = hf.W, H = hf.H;
    int top = 0;
    while ((W >> (top + 1)) >= 64 && top + 1 < MRTX_MAX_LEVELS) ++top;
    hf.top = top; hf.nx[0] = W; hf.ny[0] = H - 1;
    store.resize(top + 1);
    const T* base = (const T*)hf.base;
    for (int k = 1; k <= top; ++k) {
        hf.nx[k] = (W + (1 << k) - 1) >> k; hf.ny[k] = (H - 1 + (1 << k) - 1) >> k;
        store[k].resize((size_t)hf.nx[k] * hf.ny[k]);
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
                const T* in = store[k - 1].data();
                m = std::max(std::max(in[(size_t)(2 * J) * inx + 2 * I], in[(size_t)(2 * J) * inx + c1]),
                             std::max(in[(size_t)r1 * inx + 2 * I], in[(size_t)r1 * inx + c1]));
            }
            store[k][(size_t)J * hf.nx[k] + I] = m;
        }
        hf.level[k] = store[k].data();
    }
}

extern "C" void dbg_set(int v) { mrtx_core::g_debug = v; }

// pyramid + wall tables are cached across calls (keyed on the map pointer): they take minutes at 92160 x 46080
struct HostScene {
    const void* key = nullptr; int W = 0, H = 0;
    HeightField hf;
    std::vector<std::vector<int16_t>> s16; std::vector<std::vector<float>> s32;
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
