// K4: max-height pyramid of the displacement map (enables K5/K6; no reference counterpart -
// PlotOptiX step-marches the texture, moon_renderer.py:85-101).
//
// The surface is r = R * bilinear(D) between texel centres (renderer_navigation.py:558-599):
// columns wrap, rows clamp.  A level-0 cell (r0, c0) is one bilinear patch, bounded above by
// its four corner texels.  Level k >= 1 cell (J, I) covers level-0 cells
// rows [J*2^k, (J+1)*2^k) x cols [I*2^k, (I+1)*2^k) (clipped to H-1 rows / W cols) and stores
// the max texel over their corner footprint, (2^k+1)^2 texels - conservative under bilinear
// interpolation, which a plain 2x2 max-pool of texels would not be.
// Levels keep the map's dtype (int16 counts are monotone in D), row-major.

#include "common.cuh"

namespace {

template <typename T>
__global__ void level1_kernel(const T* __restrict__ base, int W, int H, T* __restrict__ out, int nx, int ny) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)nx * ny) return;
    const int J = (int)(idx / nx), I = (int)(idx - (long long)J * nx);
    const int r_lo = 2 * J, r_hi = min(2 * J + 2, H - 1);          // texel rows r_lo..r_hi inclusive
    const int c_lo = 2 * I, c_hi = min(2 * I + 2, W);              // texel cols, c == W wraps to 0
    T m = base[(size_t)r_lo * W + c_lo];
    for (int r = r_lo; r <= r_hi; ++r)
        for (int c = c_lo; c <= c_hi; ++c) {
            const T v = base[(size_t)r * W + (c >= W ? c - W : c)];
            m = v > m ? v : m;
        }
    out[idx] = m;
}

template <typename T>
__global__ void levelk_kernel(const T* __restrict__ in, int inx, int iny, T* __restrict__ out, int nx, int ny) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)nx * ny) return;
    const int J = (int)(idx / nx), I = (int)(idx - (long long)J * nx);
    const int r1 = min(2 * J + 1, iny - 1), c1 = min(2 * I + 1, inx - 1);
    T m = in[(size_t)(2 * J) * inx + 2 * I];
    T v = in[(size_t)(2 * J) * inx + c1];  m = v > m ? v : m;
    v = in[(size_t)r1 * inx + 2 * I];      m = v > m ? v : m;
    v = in[(size_t)r1 * inx + c1];         m = v > m ? v : m;
    out[idx] = m;
}

// dil[k](J, I) = max of level k over the cell and its neighbours (reach cells either side in longitude, wrapping; one
// row either side, clamped): an upper bound of the surface for every direction within one level-k cell of the cell.
// The beam pre-pass walks these levels with the CENTRE ray of a pixel (trace_fast.cuh, BeamCtl).
template <typename T>
__global__ void dilate_kernel(const T* __restrict__ in, int nx, int ny, int reach, T* __restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)nx * ny) return;
    const int J = (int)(idx / nx), I = (int)(idx - (long long)J * nx);
    T m = in[idx];
    for (int j = max(J - 1, 0); j <= min(J + 1, ny - 1); ++j)
        for (int d = -reach; d <= reach; ++d) {
            int i = I + d;
            i = i < 0 ? i + nx : (i >= nx ? i - nx : i);
            const T v = in[(size_t)j * nx + i];
            m = v > m ? v : m;
        }
    out[idx] = m;
}

// global min / max of the base map, as order-preserving ints
template <typename T> __device__ __forceinline__ int ordered(T v);
template <> __device__ __forceinline__ int ordered<int16_t>(int16_t v) { return (int)v; }
template <> __device__ __forceinline__ int ordered<float>(float v) {
    const int b = __float_as_int(v);
    return b >= 0 ? b : b ^ 0x7fffffff;
}

template <typename T>
__global__ void minmax_kernel(const T* __restrict__ base, size_t n, int* __restrict__ mm) {
    int lo = 0x7fffffff, hi = (int)0x80000000;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int o = ordered<T>(base[i]);
        lo = min(lo, o); hi = max(hi, o);
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((threadIdx.x & 31) == 0) { atomicMin(&mm[0], lo); atomicMax(&mm[1], hi); }
}

__global__ void wall_tables_kernel(float2* lon32, double2* lon64, float* lat32, float2* latsc32, double2* lat64, int W, int H) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= W) {
        double sn, cs;
        sincospi((2.0 * i + 1.0) / W - 1.0, &sn, &cs);
        lon64[i] = make_double2(cs, sn);
        lon32[i] = make_float2((float)cs, (float)sn);
    }
    if (i < H) {
        double sn, cs;                              // phi_j = pi/2 - pi (j+0.5)/H
        sincospi((i + 0.5) / H, &sn, &cs);
        lat64[i] = make_double2(cs, sn);            // (sin phi, cos phi)
        lat32[i] = (float)cs;
        latsc32[i] = make_float2((float)cs, (float)sn);
    }
}

int build_wall_tables(mrtx_ctx* ctx) {
    HeightField& hf = ctx->hf;
    const int W = hf.W, H = hf.H;
    const size_t n_lon = (size_t)W + 1;
    const size_t off_lon64 = 0, off_lat64 = off_lon64 + n_lon * sizeof(double2);
    const size_t off_lon32 = off_lat64 + (size_t)H * sizeof(double2), off_lat32 = off_lon32 + n_lon * sizeof(float2);
    const size_t off_latsc32 = (off_lat32 + (size_t)H * sizeof(float) + 15) & ~(size_t)15;
    const size_t total = off_latsc32 + (size_t)H * sizeof(float2);
    MRTX_CUDA(cudaMalloc(&ctx->hf_tables_owned, total));
    char* p = (char*)ctx->hf_tables_owned;
    hf.lon64 = (const double2*)(p + off_lon64); hf.lat64 = (const double2*)(p + off_lat64);
    hf.lon32 = (const float2*)(p + off_lon32);  hf.lat32 = (const float*)(p + off_lat32);
    hf.latsc32 = (const float2*)(p + off_latsc32);
    const int n = (W + 1 > H ? W + 1 : H);
    wall_tables_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>((float2*)hf.lon32, (double2*)hf.lon64, (float*)hf.lat32,
                                                                 (float2*)hf.latsc32, (double2*)hf.lat64, W, H);
    MRTX_CUDA(cudaGetLastError());
    return MRTX_OK;
}

// level (row-major) -> 8 x 8-cell tiles, tile after tile in row-major order, cells past the level's edge repeated from it
template <typename T>
__global__ void tile_kernel(const T* __restrict__ src, int nx, int ny, int tx, long long n, T* __restrict__ dst) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long tile = i >> 6;
    const int within = (int)(i & 63);
    const int J = (int)(tile / tx) * 8 + (within >> 3), I = (int)(tile % tx) * 8 + (within & 7);
    dst[i] = src[(size_t)min(J, ny - 1) * nx + min(I, nx - 1)];
}

template <typename T>
int build_levels(mrtx_ctx* ctx) {
    HeightField& hf = ctx->hf;
    const int W = hf.W, H = hf.H;
    int top = 0;
    while ((W >> (top + 1)) >= 64 && top + 1 < MRTX_MAX_LEVELS) ++top;
    hf.top = top;
    hf.nx[0] = W; hf.ny[0] = H - 1;
    size_t total = 0;
    for (int k = 1; k <= top; ++k) {
        hf.nx[k] = (W + (1 << k) - 1) >> k;
        hf.ny[k] = (H - 1 + (1 << k) - 1) >> k;
        total += (((size_t)hf.nx[k] * hf.ny[k] * sizeof(T)) + 255) & ~(size_t)255;
        if (k >= MRTX_DIL_MIN_LEVEL) total += (((size_t)hf.nx[k] * hf.ny[k] * sizeof(T)) + 255) & ~(size_t)255;
#if MRTX_TILED
        total += (size_t)((hf.nx[k] + 7) >> 3) * ((hf.ny[k] + 7) >> 3) * 64 * sizeof(T);        // the tiled copy (a multiple of 128 bytes)
#endif
    }
    cudaStream_t st = ctx->stream;
    if (top > 0) {
        MRTX_CUDA(cudaMalloc(&ctx->hf_levels_owned, total));
        char* p = (char*)ctx->hf_levels_owned;
        for (int k = 1; k <= top; ++k) {
            hf.level[k] = p;
            p += (((size_t)hf.nx[k] * hf.ny[k] * sizeof(T)) + 255) & ~(size_t)255;
        }
        hf.lvl_base = ctx->hf_levels_owned;
        const T* base = (const T*)hf.base;
        {
            const long long n = (long long)hf.nx[1] * hf.ny[1];
            level1_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(base, W, H, (T*)hf.level[1], hf.nx[1], hf.ny[1]);
        }
        for (int k = 2; k <= top; ++k) {
            const long long n = (long long)hf.nx[k] * hf.ny[k];
            levelk_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const T*)hf.level[k - 1], hf.nx[k - 1], hf.ny[k - 1],
                                                                          (T*)hf.level[k], hf.nx[k], hf.ny[k]);
        }
        // dilated copies of levels >= MRTX_DIL_MIN_LEVEL (a level whose last column is a partial cell dilates by two)
        for (int k = 1; k <= top; ++k) hf.dil[k] = nullptr;
        for (int k = MRTX_DIL_MIN_LEVEL; k <= top; ++k) {
            hf.dil[k] = p;
            p += (((size_t)hf.nx[k] * hf.ny[k] * sizeof(T)) + 255) & ~(size_t)255;
            const long long n = (long long)hf.nx[k] * hf.ny[k];
            const int reach = ((long long)hf.nx[k] << k) != W ? 2 : 1;
            dilate_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const T*)hf.level[k], hf.nx[k], hf.ny[k], reach < hf.nx[k] / 2 ? reach : hf.nx[k] / 2, (T*)hf.dil[k]);
        }
#if MRTX_TILED
        // the levels again in 8 x 8-cell tiles for the filtered walk: a ray moves through a level in both directions and
        // descends to the four children of a cell - in rows, every step in latitude and every second child is another
        // cache line; in tiles, seven of eight steps and all four children stay in the line
        for (int k = 1; k <= top; ++k) {
            const int tx = (hf.nx[k] + 7) >> 3, ty = (hf.ny[k] + 7) >> 3;
            hf.off[2 * MRTX_MAX_LEVELS + k] = (unsigned)((p - (const char*)hf.lvl_base) / sizeof(T));
            const long long n = (long long)tx * ty * 64;
            tile_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const T*)hf.level[k], hf.nx[k], hf.ny[k], tx, n, (T*)p);
            p += (size_t)n * sizeof(T);
        }
#endif
        MRTX_CUDA(cudaGetLastError());
        for (int k = 1; k <= top; ++k) {
            hf.off[k] = (unsigned)(((const char*)hf.level[k] - (const char*)hf.lvl_base) / sizeof(T));
            hf.off[MRTX_MAX_LEVELS + k] = hf.dil[k] ? (unsigned)(((const char*)hf.dil[k] - (const char*)hf.lvl_base) / sizeof(T)) : 0u;
        }
    }
    // global range -> bounding sphere (and the "surely inside" sphere)
    int* d_mm = nullptr;
    MRTX_CUDA(cudaMalloc(&d_mm, 2 * sizeof(int)));
    const int init[2] = {0x7fffffff, (int)0x80000000};
    MRTX_CUDA(cudaMemcpyAsync(d_mm, init, sizeof(init), cudaMemcpyHostToDevice, st));
    minmax_kernel<T><<<ctx->sm_count * 8, 256, 0, st>>>((const T*)hf.base, (size_t)W * H, d_mm);
    int mm[2];
    cudaError_t e = cudaMemcpyAsync(mm, d_mm, sizeof(mm), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_mm);
    if (e != cudaSuccess) { mrtx_set_error("pyramid: %s", cudaGetErrorString(e)); return MRTX_ERR_CUDA; }
    if (sizeof(T) == 2) {
        // same three float32 roundings as data_loader.py:219-242
        volatile float lo = (float)mm[0], hi = (float)mm[1];
        lo = lo * hf.scale; lo = lo + 1.0f; lo = lo / hf.radius_scale;
        hi = hi * hf.scale; hi = hi + 1.0f; hi = hi / hf.radius_scale;
        hf.dmin = lo; hf.dmax = hi;
    } else {
        auto unorder = [](int o) { const int b = o >= 0 ? o : o ^ 0x7fffffff; float f; memcpy(&f, &b, 4); return f; };
        hf.dmin = unorder(mm[0]); hf.dmax = unorder(mm[1]);
    }
    if (!(hf.dmin > 0.0f) || !isfinite(hf.dmax)) {
        mrtx_set_error("displacement factors must be finite and positive (got %g .. %g)", hf.dmin, hf.dmax);
        return MRTX_ERR_INVALID;
    }
    return MRTX_OK;
}

}  // namespace

int build_pyramid(mrtx_ctx* ctx) {
    const int rc = build_wall_tables(ctx);
    if (rc) return rc;
    return ctx->hf.is_i16 ? build_levels<int16_t>(ctx) : build_levels<float>(ctx);
}
