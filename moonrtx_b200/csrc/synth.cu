// Synthetic LOLA-shaped inputs generated directly in HBM.
//
// The real LDEM (92160x46080 int16, 8.5 GB) and colour TIFF cannot be downloaded
// offline; bench.py and the large-size tests render these instead.  Relief = fBm of 3-D
// value noise evaluated at the unit-sphere point of every texel (so it is seamless at
// the +/-180 deg meridian and regular at the poles, spectrum slope ~ -2) plus four
// octaves of hashed crater bowls with raised rims, mapped to the real LDEM count range
// (-18200 .. +21600 counts = -9.1 .. +10.8 km at 0.5 m/count, data_loader.py:160-163).

#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t hash3(int x, int y, int z, uint32_t seed) {
    uint32_t h = seed ^ (uint32_t)x * 0x8da6b343u ^ (uint32_t)y * 0xd8163841u ^ (uint32_t)z * 0xcb1ab31fu;
    h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
    return h;
}
__device__ __forceinline__ float u01(uint32_t h) { return (float)(h >> 8) * (1.0f / 16777216.0f); }

__device__ float value_noise(float x, float y, float z, uint32_t seed) {
    const float fx = floorf(x), fy = floorf(y), fz = floorf(z);
    const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
    float tx = x - fx, ty = y - fy, tz = z - fz;
    tx = tx * tx * (3.0f - 2.0f * tx); ty = ty * ty * (3.0f - 2.0f * ty); tz = tz * tz * (3.0f - 2.0f * tz);
    float c[2][2][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int d = 0; d < 2; ++d) c[a][b][d] = u01(hash3(ix + a, iy + b, iz + d, seed)) * 2.0f - 1.0f;
    const float x00 = c[0][0][0] + tx * (c[1][0][0] - c[0][0][0]);
    const float x10 = c[0][1][0] + tx * (c[1][1][0] - c[0][1][0]);
    const float x01 = c[0][0][1] + tx * (c[1][0][1] - c[0][0][1]);
    const float x11 = c[0][1][1] + tx * (c[1][1][1] - c[0][1][1]);
    const float y0 = x00 + ty * (x10 - x00), y1 = x01 + ty * (x11 - x01);
    return y0 + tz * (y1 - y0);
}

// One octave of craters: a 3-D lattice of pitch `cell`; every lattice cell may hold one
// crater whose centre is the cell's hashed point pushed onto the unit sphere.
__device__ float crater_octave(float x, float y, float z, float cell, uint32_t seed) {
    const float inv = 1.0f / cell;
    const float gx = x * inv - 0.5f, gy = y * inv - 0.5f, gz = z * inv - 0.5f;
    const int ix = (int)floorf(gx), iy = (int)floorf(gy), iz = (int)floorf(gz);
    float sum = 0.0f;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                const int cx = ix + a, cy = iy + b, cz = iz + d;
                const uint32_t h = hash3(cx, cy, cz, seed);
                if ((h & 3u) != 0u) continue;                       // a quarter of the cells are cratered
                float px = (cx + 0.25f + 0.5f * u01(hash3(cx, cy, cz, seed + 1))) * cell;
                float py = (cy + 0.25f + 0.5f * u01(hash3(cx, cy, cz, seed + 2))) * cell;
                float pz = (cz + 0.25f + 0.5f * u01(hash3(cx, cy, cz, seed + 3))) * cell;
                const float n = rsqrtf(px * px + py * py + pz * pz);
                if (fabsf(1.0f / n - 1.0f) > 0.5f * cell) continue;  // lattice point too far off the sphere
                px *= n; py *= n; pz *= n;
                const float rad = cell * (0.12f + 0.33f * u01(hash3(cx, cy, cz, seed + 4)));
                const float dx = x - px, dy = y - py, dz = z - pz;
                const float t = sqrtf(dx * dx + dy * dy + dz * dz) / rad;
                if (t >= 1.6f) continue;
                const float depth = rad * 0.18f;                    // depth ~ 0.18 radius (fresh simple crater)
                if (t < 1.0f) sum -= depth * (1.0f - t * t);
                else { const float q = (t - 1.0f) * 4.0f; sum += 0.35f * depth * __expf(-q * q); }
            }
    return sum;
}

__device__ __forceinline__ void sphere_point(int col, int row, int W, int H, float& x, float& y, float& z) {
    const float lon = ((col + 0.5f) / W - 0.5f) * 6.283185307179586f;
    const float lat = (0.5f - (row + 0.5f) / H) * 3.141592653589793f;
    float sl, cl, so, co;
    sincosf(lat, &sl, &cl); sincosf(lon, &so, &co);
    x = cl * so; y = -cl * co; z = sl;           // renderer_navigation.py:47-53
}

__global__ void synth_ldem_kernel(int16_t* __restrict__ out, int W, int H, uint32_t seed, int octaves) {
    const size_t n = (size_t)W * H;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int row = (int)(i / W), col = (int)(i - (size_t)row * W);
        float x, y, z;
        sphere_point(col, row, W, H, x, y, z);
        float f = 1.5f, a = 1.0f, s = 0.0f, norm = 0.0f;
        for (int o = 0; o < octaves; ++o) {
            s += a * value_noise(x * f + 17.0f, y * f - 5.0f, z * f + 3.0f, seed + 101u * o);
            norm += (o < 3) ? a : 0.0f;
            f *= 2.0f; a *= 0.56f;                // amplitude ~ f^-0.84
        }
        s /= norm;                                // roughly [-1, 1]
        // crater relief in units of sphere radius -> counts (1 radius = 3 474 800 counts)
        float cr = 0.0f;
        cr += crater_octave(x, y, z, 0.30f, seed + 7001u);
        cr += crater_octave(x, y, z, 0.09f, seed + 7002u);
        cr += crater_octave(x, y, z, 0.027f, seed + 7003u);
        cr += crater_octave(x, y, z, 0.008f, seed + 7004u);
        float counts = 1700.0f + 11000.0f * s + cr * 3474800.0f;
        counts = fminf(fmaxf(counts, -18200.0f), 21600.0f);
        out[i] = (int16_t)__float2int_rn(counts);
    }
}

__global__ void synth_color_kernel(uint8_t* __restrict__ out, int W, int H, uint32_t seed) {
    const size_t n = (size_t)W * H;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int row = (int)(i / W), col = (int)(i - (size_t)row * W);
        float x, y, z;
        sphere_point(col, row, W, H, x, y, z);
        float f = 2.0f, a = 1.0f, s = 0.0f;
        for (int o = 0; o < 6; ++o) {
            s += a * value_noise(x * f - 9.0f, y * f + 2.0f, z * f + 31.0f, seed + 13u * o);
            f *= 2.3f; a *= 0.6f;
        }
        const float base = 128.0f + 70.0f * s;
        const uint32_t h = hash3(col, row, 0, seed + 99u);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v = base * (0.94f + 0.03f * c) + ((float)((h >> (8 * c)) & 255u) - 127.5f) * 0.12f;
            out[i * 3 + c] = (uint8_t)__float2int_rn(fminf(fmaxf(v, 0.0f), 255.0f));
        }
    }
}

}  // namespace

int launch_synth_ldem(mrtx_ctx* ctx, int16_t* out, int W, int H, uint32_t seed) {
    // octaves until the wavelength reaches ~2 texels
    int oct = 1;
    while ((1.5f * (float)(1 << oct)) * 2.0f < (float)W / 3.14159f && oct < 16) ++oct;
    synth_ldem_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(out, W, H, seed, oct);
    MRTX_CUDA(cudaGetLastError());
    return MRTX_OK;
}

int launch_synth_color(mrtx_ctx* ctx, uint8_t* out, int W, int H, uint32_t seed) {
    synth_color_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(out, W, H, seed);
    MRTX_CUDA(cudaGetLastError());
    return MRTX_OK;
}
