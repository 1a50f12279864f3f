// K1 + K2: LDEM elevation block mean + normalisation.
//
// Replaces moonrtx/data_loader.py:223-227 (numpy reshape + two float32 means),
// :227 (*scale), :232 (+1) and :241-242 (max, /max).  Bit-exact against numpy:
//   stage 1  m[r][j] = fl32( exact_int_sum(src[r][j*ds .. j*ds+ds-1]) / fl32(ds) )
//   stage 2  a       = (((m[0]+m[1])+m[2])+...)   every add rounded to f32, rows in order
//            mean    = fl32(a / fl32(ds))
//   e = fl32(fl32(mean * fl32(scale)) + 1);  radius_scale = max e;  out = fl32(e / radius_scale)
// All float ops go through __f*_rn intrinsics so ptxas can neither contract them into
// FMAs nor replace the divisions by reciprocal multiplies.
//
// Bandwidth-bound (2 B read per texel, 4/ds^2 B written): the source is streamed with
// 256-bit loads marked L1::no_allocate + L2::evict_first (sm_100 LDG.E.NA.EFL2.256), the
// un-normalised result is stored with L2::evict_last so that pass 2 (the divide by the
// global max, which cannot start before every block mean is known) finds it in the
// 126 MB L2 instead of HBM whenever it fits.

#include "common.cuh"

namespace {

__device__ __forceinline__ void ld256_stream(const void* p, uint32_t (&v)[8]) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(p));
}
__device__ __forceinline__ void st256_keep(void* p, const float (&f)[8]) {
    asm volatile("st.global.L2::evict_last.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "r"(__float_as_uint(f[0])), "r"(__float_as_uint(f[1])), "r"(__float_as_uint(f[2])),
                    "r"(__float_as_uint(f[3])), "r"(__float_as_uint(f[4])), "r"(__float_as_uint(f[5])),
                    "r"(__float_as_uint(f[6])), "r"(__float_as_uint(f[7]))
                 : "memory");
}
__device__ __forceinline__ void ld256_last_use(const void* p, float (&f)[8]) {
    uint32_t v[8];
    asm volatile("ld.global.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(p) : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[i]);
}
__device__ __forceinline__ void st256_stream(void* p, const float (&f)[8]) {
    asm volatile("st.global.L2::evict_first.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "r"(__float_as_uint(f[0])), "r"(__float_as_uint(f[1])), "r"(__float_as_uint(f[2])),
                    "r"(__float_as_uint(f[3])), "r"(__float_as_uint(f[4])), "r"(__float_as_uint(f[5])),
                    "r"(__float_as_uint(f[6])), "r"(__float_as_uint(f[7]))
                 : "memory");
}

__host__ __device__ constexpr int gcd_c(int a, int b) { return b == 0 ? a : gcd_c(b, a % b); }
__host__ __device__ constexpr int lcm_c(int a, int b) { return a / gcd_c(a, b) * b; }
// outputs per thread: whole 32-byte loads per source row AND whole 32-byte stores
__host__ __device__ constexpr int group_of(int ds) { return lcm_c(16 / gcd_c(ds, 16), 8); }

__device__ __forceinline__ float finish(float acc, float dsf, float scale) {
    float mean = __fdiv_rn(acc, dsf);
    return __fadd_rn(__fmul_rn(mean, scale), 1.0f);
}

// Publish a block's maximum: one warp shuffle tree, one smem hop, and an atomic only
// when the block can actually raise the global value (e > 0 always, so the uint
// ordering of the bit patterns is the float ordering).
__device__ __forceinline__ void publish_max(float v, unsigned* gmax_bits) {
    unsigned b = __float_as_uint(v);
    b = __reduce_max_sync(0xffffffffu, b);
    __shared__ unsigned wmax[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) wmax[warp] = b;
    __syncthreads();
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        b = lane < nw ? wmax[lane] : 0u;
        b = __reduce_max_sync(0xffffffffu, b);
        if (lane == 0 && b > *(volatile unsigned*)gmax_bits) atomicMax(gmax_bits, b);
    }
}

// Vector path: W % 16 == 0, 32-byte aligned base, (W/DS) % G == 0.
template <int DS>
__global__ void __launch_bounds__(256)
downscale_vec_kernel(const int16_t* __restrict__ src, float* __restrict__ out, int W, int h, int w,
                     float scale, unsigned* __restrict__ gmax_bits) {
    constexpr int G = group_of(DS);             // outputs per thread
    constexpr int NV = G * DS / 16;             // 256-bit loads per source row
    constexpr int RC = (8 / NV) > 0 ? ((8 / NV) < DS ? (8 / NV) : DS) : 1;   // rows in flight
    const int groups = w / G;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float emax = 0.0f;
    if (idx < (long long)groups * h) {
        const int orow = (int)(idx / groups);
        const int g = (int)(idx - (long long)orow * groups);
        const int16_t* p = src + (size_t)orow * DS * W + (size_t)g * G * DS;
        const float dsf = (float)DS;
        float acc[G];
#pragma unroll
        for (int r0 = 0; r0 < DS; r0 += RC) {
            uint32_t v[RC][NV][8];
#pragma unroll
            for (int r = 0; r < RC; ++r)
                if (r0 + r < DS) {
#pragma unroll
                    for (int q = 0; q < NV; ++q) ld256_stream(p + (size_t)(r0 + r) * W + q * 16, v[r][q]);
                }
#pragma unroll
            for (int r = 0; r < RC; ++r)
                if (r0 + r < DS) {
#pragma unroll
                    for (int o = 0; o < G; ++o) {
                        int s = 0;
#pragma unroll
                        for (int c = 0; c < DS; ++c) {
                            const int e = o * DS + c;               // element within the row chunk
                            const uint32_t word = v[r][e / 16][(e % 16) / 2];
                            s += (e & 1) ? ((int)word >> 16) : (int)(short)(word & 0xffffu);
                        }
                        const float m = __fdiv_rn((float)s, dsf);
                        acc[o] = (r0 + r == 0) ? m : __fadd_rn(acc[o], m);
                    }
                }
        }
        float* q = out + (size_t)orow * w + (size_t)g * G;
#pragma unroll
        for (int o8 = 0; o8 < G; o8 += 8) {
            float e[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                e[i] = finish(acc[o8 + i], dsf, scale);
                emax = fmaxf(emax, e[i]);
            }
            st256_keep(q + o8, e);
        }
    }
    publish_max(emax, gmax_bits);
}

// Any ds, any alignment, output columns [c0, w): one thread per output texel.
__global__ void __launch_bounds__(256)
downscale_generic_kernel(const int16_t* __restrict__ src, float* __restrict__ out, int W, int h, int w,
                         int ds, int c0, float scale, unsigned* __restrict__ gmax_bits) {
    const int wc = w - c0;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float emax = 0.0f;
    if (idx < (long long)wc * h) {
        const int orow = (int)(idx / wc);
        const int ocol = c0 + (int)(idx - (long long)orow * wc);
        const int16_t* p = src + (size_t)orow * ds * W + (size_t)ocol * ds;
        const float dsf = (float)ds;
        float acc = 0.0f;
        for (int r = 0; r < ds; ++r) {
            int s = 0;
            for (int c = 0; c < ds; ++c) s += (int)__ldg(p + (size_t)r * W + c);
            // |s| < 2^24 for ds <= 512, so (float)s is exact like numpy's f32 row sum
            const float m = __fdiv_rn((float)s, dsf);
            acc = r == 0 ? m : __fadd_rn(acc, m);
        }
        const float e = finish(acc, dsf, scale);
        out[(size_t)orow * w + ocol] = e;
        emax = e;
    }
    publish_max(emax, gmax_bits);
}

// K2: out /= max.  256-bit vectors for the aligned body, scalars for the tail.
__global__ void __launch_bounds__(256)
normalise_kernel(float* __restrict__ out, size_t n, size_t nvec, const unsigned* __restrict__ gmax_bits,
                 float* __restrict__ host_rs) {
    const float mx = __uint_as_float(*gmax_bits);
    // radius_scale goes to the caller through a mapped pinned word: no copy engine round trip behind the kernel
    if (host_rs && blockIdx.x == 0 && threadIdx.x == 0) { *host_rs = mx; __threadfence_system(); }
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        float f[8];
        ld256_last_use(out + i * 8, f);
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = __fdiv_rn(f[k], mx);
        st256_stream(out + i * 8, f);
    }
    for (size_t i = nvec * 8 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = __fdiv_rn(out[i], mx);
}

template <int DS>
void launch_vec(const int16_t* src, float* out, int W, int h, int w, float scale, unsigned* gmax,
                cudaStream_t st) {
    const long long threads = (long long)(w / group_of(DS)) * h;
    const unsigned blocks = (unsigned)((threads + 255) / 256);
    downscale_vec_kernel<DS><<<blocks, 256, 0, st>>>(src, out, W, h, w, scale, gmax);
}

}  // namespace

// The three steps of a downscale, so that a map arriving from the host band by band can be reduced while it arrives:
// begin (clear the running max), band (block means of source rows [0, Hb) of `src` -> rows of `out`), finish (/ max).
int downscale_begin(mrtx_ctx* ctx) {
    MRTX_CUDA(cudaMemsetAsync(ctx->d_max_bits, 0, sizeof(unsigned), ctx->stream));
    return MRTX_OK;
}

int downscale_band(mrtx_ctx* ctx, const int16_t* src, int W, int Hb, int ds, float* out) {
    const int h = Hb / ds, w = W / ds;
    // data_loader.py:216 - python float scale, applied to a float32 array => fl32(scale)
    const float scale = (float)(0.5 / 1737400.0);
    cudaStream_t st = ctx->stream;
    int c0 = 0;   // first output column left to the generic kernel
    const bool aligned = (W % 16 == 0) && ((uintptr_t)src % 32 == 0) && ((uintptr_t)out % 32 == 0) && (w % 8 == 0);
    if (aligned) {
#define MRTX_DS_CASE(D)                                                               \
    case D:                                                                           \
        if (w % group_of(D) == 0) {                                                   \
            launch_vec<D>(src, out, W, h, w, scale, ctx->d_max_bits, st);             \
            c0 = w;                                                                   \
        }                                                                             \
        break;
        switch (ds) {
            MRTX_DS_CASE(1) MRTX_DS_CASE(2) MRTX_DS_CASE(3) MRTX_DS_CASE(4) MRTX_DS_CASE(5)
            MRTX_DS_CASE(6) MRTX_DS_CASE(8) MRTX_DS_CASE(12) MRTX_DS_CASE(16)
            default: break;
        }
#undef MRTX_DS_CASE
    }
    if (c0 < w) {
        const long long threads = (long long)(w - c0) * h;
        const unsigned blocks = (unsigned)((threads + 255) / 256);
        downscale_generic_kernel<<<blocks, 256, 0, st>>>(src, out, W, h, w, ds, c0, scale, ctx->d_max_bits);
    }
    MRTX_CUDA(cudaGetLastError());
    return MRTX_OK;
}

int downscale_finish(mrtx_ctx* ctx, float* out, size_t n, float* host_rs_dev) {
    // 256-bit vectors need a 32-byte aligned destination; otherwise everything is "tail"
    const size_t nvec = ((uintptr_t)out % 32 == 0) ? n / 8 : 0;
    size_t want = ((nvec ? nvec : n) + 255) / 256;
    const size_t cap = (size_t)ctx->sm_count * 16;
    const unsigned blocks = (unsigned)(want < 1 ? 1 : (want > cap ? cap : want));
    normalise_kernel<<<blocks, 256, 0, ctx->stream>>>(out, n, nvec, ctx->d_max_bits, host_rs_dev);
    MRTX_CUDA(cudaGetLastError());
    return MRTX_OK;
}

int launch_downscale_i16(mrtx_ctx* ctx, const int16_t* src, int W, int H, int ds, float* out, float* host_rs_dev) {
    int rc = downscale_begin(ctx);
    if (!rc) rc = downscale_band(ctx, src, W, H, ds, out);
    if (!rc) rc = downscale_finish(ctx, out, (size_t)(H / ds) * (W / ds), host_rs_dev);
    return rc;
}
