// K3: colour map reduce + albedo LUT + BGR->RGBA.
//
// Replaces moonrtx/data_loader.py:331 (cv2.imread IMREAD_REDUCED_COLOR_k: for a TIFF,
// OpenCV 4.13 decodes in full and resizes with INTER_LINEAR_EXACT, which for sizes
// divisible by k is the round-half-up mean (a+b+c+d+2)>>2 of the CENTRAL 2x2 texels of
// each k x k block) and :345-368 (_moon_texture: lut[] per channel, BGR -> RGBA, A=255).
// Integer arithmetic throughout, so the result is bit-exact.
//
// Only the two central source rows of every k are touched; a warp covers 32*4 = 128
// consecutive output texels, so the byte gathers of one warp fall in a contiguous
// 128*k*3-byte span of each row and are served out of L1 after the first touch.

#include "common.cuh"

namespace {

template <int K>
__global__ void __launch_bounds__(256)
color_reduce_kernel(const uint8_t* __restrict__ bgr, int W, int h, int w,
                    const uint8_t* __restrict__ lut_g, uchar4* __restrict__ out) {
    __shared__ uint8_t lut[256];
    lut[threadIdx.x & 255] = lut_g[threadIdx.x & 255];
    __syncthreads();
    const int quads = (w + 3) / 4;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)quads * h) return;
    const int orow = (int)(idx / quads);
    const int oc0 = (int)(idx - (long long)orow * quads) * 4;
    constexpr int O = K / 2 - 1;                      // first central row / column (K >= 2)
    uchar4 px[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int oc = oc0 + j;
        if (oc >= w) break;
        unsigned b, g, r;
        if (K == 1) {
            const uint8_t* p = bgr + ((size_t)orow * W + oc) * 3;
            b = __ldg(p); g = __ldg(p + 1); r = __ldg(p + 2);
        } else {
            const uint8_t* p0 = bgr + ((size_t)(orow * K + O) * W + (size_t)oc * K + O) * 3;
            const uint8_t* p1 = p0 + (size_t)W * 3;
            b = (__ldg(p0) + __ldg(p0 + 3) + __ldg(p1) + __ldg(p1 + 3) + 2u) >> 2;
            g = (__ldg(p0 + 1) + __ldg(p0 + 4) + __ldg(p1 + 1) + __ldg(p1 + 4) + 2u) >> 2;
            r = (__ldg(p0 + 2) + __ldg(p0 + 5) + __ldg(p1 + 2) + __ldg(p1 + 5) + 2u) >> 2;
        }
        px[j] = make_uchar4(lut[r], lut[g], lut[b], 255);
    }
    uchar4* q = out + (size_t)orow * w + oc0;
    if (oc0 + 3 < w && (((uintptr_t)q) & 15) == 0) {
        *reinterpret_cast<uint4*>(q) = *reinterpret_cast<uint4*>(px);
    } else {
        for (int j = 0; j < 4 && oc0 + j < w; ++j) q[j] = px[j];
    }
}

// rt.set_background(star_map, gamma=g, rt_format="UByte4"), moon_renderer.py:606-607: float RGB in [0, 1] -> 8-bit texture
// of LINEAR radiance v^gamma (the Gamma post-process raises to 1/gamma again, as for the albedo texture), alpha 255.
__global__ void background_texture_kernel(const float* __restrict__ rgb, size_t n, double gamma, uchar4* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned c[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            double v = (double)rgb[i * 3 + q];
            v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
            c[q] = (unsigned)rint(255.0 * pow(v, gamma));
        }
        out[i] = make_uchar4((unsigned char)c[0], (unsigned char)c[1], (unsigned char)c[2], 255);
    }
}

// cv2.resize(..., interpolation=cv2.INTER_CUBIC) of a float32 image (load_starmap, data_loader.py:412-414): Keys cubic
// with A = -0.75, source position (x + 0.5) * W / w - 0.5, taps clamped to the image, horizontal pass then vertical pass
// in float32 (OpenCV's order; its SIMD paths may contract differently: equal to a few ulp, not bit for bit), clipped
// to [0, 1] as :415 does.
__device__ __forceinline__ void cubic_weights(float t, float (&c)[4]) {
    const float A = -0.75f;
    c[0] = ((A * (t + 1.0f) - 5.0f * A) * (t + 1.0f) + 8.0f * A) * (t + 1.0f) - 4.0f * A;
    c[1] = ((A + 2.0f) * t - (A + 3.0f)) * t * t + 1.0f;
    c[2] = ((A + 2.0f) * (1.0f - t) - (A + 3.0f)) * (1.0f - t) * (1.0f - t) + 1.0f;
    c[3] = 1.0f - c[0] - c[1] - c[2];
}
__global__ void resize_cubic_kernel(const float* __restrict__ src, int W, int H, int C, float* __restrict__ dst, int w, int h) {
    const size_t n = (size_t)w * h;
    const double sx = (double)W / w, sy = (double)H / h;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % w), y = (int)(i / w);
        const float fx = (float)((x + 0.5) * sx - 0.5), fy = (float)((y + 0.5) * sy - 0.5);
        const int ix = (int)floorf(fx), iy = (int)floorf(fy);
        float cx[4], cy[4];
        cubic_weights(fx - (float)ix, cx);
        cubic_weights(fy - (float)iy, cy);
        for (int q = 0; q < C; ++q) {
            float rows[4];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int yy = min(max(iy - 1 + b, 0), H - 1);
                float acc = 0.0f;
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const int xx = min(max(ix - 1 + a, 0), W - 1);
                    acc = __fadd_rn(acc, __fmul_rn(src[((size_t)yy * W + xx) * C + q], cx[a]));
                }
                rows[b] = acc;
            }
            float v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(rows[0], cy[0]), __fmul_rn(rows[1], cy[1])), __fmul_rn(rows[2], cy[2])), __fmul_rn(rows[3], cy[3]));
            dst[i * C + q] = fminf(fmaxf(v, 0.0f), 1.0f);
        }
    }
}

}  // namespace

int launch_background_texture(mrtx_ctx* ctx, const float* rgb, int W, int H, float gamma, uint8_t* rgba) {
    background_texture_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(rgb, (size_t)W * H, (double)gamma, (uchar4*)rgba);
    MRTX_CUDA(cudaGetLastError());
    return MRTX_OK;
}

int launch_resize_cubic(mrtx_ctx* ctx, const float* src, int W, int H, int C, float* dst, int w, int h) {
    resize_cubic_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(src, W, H, C, dst, w, h);
    MRTX_CUDA(cudaGetLastError());
    return MRTX_OK;
}

int launch_color_reduce(mrtx_ctx* ctx, const uint8_t* bgr, int W, int H, int k,
                        const uint8_t* lut_dev, uint8_t* out) {
    const int h = H / k, w = W / k;
    const long long threads = (long long)((w + 3) / 4) * h;
    const unsigned blocks = (unsigned)((threads + 255) / 256);
    uchar4* o = reinterpret_cast<uchar4*>(out);
    switch (k) {
        case 1: color_reduce_kernel<1><<<blocks, 256, 0, ctx->stream>>>(bgr, W, h, w, lut_dev, o); break;
        case 2: color_reduce_kernel<2><<<blocks, 256, 0, ctx->stream>>>(bgr, W, h, w, lut_dev, o); break;
        case 4: color_reduce_kernel<4><<<blocks, 256, 0, ctx->stream>>>(bgr, W, h, w, lut_dev, o); break;
        case 8: color_reduce_kernel<8><<<blocks, 256, 0, ctx->stream>>>(bgr, W, h, w, lut_dev, o); break;
        default:
            mrtx_set_error("color downscale must be one of 1, 2, 4, 8 (data_loader.py:253)");
            return MRTX_ERR_INVALID;
    }
    MRTX_CUDA(cudaGetLastError());
    return MRTX_OK;
}
