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

}  // namespace

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
