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

#include "trace_common.cuh"

namespace {

// is re-based in float64 per candidate) in one tenth of the instructions, and is laid out for the SIMD
// width instead of around it:
//   * a warp owns 32 >> g_log2 neighbouring pixels, 2^g_log2 lanes share one pixel and trace one sample each,
//     so the lanes of a warp walk the same pyramid nodes (coherent loads, similar trip counts);
//   * while-while: lanes traverse until each holds a candidate patch (or has left the sphere), then all
//     candidates are tested at one instruction; primary and shadow rays run through the same loop body;
//   * per-pixel sums are reduced with shuffles: one accumulator read-modify-write per pixel;
//   * warps claim tasks from a counter, so limb / terminator warps that walk hundreds of cells do not leave
//     SMs idle at the end of the frame.
// A sample the filter cannot certify (FT_DEFER) contributes nothing here; its bit is set in the pixel's entry of
// the deferred list and trace_kernel_referee traces it again afterwards.
#ifndef MRTX_FAST_MINBLOCKS
#define MRTX_FAST_MINBLOCKS 8
#endif
template <bool I16>
__global__ void __launch_bounds__(128, MRTX_FAST_MINBLOCKS)
trace_kernel_fast(const __grid_constant__ RenderArgs A) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int gl = A.g_log2, g = 1 << gl;
    const int sub = lane & (g - 1), pw = lane >> gl;         // lane within the pixel's group, pixel within the warp
    const unsigned ppw = 32u >> gl;
    const unsigned n_limb = A.work_counter[4];
    const unsigned nkept = n_limb + A.work_counter[1];
    const unsigned ntasks = (nkept + ppw - 1u) / ppw;
    const float Rf = A.K.R;
    const unsigned rounds = (A.nsamples + (unsigned)g - 1u) >> gl;

    Counters cnt = {0u, 0u, 0u};
    RayStats rs = {0u, 0u, 0u, 0u, 0u};
    unsigned n_defer = 0;

    for (;;) {
        unsigned task = 0;
        if (lane == 0) task = atomicAdd(&A.work_counter[2], 1u);
        task = __shfl_sync(FULL, task, 0);
        if (task >= ntasks) break;
        const unsigned p = task * ppw + (unsigned)pw;
        const bool valid = p < nkept;
        const unsigned packed = valid ? list_pixel(A, p, n_limb) : 0u;
        const int x = (int)(packed & 0xffffu), y = (int)(packed >> 16);
        const uint32_t pixel = (uint32_t)y * (uint32_t)A.width + (uint32_t)x;
        float3 acc = make_float3(0.f, 0.f, 0.f);
        unsigned dmask = 0;                                  // group leader: deferred samples of this pixel

#pragma unroll 1
        for (unsigned rd = 0; rd < rounds; ++rd) {
            const unsigned k = rd * (unsigned)g + (unsigned)sub;
            const unsigned sm = A.sample0 + k;
            const bool active = valid && k < A.nsamples;
            Ray64 R;
            Walk st;
            FastHit fh;
            float3 lit = make_float3(0.f, 0.f, 0.f);
            bool defer = false, want = active, hit = false, entered = false, occluded = false, shadowed = false;
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {           // 0: primary ray, 1: shadow ray
                bool alive = false;
                if (want) {
                    if (pass == 0) primary_ray_fast(A, x, y, pixel, sm, R);
                    alive = walk_begin(A.hf, A.sp.radius, R, 0.0, pass ? (int)A.sp.start_shadow : A.hf.top - (int)A.sp.start_primary, st);
                    if (pass == 0) entered = alive;
                }
                int res = FT_MISS;
                while (__any_sync(FULL, alive)) {
                    RawPatch P;
                    float sx = 0.f;
                    int face = 4;
                    bool cand = false;
                    if (alive) {
                        for (;;) {
                            const int r = walk_step<I16>(A.hf, Rf, A.inv_rs, st, P, sx, face, cnt);
                            if (r == TR_CONTINUE) continue;
                            if (r == TR_END) alive = false; else cand = true;
                            break;
                        }
                    }
                    __syncwarp();
                    if ((alive || cand) && st.steps > (int)A.sp.long_walk) {
                        res = FT_DEFER; alive = false; cand = false;
                        atomicAdd(&A.defer_stats[15 + (pass ? 16 : 0)], 1ull);
                    }
                    if (cand) {
                        ++cnt.tests;
                        const int t = fast_test<I16>(A.hf, A.K, R, st.s_in, 0.0, st.s, sx, st.smax, P, pass != 0, fh);
                        if (t == FT_MISS) { if (!walk_advance(A.hf, st, sx, face)) alive = false; }
                        else {
                            res = t & 3; alive = false;
                            if (res == FT_DEFER) atomicAdd(&A.defer_stats[(t >> 2) + (pass ? 16 : 0)], 1ull);
                        }
                    }
                }
                if (res == FT_DEFER) defer = true;
                if (pass == 0) {
                    hit = res == FT_HIT;
                    want = false;
                    if (hit) {
                        Ray64 S;
                        want = shade_fast(A, R, fh, x, y, pixel, sm, lit, S);
                        shadowed = want;
                        if (want) R = S;
                    }
                    if (!__any_sync(FULL, want)) break;
                } else if (want) {
                    occluded = res == FT_HIT;
                }
            }
            if (active && !defer) {
                ++rs.primary;
                if (entered) ++rs.inside;
                if (hit) {
                    ++rs.hits;
                    if (shadowed) { ++rs.shadow; if (occluded) ++rs.occluded; }
                    if (!occluded) { acc.x += lit.x; acc.y += lit.y; acc.z += lit.z; }
                } else write_miss(A, x, y, sm == A.sample0);
            }
            // deferred samples of each pixel -> its leader's mask (bit = sample index in this launch)
            const unsigned dm = __ballot_sync(FULL, defer);
            if (dm) {
                const unsigned gm = g == 32 ? dm : (dm >> (pw << gl)) & ((1u << g) - 1u);
                dmask |= gm << (rd << gl);
                if (defer) ++n_defer;
            }
        }
        // per-pixel sum over the group's lanes, one read-modify-write per pixel
        for (int o = g >> 1; o > 0; o >>= 1) {
            acc.x += __shfl_xor_sync(FULL, acc.x, o);
            acc.y += __shfl_xor_sync(FULL, acc.y, o);
            acc.z += __shfl_xor_sync(FULL, acc.z, o);
        }
        if (valid && sub == 0) {
            float4* ap = A.accum + (size_t)y * A.width + x;
            float4 old = *ap;
            old.x += acc.x; old.y += acc.y; old.z += acc.z; old.w += (float)A.nsamples;
            *ap = old;
            if (dmask) A.defer_list[atomicAdd(&A.work_counter[3], 1u)] = make_uint2(packed, dmask);
        }
    }
    flush_counters(A, rs, cnt, lane);
    const unsigned nd = __reduce_add_sync(FULL, n_defer);
    if (lane == 0 && nd) atomicAdd(&A.defer_stats[0], (unsigned long long)nd);
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
    unsigned kernel = ctx->sp.kernel;
    if (kernel >= 2 && !make_fast_consts(ctx->hf, ctx->sp.radius).enabled) kernel = 1;   // map too coarse for the filter: everything would defer
    if (kernel != 2) return launch_trace_alt(ctx, x0, y0, x1, y1, s0, ns, kernel);
    RenderArgs A;
    fill_render_args(ctx, x0, y0, x1, y1, s0, ns, A);
    const bool i16 = ctx->hf.is_i16 != 0;
    int rc = launch_cull(ctx, A);
    if (rc) return rc;
    const long long npix = (long long)(x1 - x0) * (y1 - y0);
    // filtered kernel in chunks of <= 32 samples (one mask bit per sample in the deferred list), each
    // followed by the referee kernel over whatever it deferred
    int per_sm = 0;
    if (i16) MRTX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trace_kernel_fast<true>, 128, 0));
    else     MRTX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trace_kernel_fast<false>, 128, 0));
    if (per_sm < 1) per_sm = 1;
    for (unsigned done = 0; done < ns; done += 32u) {
        const unsigned n = ns - done < 32u ? ns - done : 32u;
        A.sample0 = s0 + done; A.nsamples = n;
        int gl = 0;
        while ((2u << gl) <= n && gl < 5) ++gl;
        A.g_log2 = gl;
        if (done) MRTX_CUDA(cudaMemsetAsync(A.work_counter + 2, 0, 2 * sizeof(unsigned), ctx->stream));
        const long long warps_needed = ((npix << gl) + 31) / 32;
        long long blocks = (long long)ctx->sm_count * per_sm;
        if (blocks * 4 > warps_needed) blocks = (warps_needed + 3) / 4;
        if (blocks < 1) blocks = 1;
        if (i16) trace_kernel_fast<true><<<(unsigned)blocks, 128, 0, ctx->stream>>>(A);
        else     trace_kernel_fast<false><<<(unsigned)blocks, 128, 0, ctx->stream>>>(A);
        if (i16) trace_kernel_referee<true, false><<<ctx->sm_count * 8, 64, 0, ctx->stream>>>(A);
        else     trace_kernel_referee<false, false><<<ctx->sm_count * 8, 64, 0, ctx->stream>>>(A);
    }
    MRTX_CUDA(cudaGetLastError());
    return MRTX_OK;
}

int launch_resolve(mrtx_ctx* ctx) {
    return launch_resolve_to(ctx, ctx->tex[1].data, ctx->rgba8);
}

int launch_resolve_to(mrtx_ctx* ctx, const uchar4* overlay, uchar4* out) {
    const size_t n = (size_t)ctx->width * ctx->height;
    resolve_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ctx->accum, overlay, out, n, ctx->sp.exposure, ctx->sp.inv_gamma);
    MRTX_CUDA(cudaGetLastError());
    return MRTX_OK;
}
