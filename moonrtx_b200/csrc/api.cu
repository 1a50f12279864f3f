// C ABI of libmoonb200.so: context lifetime, memory, timers, data_loader entry points,
// scene setters and frame read-back.  Kernels live in their own translation units.

#include "common.cuh"

#include <thread>
#include <atomic>
#include <vector>
#include <algorithm>
#include <fcntl.h>
#include <unistd.h>
#include <sys/types.h>

#include <stdarg.h>
#include <stdlib.h>

// ---- errors ------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void mrtx_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" {

int mrtx_abi_version(void) { return MRTX_ABI_VERSION; }
const char* mrtx_last_error(void) { return g_err; }

int mrtx_device_count(int* count) {
    MRTX_REQUIRE(count, "null argument");
    MRTX_CUDA(cudaGetDeviceCount(count));
    return MRTX_OK;
}

// ---- context -------------------------------------------------------------------------
int mrtx_create(int device, mrtx_ctx** out_ctx) {
    MRTX_REQUIRE(out_ctx, "null argument");
    *out_ctx = nullptr;
    int n = 0;
    MRTX_CUDA(cudaGetDeviceCount(&n));
    MRTX_REQUIRE(device >= 0 && device < n, "device %d out of range (%d visible)", device, n);
    MRTX_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    MRTX_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        // no fallback path exists: the kernels use sm_100-only instructions
        mrtx_set_error("device %d is sm_%d%d; libmoonb200 is built for sm_100a (B200) only",
                       device, prop.major, prop.minor);
        return MRTX_ERR_STATE;
    }
    mrtx_ctx* c = (mrtx_ctx*)calloc(1, sizeof(mrtx_ctx));
    MRTX_REQUIRE(c, "out of host memory");
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->l2_bytes = prop.l2CacheSize;
    c->hbm_bytes = prop.totalGlobalMem;
    MRTX_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    MRTX_CUDA(cudaEventCreate(&c->ev0));
    MRTX_CUDA(cudaEventCreate(&c->ev1));
    MRTX_CUDA(cudaMalloc(&c->d_max_bits, sizeof(unsigned)));
    MRTX_CUDA(cudaMalloc(&c->d_work, 16 * sizeof(unsigned)));
    MRTX_CUDA(cudaMalloc(&c->hard_buf, 256 * 2560));            // HARD_MAX_RAYS x sizeof(HardRay) (checked in trace.cu)
    MRTX_CUDA(cudaMalloc(&c->d_counters, 16 * sizeof(unsigned long long)));
    MRTX_CUDA(cudaMemset(c->d_counters, 0, 16 * sizeof(unsigned long long)));
    MRTX_CUDA(cudaMalloc(&c->d_defer_stats, 32 * sizeof(unsigned long long)));
    MRTX_CUDA(cudaMemset(c->d_defer_stats, 0, 32 * sizeof(unsigned long long)));
    // scene defaults = the reference's (moon_renderer.py:37, 85-101, 597-599, 620-621)
    SceneParams& sp = c->sp;
    sp.radius = 10.0;
    const double u[3] = {0, 0, 1}, v[3] = {0, -1, 0}, pos[3] = {0, 0, 0};
    mrtx_set_frame(c, pos, u, v, 10.0);
    sp.light_pos[0] = 21460.0; sp.light_radius = 100.0; sp.light_radiance = 80.0 * 460.5316;
    sp.scene_epsilon = 1.0e-4;
    sp.exposure = 0.9f; sp.inv_gamma = 1.0f / 2.2f;
    sp.jitter = 0; sp.shadows = 1; sp.debug_hits = 0; sp.kernel = 2; sp.beam = 0; sp.beam_drop = 2; sp.ceiling = 0; sp.shadow_queue = 4; sp.start_primary = 3; sp.start_shadow = 2; sp.long_walk = 2048u; sp.referee_budget = 500u; sp.hard_rays = 1u;
    const double eye[3] = {0, -300, 0}, tgt[3] = {0, 0, 0}, up[3] = {0, 0, 1};
    mrtx_set_camera(c, eye, tgt, up, 4.242192793);
    *out_ctx = c;
    return MRTX_OK;
}

static void free_frame(mrtx_ctx* c) {
    cudaFree(c->accum); cudaFree(c->rgba8); cudaFree(c->hit); cudaFree(c->hit64); cudaFree(c->pixel_list); cudaFree(c->defer_list); cudaFree(c->defer_mask);
    cudaFree(c->beam_s); cudaFree(c->beam_l); cudaFree(c->accfix);
    c->accfix = nullptr;
    c->pixel_list = nullptr; c->defer_list = nullptr; c->defer_mask = nullptr; c->beam_s = nullptr; c->beam_l = nullptr;
    c->accum = nullptr; c->rgba8 = nullptr; c->hit = nullptr; c->hit64 = nullptr;
}

static void pipe_release(mrtx_ctx* ctx);

int mrtx_destroy(mrtx_ctx* ctx) {
    if (!ctx) return MRTX_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    mrtx_comm_destroy(ctx);
    free_heightfield(ctx);
    for (int s = 0; s < 3; ++s) cudaFree(ctx->tex_owned[s]);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    pipe_release(ctx);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->comm_stream) { cudaStreamSynchronize(ctx->comm_stream); cudaStreamDestroy(ctx->comm_stream); }
    p2p_release(ctx);
    for (int q = 0; q < 2; ++q) { cudaFree(ctx->recv_buf[q]); if (ctx->recv_ev[q]) cudaEventDestroy(ctx->recv_ev[q]); }
    free_frame(ctx);
    cudaFree(ctx->d_max_bits);
    cudaFree(ctx->d_work); cudaFree(ctx->hard_buf);
    cudaFree(ctx->d_counters);
    cudaFree(ctx->d_defer_stats);
    cudaFree(ctx->wave_buf);
    cudaFree(ctx->sq_buf); cudaFree(ctx->hq_buf); cudaFree(ctx->bq_buf[0]); cudaFree(ctx->bq_buf[1]); cudaFree(ctx->pool_buf);
    cudaFree(ctx->tube_seg); cudaFree(ctx->tube_tiles);
    if (ctx->prof_ev) { for (int i = 0; i < MRTX_PROF_MAX * MRTX_PROF_EVENTS; ++i) cudaEventDestroy(ctx->prof_ev[i]); free(ctx->prof_ev); }
    cudaFree(ctx->flush_buf);
    if (ctx->h_rs) cudaFreeHost(ctx->h_rs);
    for (int b = 0; b < 2; ++b) { if (ctx->stage[b]) cudaFreeHost(ctx->stage[b]); if (ctx->stage_ev[b]) cudaEventDestroy(ctx->stage_ev[b]); }
    cudaFree(ctx->gather_buf);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    cudaStreamDestroy(ctx->own_stream);
    free(ctx);
    return MRTX_OK;
}

int mrtx_synchronize(mrtx_ctx* ctx) {
    MRTX_CTX(ctx);
    MRTX_CUDA(cudaStreamSynchronize(ctx->stream));
    return MRTX_OK;
}

int mrtx_set_stream(mrtx_ctx* ctx, void* cuda_stream) {
    MRTX_CTX(ctx);
    MRTX_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return MRTX_OK;
}

int mrtx_get_stream(mrtx_ctx* ctx, void** cuda_stream) {
    MRTX_CTX(ctx);
    MRTX_REQUIRE(cuda_stream, "null argument");
    *cuda_stream = (void*)ctx->stream;
    return MRTX_OK;
}

int mrtx_device_props(mrtx_ctx* ctx, int* sm_count, int* l2_bytes, size_t* hbm_bytes) {
    MRTX_CTX(ctx);
    if (sm_count) *sm_count = ctx->sm_count;
    if (l2_bytes) *l2_bytes = ctx->l2_bytes;
    if (hbm_bytes) *hbm_bytes = ctx->hbm_bytes;
    return MRTX_OK;
}

int mrtx_timer_start(mrtx_ctx* ctx) {
    MRTX_CTX(ctx);
    MRTX_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    return MRTX_OK;
}

int mrtx_timer_stop(mrtx_ctx* ctx, float* elapsed_ms) {
    MRTX_CTX(ctx);
    MRTX_REQUIRE(elapsed_ms, "null argument");
    MRTX_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    MRTX_CUDA(cudaEventSynchronize(ctx->ev1));
    MRTX_CUDA(cudaEventElapsedTime(elapsed_ms, ctx->ev0, ctx->ev1));
    return MRTX_OK;
}

// ---- memory ------------------------------------------------------------------------
int mrtx_dev_alloc(mrtx_ctx* ctx, size_t bytes, void** out_dev) {
    MRTX_CTX(ctx);
    MRTX_REQUIRE(out_dev, "null argument");
    MRTX_CUDA(cudaMalloc(out_dev, bytes ? bytes : 1));
    return MRTX_OK;
}
int mrtx_dev_free(mrtx_ctx* ctx, void* ptr_dev) {
    MRTX_CTX(ctx);
    MRTX_CUDA(cudaStreamSynchronize(ctx->stream));
    MRTX_CUDA(cudaFree(ptr_dev));
    return MRTX_OK;
}
int mrtx_host_alloc(size_t bytes, void** out_pinned) {
    MRTX_REQUIRE(out_pinned, "null argument");
    MRTX_CUDA(cudaMallocHost(out_pinned, bytes ? bytes : 1));
    return MRTX_OK;
}
int mrtx_host_free(void* pinned) {
    MRTX_CUDA(cudaFreeHost(pinned));
    return MRTX_OK;
}
int mrtx_h2d(mrtx_ctx* ctx, void* dst_dev, const void* src, size_t bytes) {
    MRTX_CTX(ctx);
    MRTX_CUDA(cudaMemcpyAsync(dst_dev, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return MRTX_OK;
}
int mrtx_d2h(mrtx_ctx* ctx, void* dst, const void* src_dev, size_t bytes) {
    MRTX_CTX(ctx);
    MRTX_CUDA(cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    MRTX_CUDA(cudaStreamSynchronize(ctx->stream));
    return MRTX_OK;
}
int mrtx_l2_flush(mrtx_ctx* ctx) {
    MRTX_CTX(ctx);
    const size_t want = (size_t)ctx->l2_bytes * 2 + (64u << 20);
    if (ctx->flush_bytes < want) {
        cudaFree(ctx->flush_buf);
        ctx->flush_buf = nullptr; ctx->flush_bytes = 0;
        MRTX_CUDA(cudaMalloc(&ctx->flush_buf, want));
        ctx->flush_bytes = want;
    }
    MRTX_CUDA(cudaMemsetAsync(ctx->flush_buf, 0x5a, ctx->flush_bytes, ctx->stream));
    return MRTX_OK;
}

// ---- data_loader ---------------------------------------------------------------------
static int check_downscale_args(const void* src, int W, int H, int ds, const void* out) {
    MRTX_REQUIRE(src && out, "null buffer");
    MRTX_REQUIRE(W > 0 && H > 0, "bad map size %d x %d", W, H);
    MRTX_REQUIRE(ds >= 1 && ds <= 512, "downscale %d outside 1..512 (row sums must stay exact in float32)", ds);
    // numpy's reshape(1, h, ds, w, ds) raises ValueError here (data_loader.py:225)
    MRTX_REQUIRE(W % ds == 0 && H % ds == 0,
                 "cannot reshape array of size %lld into shape (1,%d,%d,%d,%d)",
                 (long long)W * H, H / ds, ds, W / ds, ds);
    return MRTX_OK;
}

static int ensure_rs_word(mrtx_ctx* ctx) {
    if (ctx->h_rs) return MRTX_OK;
    MRTX_CUDA(cudaHostAlloc((void**)&ctx->h_rs, sizeof(float), cudaHostAllocMapped));
    MRTX_CUDA(cudaHostGetDevicePointer((void**)&ctx->h_rs_dev, ctx->h_rs, 0));
    return MRTX_OK;
}

int mrtx_downscale_i16_dev(mrtx_ctx* ctx, const int16_t* src_dev, int W, int H, int ds,
                           float* out_dev, float* radius_scale) {
    MRTX_CTX(ctx);
    int rc = check_downscale_args(src_dev, W, H, ds, out_dev);
    if (rc) return rc;
    if (radius_scale) { rc = ensure_rs_word(ctx); if (rc) return rc; }
    rc = launch_downscale_i16(ctx, src_dev, W, H, ds, out_dev, radius_scale ? ctx->h_rs_dev : nullptr);
    if (rc) return rc;
    if (radius_scale) {
        // the kernel has written the value to mapped host memory: nothing to copy, only to wait for
        MRTX_CUDA(cudaStreamSynchronize(ctx->stream));
        *radius_scale = *(volatile float*)ctx->h_rs;
    }
    return MRTX_OK;
}

// memcpy on a few host threads (one thread moves ~10 GB/s, the copy engine five times that)
static void parallel_memcpy(void* dst, const void* src, size_t bytes) {
    unsigned nt = std::thread::hardware_concurrency() / 2;
    nt = nt < 1 ? 1 : (nt > 8 ? 8 : nt);
    if (bytes < ((size_t)4 << 20) || nt == 1) { memcpy(dst, src, bytes); return; }
    std::thread th[8];
    const size_t per = ((bytes / nt) + 4095) & ~(size_t)4095;
    unsigned used = 0;
    for (unsigned i = 0; i < nt; ++i) {
        const size_t off = (size_t)i * per;
        if (off >= bytes) break;
        const size_t len = bytes - off < per ? bytes - off : per;
        th[used++] = std::thread([=] { memcpy((char*)dst + off, (const char*)src + off, len); });
    }
    for (unsigned i = 0; i < used; ++i) th[i].join();
}

static int ensure_staging(mrtx_ctx* ctx, size_t bytes) {
    if (ctx->stage_bytes >= bytes) return MRTX_OK;
    MRTX_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int b = 0; b < 2; ++b) {
        if (ctx->stage[b]) cudaFreeHost(ctx->stage[b]);
        ctx->stage[b] = nullptr;
        MRTX_CUDA(cudaHostAlloc(&ctx->stage[b], bytes, cudaHostAllocDefault));
        if (!ctx->stage_ev[b]) MRTX_CUDA(cudaEventCreateWithFlags(&ctx->stage_ev[b], cudaEventDisableTiming));
    }
    ctx->stage_bytes = bytes;
    return MRTX_OK;
}

// Host buffers in and out (what load_elevation_data calls, data_loader.py:215-247).  The caller's arrays are pageable:
// the map goes up in bands through two pinned staging buffers - host threads fill one while the copy engine drains the
// other and the block-mean kernel reduces the band before - so the call costs about what the slowest of the three
// (the host-side memcpy) costs; the result comes down the same way.
// `fill(dst, first_row, rows)` puts source rows into a pinned staging buffer (from the caller's array, or straight from the
// strips of a TIFF file); `npy`, if given, receives the result as it comes down (the downscale cache, data_loader.py:88-95).
extern "C++" {
template <typename Fill>
static int downscale_stream(mrtx_ctx* ctx, int W, int H, int ds, float* out, float* radius_scale, Fill fill, FILE* npy) {
    int rc = ensure_rs_word(ctx);
    if (rc) return rc;
    const int h = H / ds, w = W / ds;
    const size_t in_bytes = (size_t)W * H * sizeof(int16_t);
    const size_t out_bytes = (size_t)w * h * sizeof(float);
    // bands of whole output rows, about 32 MB of source each
    size_t band_out_rows = ((size_t)32 << 20) / ((size_t)W * 2 * ds);
    if (band_out_rows < 1) band_out_rows = 1;
    const size_t band_rows = band_out_rows * ds, band_bytes = band_rows * (size_t)W * 2;
    rc = ensure_staging(ctx, band_bytes);
    if (rc) return rc;
    int16_t* d_src = nullptr; float* d_out = nullptr;
    MRTX_CUDA(cudaMalloc(&d_src, in_bytes));
    cudaError_t e = cudaMalloc(&d_out, out_bytes);
    if (e != cudaSuccess) { cudaFree(d_src); mrtx_set_error("cudaMalloc: %s", cudaGetErrorString(e)); return MRTX_ERR_CUDA; }
    cudaStream_t st = ctx->stream;
    rc = downscale_begin(ctx);
    int b = 0;
    for (size_t r0 = 0; !rc && r0 < (size_t)H; r0 += band_rows, b ^= 1) {
        const size_t rows = (size_t)H - r0 < band_rows ? (size_t)H - r0 : band_rows;
        const size_t bytes = rows * (size_t)W * 2;
        if (cudaEventSynchronize(ctx->stage_ev[b]) != cudaSuccess) { rc = MRTX_ERR_CUDA; break; }     // the copy out of this buffer two bands ago
        rc = fill(ctx->stage[b], r0, rows);
        if (rc) break;
        if (cudaMemcpyAsync(d_src + r0 * (size_t)W, ctx->stage[b], bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) { rc = MRTX_ERR_CUDA; break; }
        if (cudaEventRecord(ctx->stage_ev[b], st) != cudaSuccess) { rc = MRTX_ERR_CUDA; break; }
        rc = downscale_band(ctx, d_src + r0 * (size_t)W, W, (int)rows, ds, d_out + (r0 / ds) * (size_t)w);
    }
    if (!rc) rc = downscale_finish(ctx, d_out, (size_t)w * h, ctx->h_rs_dev);
    // the result: device -> pinned staging -> the caller's array (and the cache file), two chunks in flight
    const size_t chunk = ctx->stage_bytes;
    size_t done = 0, copied = 0;
    int q = 0;
    size_t pend_off[2] = {0, 0}, pend_len[2] = {0, 0};
    bool pend[2] = {false, false};
    while (!rc && (done < out_bytes || pend[0] || pend[1])) {
        if (pend[q]) {
            if (cudaEventSynchronize(ctx->stage_ev[q]) != cudaSuccess) { rc = MRTX_ERR_CUDA; break; }
            parallel_memcpy((char*)out + pend_off[q], ctx->stage[q], pend_len[q]);
            if (npy && fwrite(ctx->stage[q], 1, pend_len[q], npy) != pend_len[q]) { mrtx_set_error("cache file: short write"); rc = MRTX_ERR_INVALID; break; }
            copied += pend_len[q]; pend[q] = false;
        }
        if (done < out_bytes) {
            const size_t len = out_bytes - done < chunk ? out_bytes - done : chunk;
            if (cudaMemcpyAsync(ctx->stage[q], (const char*)d_out + done, len, cudaMemcpyDeviceToHost, st) != cudaSuccess) { rc = MRTX_ERR_CUDA; break; }
            if (cudaEventRecord(ctx->stage_ev[q], st) != cudaSuccess) { rc = MRTX_ERR_CUDA; break; }
            pend_off[q] = done; pend_len[q] = len; pend[q] = true; done += len;
        }
        q ^= 1;
    }
    if (!rc && cudaStreamSynchronize(st) != cudaSuccess) rc = MRTX_ERR_CUDA;
    if (!rc) *radius_scale = *(volatile float*)ctx->h_rs;
    if (rc == MRTX_ERR_CUDA) mrtx_set_error("downscale: %s", cudaGetErrorString(cudaGetLastError()));
    cudaStreamSynchronize(st);
    cudaFree(d_src); cudaFree(d_out);
    (void)copied;
    return rc;
}
}  // extern "C++"

int mrtx_downscale_i16(mrtx_ctx* ctx, const int16_t* src, int W, int H, int ds,
                       float* out, float* radius_scale) {
    MRTX_CTX(ctx);
    int rc = check_downscale_args(src, W, H, ds, out);
    if (rc) return rc;
    MRTX_REQUIRE(radius_scale, "null radius_scale");
    return downscale_stream(ctx, W, H, ds, out, radius_scale, [=](void* dst, size_t r0, size_t rows) {
        parallel_memcpy(dst, src + r0 * (size_t)W, rows * (size_t)W * 2);
        return (int)MRTX_OK;
    }, nullptr);
}

// ---- the LDEM file itself -------------------------------------------------------------------------
// plotoptix.utils.read_image (data_loader.py:206) decodes the whole 8.5 GB TIFF into a host array before the first
// element is reduced.  The LOLA LDEM is an uncompressed single-channel 16-bit strip TIFF (BigTIFF at 128 px/deg): its
// strips ARE the row-major array, so they are read with pread() straight into the pinned staging buffers of the banded
// upload above - no decoded copy, no second pass over host memory.
struct TiffInfo { int W, H, bits, samples, compression, little, tiled, format; unsigned long long rows_per_strip; std::vector<unsigned long long> off, cnt; };

static unsigned long long tiff_get(const unsigned char* p, int bytes, bool le) {
    unsigned long long v = 0;
    for (int i = 0; i < bytes; ++i) v |= (unsigned long long)p[le ? i : bytes - 1 - i] << (8 * i);
    return v;
}

static int tiff_parse(const char* path, TiffInfo& T) {
    FILE* f = fopen(path, "rb");
    if (!f) { mrtx_set_error("cannot open %s", path); return MRTX_ERR_INVALID; }
    unsigned char hd[16];
    int rc = MRTX_ERR_INVALID;
    T = TiffInfo();
    T.samples = 1; T.compression = 1; T.format = 1; T.rows_per_strip = ~0ull;
    do {
        if (fread(hd, 1, 16, f) != 16) { mrtx_set_error("%s: not a TIFF file", path); break; }
        const bool le = hd[0] == 'I' && hd[1] == 'I';
        if (!le && !(hd[0] == 'M' && hd[1] == 'M')) { mrtx_set_error("%s: not a TIFF file", path); break; }
        T.little = le;
        const unsigned magic = (unsigned)tiff_get(hd + 2, 2, le);
        const bool big = magic == 43;
        if (magic != 42 && !big) { mrtx_set_error("%s: not a TIFF file", path); break; }
        const unsigned long long ifd = big ? tiff_get(hd + 8, 8, le) : tiff_get(hd + 4, 4, le);
        if (fseeko(f, (off_t)ifd, SEEK_SET)) { mrtx_set_error("%s: bad directory offset", path); break; }
        unsigned char nb[8];
        if (fread(nb, 1, big ? 8 : 2, f) != (size_t)(big ? 8 : 2)) { mrtx_set_error("%s: truncated", path); break; }
        const unsigned long long n = tiff_get(nb, big ? 8 : 2, le);
        const int esz = big ? 20 : 12;
        if (n > 4096) { mrtx_set_error("%s: implausible directory", path); break; }
        std::vector<unsigned char> dir((size_t)n * esz);
        if (fread(dir.data(), 1, dir.size(), f) != dir.size()) { mrtx_set_error("%s: truncated", path); break; }
        static const int tsz[] = {0, 1, 1, 2, 4, 8, 1, 1, 2, 4, 8, 4, 8, 4, 0, 0, 8, 8, 8};
        bool bad = false;
        auto values = [&](const unsigned char* e, std::vector<unsigned long long>& out) {
            const unsigned type = (unsigned)tiff_get(e + 2, 2, le);
            const unsigned long long count = big ? tiff_get(e + 4, 8, le) : tiff_get(e + 4, 4, le);
            const int sz = type < 19 ? tsz[type] : 0;
            if (!sz || count > ((unsigned long long)1 << 28)) { bad = true; return; }
            const unsigned char* vp = e + (big ? 12 : 8);
            const size_t inline_bytes = big ? 8 : 4;
            std::vector<unsigned char> buf;
            if (count * sz > inline_bytes) {
                const unsigned long long at = big ? tiff_get(vp, 8, le) : tiff_get(vp, 4, le);
                buf.resize((size_t)count * sz);
                if (fseeko(f, (off_t)at, SEEK_SET) || fread(buf.data(), 1, buf.size(), f) != buf.size()) { bad = true; return; }
                vp = buf.data();
            }
            out.resize((size_t)count);
            for (size_t i = 0; i < (size_t)count; ++i) out[i] = tiff_get(vp + i * sz, sz, le);
        };
        for (unsigned long long k = 0; k < n && !bad; ++k) {
            const unsigned char* e = dir.data() + (size_t)k * esz;
            const unsigned tag = (unsigned)tiff_get(e, 2, le);
            std::vector<unsigned long long> v;
            switch (tag) {
                case 256: values(e, v); if (!v.empty()) T.W = (int)v[0]; break;
                case 257: values(e, v); if (!v.empty()) T.H = (int)v[0]; break;
                case 258: values(e, v); if (!v.empty()) T.bits = (int)v[0]; break;
                case 259: values(e, v); if (!v.empty()) T.compression = (int)v[0]; break;
                case 277: values(e, v); if (!v.empty()) T.samples = (int)v[0]; break;
                case 278: values(e, v); if (!v.empty()) T.rows_per_strip = v[0]; break;
                case 339: values(e, v); if (!v.empty()) T.format = (int)v[0]; break;
                case 273: values(e, T.off); break;
                case 279: values(e, T.cnt); break;
                case 322: case 324: T.tiled = 1; break;
                default: break;
            }
        }
        if (bad || T.W <= 0 || T.H <= 0) { mrtx_set_error("%s: unreadable TIFF directory", path); break; }
        rc = MRTX_OK;
    } while (0);
    fclose(f);
    return rc;
}

// can the strips be streamed as they lie?  (else the caller decodes the file on the host, as the reference does)
static bool tiff_streamable(const TiffInfo& T) {
    if (T.tiled || T.compression != 1 || T.samples != 1 || T.bits != 16 || !T.little || T.off.empty() || T.off.size() != T.cnt.size()) return false;
    const unsigned long long rps = T.rows_per_strip > (unsigned long long)T.H ? (unsigned long long)T.H : T.rows_per_strip;
    if (!rps || T.off.size() != ((unsigned long long)T.H + rps - 1) / rps) return false;
    for (size_t i = 0; i < T.off.size(); ++i) {
        const unsigned long long rows = i + 1 < T.off.size() ? rps : (unsigned long long)T.H - rps * i;
        if (T.cnt[i] != rows * (unsigned long long)T.W * 2) return false;
    }
    return true;
}

int mrtx_tiff_info(const char* path, int* W, int* H, int* bits_per_sample, int* streamable) {
    MRTX_REQUIRE(path && W && H && bits_per_sample && streamable, "null argument");
    TiffInfo T;
    const int rc = tiff_parse(path, T);
    if (rc) return rc;
    *W = T.W; *H = T.H; *bits_per_sample = T.bits; *streamable = tiff_streamable(T) ? 1 : 0;
    return MRTX_OK;
}

int mrtx_downscale_tiff_i16(mrtx_ctx* ctx, const char* path, int ds, float* out, float* radius_scale, const char* npy_cache_path) {
    MRTX_CTX(ctx);
    MRTX_REQUIRE(path && out && radius_scale, "null argument");
    TiffInfo T;
    int rc = tiff_parse(path, T);
    if (rc) return rc;
    if (!tiff_streamable(T)) { mrtx_set_error("%s: not an uncompressed little-endian 16-bit single-channel strip TIFF", path); return MRTX_ERR_INVALID; }
    rc = check_downscale_args(path, T.W, T.H, ds, out);
    if (rc) return rc;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) { mrtx_set_error("cannot open %s", path); return MRTX_ERR_INVALID; }
    FILE* npy = nullptr;
    if (npy_cache_path) {
        npy = fopen(npy_cache_path, "wb");
        if (!npy) { close(fd); mrtx_set_error("cannot write %s", npy_cache_path); return MRTX_ERR_INVALID; }
        // numpy format 1.0: magic, version, little-endian u16 header length, dict padded with spaces to a multiple of 64, '\n'
        char dict[160];
        int n = snprintf(dict, sizeof dict, "{'descr': '<f4', 'fortran_order': False, 'shape': (%d, %d), }", T.H / ds, T.W / ds);
        const int total = ((10 + n + 1 + 63) / 64) * 64;
        const int hlen = total - 10;
        unsigned char pre[10] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0, (unsigned char)(hlen & 255), (unsigned char)(hlen >> 8)};
        fwrite(pre, 1, 10, npy);
        fwrite(dict, 1, (size_t)n, npy);
        for (int i = n; i < hlen - 1; ++i) fputc(' ', npy);
        fputc('\n', npy);
    }
    const unsigned long long rps = T.rows_per_strip > (unsigned long long)T.H ? (unsigned long long)T.H : T.rows_per_strip;
    const size_t row_bytes = (size_t)T.W * 2;
    rc = downscale_stream(ctx, T.W, T.H, ds, out, radius_scale, [&](void* dst, size_t r0, size_t rows) {
        // rows [r0, r0 + rows) lie in strips r0 / rps ...: one pread per strip piece, in parallel for large bands
        struct Piece { off_t at; size_t len, dst; };
        std::vector<Piece> pieces;
        for (size_t r = r0; r < r0 + rows;) {
            const size_t sidx = r / rps, in_strip = r - sidx * rps;
            const size_t take = std::min<size_t>(rps - in_strip, r0 + rows - r);
            pieces.push_back({(off_t)(T.off[sidx] + in_strip * row_bytes), take * row_bytes, (r - r0) * row_bytes});
            r += take;
        }
        std::atomic<int> failed{0};
        auto run = [&](size_t a, size_t b) {
            for (size_t i = a; i < b; ++i) {
                size_t got = 0;
                while (got < pieces[i].len) {
                    const ssize_t k = pread(fd, (char*)dst + pieces[i].dst + got, pieces[i].len - got, pieces[i].at + (off_t)got);
                    if (k <= 0) { failed = 1; return; }
                    got += (size_t)k;
                }
            }
        };
        unsigned nt = std::thread::hardware_concurrency() / 2;
        nt = nt < 1 ? 1 : (nt > 8 ? 8 : nt);
        if (pieces.size() < 2 * nt) run(0, pieces.size());
        else {
            std::thread th[8];
            const size_t per = (pieces.size() + nt - 1) / nt;
            unsigned used = 0;
            for (unsigned i = 0; i < nt && (size_t)i * per < pieces.size(); ++i)
                th[used++] = std::thread(run, (size_t)i * per, std::min(pieces.size(), (size_t)(i + 1) * per));
            for (unsigned i = 0; i < used; ++i) th[i].join();
        }
        if (failed) { mrtx_set_error("%s: read error", path); return (int)MRTX_ERR_INVALID; }
        return (int)MRTX_OK;
    }, npy);
    close(fd);
    if (npy) { if (fclose(npy) != 0 && !rc) { mrtx_set_error("cache file: write error"); rc = MRTX_ERR_INVALID; } if (rc) remove(npy_cache_path); }
    return rc;
}

static int check_color_args(const void* bgr, int W, int H, int k, const void* lut, const void* out) {
    MRTX_REQUIRE(bgr && lut && out, "null buffer");
    MRTX_REQUIRE(W > 0 && H > 0, "bad image size %d x %d", W, H);
    MRTX_REQUIRE(k == 1 || k == 2 || k == 4 || k == 8, "color downscale must be one of 1, 2, 4, 8");
    MRTX_REQUIRE(W % k == 0 && H % k == 0, "image size %d x %d is not divisible by %d", W, H, k);
    return MRTX_OK;
}

int mrtx_color_reduce_lut_dev(mrtx_ctx* ctx, const uint8_t* bgr_dev, int W, int H, int k,
                              const uint8_t lut[256], uint8_t* out_rgba_dev) {
    MRTX_CTX(ctx);
    int rc = check_color_args(bgr_dev, W, H, k, lut, out_rgba_dev);
    if (rc) return rc;
    uint8_t* d_lut = nullptr;
    MRTX_CUDA(cudaMalloc(&d_lut, 256));
    MRTX_CUDA(cudaMemcpyAsync(d_lut, lut, 256, cudaMemcpyHostToDevice, ctx->stream));
    rc = launch_color_reduce(ctx, bgr_dev, W, H, k, d_lut, out_rgba_dev);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_lut);
    return rc;
}

int mrtx_color_reduce_lut(mrtx_ctx* ctx, const uint8_t* bgr, int W, int H, int k,
                          const uint8_t lut[256], uint8_t* out_rgba) {
    MRTX_CTX(ctx);
    int rc = check_color_args(bgr, W, H, k, lut, out_rgba);
    if (rc) return rc;
    const size_t in_bytes = (size_t)W * H * 3, out_bytes = (size_t)(W / k) * (H / k) * 4;
    uint8_t *d_in = nullptr, *d_out = nullptr;
    MRTX_CUDA(cudaMalloc(&d_in, in_bytes));
    cudaError_t e = cudaMalloc(&d_out, out_bytes);
    if (e != cudaSuccess) { cudaFree(d_in); mrtx_set_error("cudaMalloc: %s", cudaGetErrorString(e)); return MRTX_ERR_CUDA; }
    rc = MRTX_OK;
    if (cudaMemcpyAsync(d_in, bgr, in_bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) rc = MRTX_ERR_CUDA;
    if (!rc) rc = mrtx_color_reduce_lut_dev(ctx, d_in, W, H, k, lut, d_out);
    if (!rc && cudaMemcpyAsync(out_rgba, d_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) rc = MRTX_ERR_CUDA;
    if (!rc && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = MRTX_ERR_CUDA;
    if (rc == MRTX_ERR_CUDA) mrtx_set_error("color_reduce: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(d_in); cudaFree(d_out);
    return rc;
}

int mrtx_synth_ldem_i16_dev(mrtx_ctx* ctx, int16_t* out_dev, int W, int H, uint32_t seed) {
    MRTX_CTX(ctx);
    MRTX_REQUIRE(out_dev && W > 0 && H > 0, "bad arguments");
    return launch_synth_ldem(ctx, out_dev, W, H, seed);
}
int mrtx_synth_color_bgr_dev(mrtx_ctx* ctx, uint8_t* out_dev, int W, int H, uint32_t seed) {
    MRTX_CTX(ctx);
    MRTX_REQUIRE(out_dev && W > 0 && H > 0, "bad arguments");
    return launch_synth_color(ctx, out_dev, W, H, seed);
}

// ---- scene ---------------------------------------------------------------------------
static void norm3(double* a) {
    const double n = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    if (n > 0) { a[0] /= n; a[1] /= n; a[2] /= n; }
}
static void cross3(const double* a, const double* b, double* c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

}  // extern "C"

int prof_mark(mrtx_ctx* ctx, int which) {
    if (!ctx->prof_on || ctx->prof_n >= MRTX_PROF_MAX) return MRTX_OK;
    if (!ctx->prof_ev) {
        ctx->prof_ev = (cudaEvent_t*)calloc((size_t)MRTX_PROF_MAX * MRTX_PROF_EVENTS, sizeof(cudaEvent_t));
        if (!ctx->prof_ev) { mrtx_set_error("out of host memory"); return MRTX_ERR_INVALID; }
        for (int i = 0; i < MRTX_PROF_MAX * MRTX_PROF_EVENTS; ++i) MRTX_CUDA(cudaEventCreate(&ctx->prof_ev[i]));
    }
    MRTX_CUDA(cudaEventRecord(ctx->prof_ev[(size_t)ctx->prof_n * MRTX_PROF_EVENTS + which], ctx->stream));
    if (which == MRTX_PROF_EVENTS - 1) ctx->prof_n += 1;
    return MRTX_OK;
}

void free_heightfield(mrtx_ctx* ctx) {
    if (ctx->hf_owned_base) cudaFree(ctx->hf_owned_base);
    if (ctx->hf_levels_owned) cudaFree(ctx->hf_levels_owned);
    if (ctx->hf_tables_owned) cudaFree(ctx->hf_tables_owned);
    ctx->hf_tables_owned = nullptr;
    ctx->hf_owned_base = nullptr;
    ctx->hf_levels_owned = nullptr;
    memset(&ctx->hf, 0, sizeof(ctx->hf));
}

extern "C" {

static int set_displacement_common(mrtx_ctx* ctx, const void* map, int W, int H, int is_i16,
                                   float scale, float radius_scale, int from_host, int copy) {
    MRTX_CTX(ctx);
    MRTX_REQUIRE(map, "null map");
    MRTX_REQUIRE(W >= 8 && H >= 4, "displacement map %d x %d too small (need >= 8 x 4)", W, H);
    MRTX_REQUIRE(!is_i16 || (radius_scale > 0.0f), "radius_scale must be positive");
    MRTX_CUDA(cudaStreamSynchronize(ctx->stream));
    free_heightfield(ctx);
    const size_t bytes = (size_t)W * H * (is_i16 ? 2 : 4);
    const void* base = map;
    if (from_host || copy) {
        MRTX_CUDA(cudaMalloc(&ctx->hf_owned_base, bytes));
        MRTX_CUDA(cudaMemcpyAsync(ctx->hf_owned_base, map, bytes,
                                  from_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, ctx->stream));
        base = ctx->hf_owned_base;
    }
    ctx->hf.base = base;
    ctx->hf.is_i16 = is_i16;
    ctx->hf.W = W; ctx->hf.H = H;
    ctx->hf.scale = scale; ctx->hf.radius_scale = radius_scale;
    int rc = build_pyramid(ctx);
    if (rc) free_heightfield(ctx);
    return rc;
}

int mrtx_set_displacement_f32(mrtx_ctx* ctx, const float* map, int W, int H) {
    return set_displacement_common(ctx, map, W, H, 0, 0.f, 1.f, 1, 1);
}
int mrtx_set_displacement_f32_dev(mrtx_ctx* ctx, const float* map_dev, int W, int H, int copy) {
    return set_displacement_common(ctx, map_dev, W, H, 0, 0.f, 1.f, 0, copy);
}
int mrtx_set_displacement_i16(mrtx_ctx* ctx, const int16_t* map, int W, int H, float scale, float radius_scale) {
    return set_displacement_common(ctx, map, W, H, 1, scale, radius_scale, 1, 1);
}
int mrtx_set_displacement_i16_dev(mrtx_ctx* ctx, const int16_t* map_dev, int W, int H, float scale,
                                  float radius_scale, int copy) {
    return set_displacement_common(ctx, map_dev, W, H, 1, scale, radius_scale, 0, copy);
}

int mrtx_set_texture_rgba8(mrtx_ctx* ctx, int slot, const uint8_t* rgba, int W, int H) {
    MRTX_CTX(ctx);
    MRTX_REQUIRE(slot >= 0 && slot <= 2, "texture slot %d (0 = moon_color, 1 = frame_overlay, 2 = environment)", slot);
    MRTX_CUDA(cudaStreamSynchronize(ctx->stream));
    if (!rgba) {
        cudaFree(ctx->tex_owned[slot]);
        ctx->tex_owned[slot] = nullptr;
        ctx->tex[slot].data = nullptr; ctx->tex[slot].W = ctx->tex[slot].H = 0;
        return MRTX_OK;
    }
    MRTX_REQUIRE(W > 0 && H > 0, "bad texture size");
    const size_t bytes = (size_t)W * H * 4;
    if (ctx->tex[slot].W != W || ctx->tex[slot].H != H || !ctx->tex_owned[slot]) {
        cudaFree(ctx->tex_owned[slot]);
        ctx->tex_owned[slot] = nullptr;
        MRTX_CUDA(cudaMalloc(&ctx->tex_owned[slot], bytes));
    }
    MRTX_CUDA(cudaMemcpyAsync(ctx->tex_owned[slot], rgba, bytes, cudaMemcpyHostToDevice, ctx->stream));
    MRTX_CUDA(cudaStreamSynchronize(ctx->stream));   // the caller may drop its array right away
    ctx->tex[slot].data = (const uchar4*)ctx->tex_owned[slot];
    ctx->tex[slot].W = W; ctx->tex[slot].H = H;
    return MRTX_OK;
}

int mrtx_set_background_f32(mrtx_ctx* ctx, const float* rgb, int W, int H, float gamma) {
    MRTX_CTX(ctx);
    if (!rgb) return mrtx_set_texture_rgba8(ctx, 2, nullptr, 0, 0);
    MRTX_REQUIRE(W > 1 && H > 1 && gamma > 0.0f, "bad background %d x %d, gamma %g", W, H, (double)gamma);
    MRTX_CUDA(cudaStreamSynchronize(ctx->stream));
    const size_t n = (size_t)W * H;
    float* d_rgb = nullptr;
    MRTX_CUDA(cudaMalloc(&d_rgb, n * 3 * sizeof(float)));
    cudaFree(ctx->tex_owned[2]);
    ctx->tex_owned[2] = nullptr; ctx->tex[2].data = nullptr;
    cudaError_t e = cudaMalloc(&ctx->tex_owned[2], n * 4);
    if (e != cudaSuccess) { cudaFree(d_rgb); mrtx_set_error("cudaMalloc: %s", cudaGetErrorString(e)); return MRTX_ERR_CUDA; }
    int rc = MRTX_OK;
    if (cudaMemcpyAsync(d_rgb, rgb, n * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) rc = MRTX_ERR_CUDA;
    if (!rc) rc = launch_background_texture(ctx, d_rgb, W, H, gamma, (uint8_t*)ctx->tex_owned[2]);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = MRTX_ERR_CUDA;
    cudaFree(d_rgb);
    if (rc) { mrtx_set_error("background: %s", cudaGetErrorString(cudaGetLastError())); return rc; }
    ctx->tex[2].data = (const uchar4*)ctx->tex_owned[2];
    ctx->tex[2].W = W; ctx->tex[2].H = H;
    return MRTX_OK;
}

int mrtx_read_background_rgba8(mrtx_ctx* ctx, uint8_t* out, int* W, int* H) {
    MRTX_CTX(ctx);
    if (W) *W = ctx->tex[2].W;
    if (H) *H = ctx->tex[2].H;
    if (out && ctx->tex[2].data) {
        MRTX_CUDA(cudaMemcpyAsync(out, ctx->tex[2].data, (size_t)ctx->tex[2].W * ctx->tex[2].H * 4, cudaMemcpyDeviceToHost, ctx->stream));
        MRTX_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return MRTX_OK;
}

int mrtx_set_tubes(mrtx_ctx* ctx, const float* segments, int n) {
    MRTX_CTX(ctx);
    MRTX_REQUIRE(n >= 0 && (n == 0 || segments), "bad segment list");
    MRTX_CUDA(cudaStreamSynchronize(ctx->stream));             // (a frame in flight may still read the old list)
    if (ctx->copy_stream) MRTX_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    if ((unsigned)n > ctx->tube_cap) {
        cudaFree(ctx->tube_seg); ctx->tube_seg = nullptr; ctx->tube_cap = 0; ctx->n_tubes = 0;
        const unsigned cap = (unsigned)n + (unsigned)n / 2u + 64u;
        MRTX_CUDA(cudaMalloc(&ctx->tube_seg, (size_t)cap * 3 * sizeof(float4)));
        ctx->tube_cap = cap;
    }
    if (n) MRTX_CUDA(cudaMemcpy(ctx->tube_seg, segments, (size_t)n * 3 * sizeof(float4), cudaMemcpyHostToDevice));
    ctx->n_tubes = (unsigned)n;
    return MRTX_OK;
}

int mrtx_set_sun_disk(mrtx_ctx* ctx, const double center[3], double radius, const float color[3]) {
    MRTX_REQUIRE(ctx, "null context");
    SceneParams& sp = ctx->sp;
    sp.sun_disk_radius = radius > 0.0 && center && color ? radius : 0.0;
    if (sp.sun_disk_radius > 0.0)
        for (int i = 0; i < 3; ++i) { sp.sun_disk_pos[i] = center[i]; sp.sun_disk_color[i] = color[i]; }
    return MRTX_OK;
}

int mrtx_resize_cubic_f32(mrtx_ctx* ctx, const float* src, int W, int H, int channels, float* dst, int w, int h) {
    MRTX_CTX(ctx);
    MRTX_REQUIRE(src && dst && W > 0 && H > 0 && w > 0 && h > 0 && channels >= 1 && channels <= 4, "bad resize arguments");
    const size_t nin = (size_t)W * H * channels * sizeof(float), nout = (size_t)w * h * channels * sizeof(float);
    float* d_in = nullptr; float* d_out = nullptr;
    MRTX_CUDA(cudaMalloc(&d_in, nin));
    cudaError_t e = cudaMalloc(&d_out, nout);
    if (e != cudaSuccess) { cudaFree(d_in); mrtx_set_error("cudaMalloc: %s", cudaGetErrorString(e)); return MRTX_ERR_CUDA; }
    int rc = MRTX_OK;
    if (cudaMemcpyAsync(d_in, src, nin, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) rc = MRTX_ERR_CUDA;
    if (!rc) rc = launch_resize_cubic(ctx, d_in, W, H, channels, d_out, w, h);
    if (!rc && cudaMemcpyAsync(dst, d_out, nout, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) rc = MRTX_ERR_CUDA;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = MRTX_ERR_CUDA;
    cudaFree(d_in); cudaFree(d_out);
    if (rc == MRTX_ERR_CUDA) mrtx_set_error("resize: %s", cudaGetErrorString(cudaGetLastError()));
    return rc;
}

int mrtx_set_frame(mrtx_ctx* ctx, const double pos[3], const double u[3], const double v[3], double radius) {
    MRTX_REQUIRE(ctx && pos && u && v, "null argument");
    MRTX_REQUIRE(radius > 0, "radius must be positive");
    SceneParams& sp = ctx->sp;
    double ez[3] = {u[0], u[1], u[2]};
    norm3(ez);
    // v is the scene direction of longitude 0 (body -Y); make it orthogonal to u
    double vv[3] = {v[0], v[1], v[2]};
    const double d = vv[0] * ez[0] + vv[1] * ez[1] + vv[2] * ez[2];
    for (int i = 0; i < 3; ++i) vv[i] -= d * ez[i];
    MRTX_REQUIRE(vv[0] * vv[0] + vv[1] * vv[1] + vv[2] * vv[2] > 1e-24, "u and v are parallel");
    norm3(vv);
    double ex[3];
    cross3(ez, vv, ex);          // body +X (lon +90 E) = u x v
    for (int i = 0; i < 3; ++i) {
        sp.ex[i] = ex[i]; sp.ey[i] = -vv[i]; sp.ez[i] = ez[i]; sp.pos[i] = pos[i];
    }
    sp.radius = radius;
    return MRTX_OK;
}

int mrtx_set_camera(mrtx_ctx* ctx, const double eye[3], const double target[3], const double up[3], double fov_deg) {
    MRTX_REQUIRE(ctx && eye && target && up, "null argument");
    MRTX_REQUIRE(fov_deg > 0.0 && fov_deg < 180.0, "fov %g outside (0, 180)", fov_deg);
    Camera& c = ctx->cam;
    double w[3] = {target[0] - eye[0], target[1] - eye[1], target[2] - eye[2]};
    MRTX_REQUIRE(w[0] * w[0] + w[1] * w[1] + w[2] * w[2] > 0, "eye == target");
    norm3(w);
    double r[3];
    cross3(w, up, r);
    MRTX_REQUIRE(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] > 1e-24, "up is parallel to the view direction");
    norm3(r);
    double u2[3];
    cross3(r, w, u2);
    for (int i = 0; i < 3; ++i) { c.eye[i] = eye[i]; c.w[i] = w[i]; c.right[i] = r[i]; c.up[i] = u2[i]; }
    c.tan_half_fov = tan(fov_deg * 0.5 * 3.14159265358979323846 / 180.0);
    return MRTX_OK;
}

int mrtx_set_light(mrtx_ctx* ctx, const double pos[3], double radius, double radiance) {
    MRTX_REQUIRE(ctx && pos, "null argument");
    MRTX_REQUIRE(radius >= 0 && radiance >= 0, "negative light radius / radiance");
    for (int i = 0; i < 3; ++i) ctx->sp.light_pos[i] = pos[i];
    ctx->sp.light_radius = radius;
    ctx->sp.light_radiance = radiance;
    return MRTX_OK;
}

int mrtx_set_float(mrtx_ctx* ctx, const char* name, double value) {
    MRTX_REQUIRE(ctx && name, "null argument");
    SceneParams& sp = ctx->sp;
    if (!strcmp(name, "scene_epsilon")) { MRTX_REQUIRE(value >= 0, "scene_epsilon < 0"); sp.scene_epsilon = value; }
    else if (!strcmp(name, "tonemap_exposure")) sp.exposure = (float)value;
    else if (!strcmp(name, "tonemap_gamma")) { MRTX_REQUIRE(value > 0, "gamma <= 0"); sp.inv_gamma = (float)(1.0 / value); }
    else if (!strcmp(name, "marching_step") || !strcmp(name, "marching_step_eps")) { /* exact intersection: unused */ }
    else { mrtx_set_error("unknown float parameter '%s'", name); return MRTX_ERR_INVALID; }
    return MRTX_OK;
}

int mrtx_set_uint(mrtx_ctx* ctx, const char* name, unsigned a, unsigned b) {
    MRTX_REQUIRE(ctx && name, "null argument");
    if (!strcmp(name, "path_seg_range")) {
        // (min, max) ray segments of a path, moon_renderer.py:583: the camera ray and the light ray are two; every segment
        // beyond is one diffuse interreflection bounce (traced to the maximum, no Russian roulette)
        ctx->sp.n_bounce = b > 2u ? (b - 2u > 4u ? 4u : b - 2u) : 0u;
    }
    else if (!strcmp(name, "jitter")) ctx->sp.jitter = a ? 1u : 0u;
    else if (!strcmp(name, "shadows")) ctx->sp.shadows = a ? 1u : 0u;
    else if (!strcmp(name, "debug_hits")) ctx->sp.debug_hits = a ? 1u : 0u;
    else if (!strcmp(name, "start_levels")) { ctx->sp.start_primary = a; ctx->sp.start_shadow = b; }
    else if (!strcmp(name, "long_walk")) { MRTX_REQUIRE(a >= 1u, "long_walk must be >= 1"); ctx->sp.long_walk = a; }
    else if (!strcmp(name, "hard_rays")) ctx->sp.hard_rays = a ? 1u : 0u;
    else if (!strcmp(name, "referee_budget")) { MRTX_REQUIRE(a >= 1u, "referee_budget must be >= 1"); ctx->sp.referee_budget = a; }
    else if (!strcmp(name, "ceiling")) ctx->sp.ceiling = a;
    else if (!strcmp(name, "profile")) ctx->prof_on = a ? 1 : 0;
    else if (!strcmp(name, "blocks_per_sm")) ctx->sp.blocks_per_sm = a;
    else if (!strcmp(name, "shadow_queue")) ctx->sp.shadow_queue = a > 4u ? 4u : a;
    else if (!strcmp(name, "beam")) { ctx->sp.beam = a ? 1u : 0u; ctx->sp.beam_drop = b; }
    else if (!strcmp(name, "kernel")) { MRTX_REQUIRE(a <= 3u, "kernel must be 0, 1, 2 or 3"); ctx->sp.kernel = a; }
    else { mrtx_set_error("unknown uint parameter '%s'", name); return MRTX_ERR_INVALID; }
    return MRTX_OK;
}

static int alloc_frame(mrtx_ctx* ctx, size_t n);

int mrtx_resize(mrtx_ctx* ctx, int width, int height) {
    MRTX_CTX(ctx);
    MRTX_REQUIRE(width > 0 && height > 0 && width <= 65536 && height <= 65536, "bad frame size %d x %d", width, height);
    if (width == ctx->width && height == ctx->height && ctx->accum) return MRTX_OK;
    MRTX_CUDA(cudaStreamSynchronize(ctx->stream));
    free_frame(ctx);
    ctx->width = ctx->height = 0;                           // until every buffer of the new size exists
    const size_t n = (size_t)width * height;
    const int rc = alloc_frame(ctx, n);
    if (rc) { free_frame(ctx); return rc; }
    ctx->width = width; ctx->height = height;
    return MRTX_OK;
}

static int alloc_frame(mrtx_ctx* ctx, size_t n) {
    MRTX_CUDA(cudaMalloc(&ctx->accum, n * sizeof(float4)));
    MRTX_CUDA(cudaMalloc(&ctx->rgba8, n * sizeof(uchar4)));
    MRTX_CUDA(cudaMalloc(&ctx->hit, n * sizeof(float4)));
    MRTX_CUDA(cudaMalloc(&ctx->pixel_list, n * sizeof(unsigned)));
    MRTX_CUDA(cudaMalloc(&ctx->defer_list, n * sizeof(uint2)));
    MRTX_CUDA(cudaMalloc(&ctx->defer_mask, n * sizeof(unsigned)));
    MRTX_CUDA(cudaMemsetAsync(ctx->defer_mask, 0, n * sizeof(unsigned), ctx->stream));
    MRTX_CUDA(cudaMalloc(&ctx->accfix, n * 3 * sizeof(unsigned long long)));
    MRTX_CUDA(cudaMemsetAsync(ctx->accfix, 0, n * 3 * sizeof(unsigned long long), ctx->stream));
    MRTX_CUDA(cudaMalloc(&ctx->beam_s, n * sizeof(double)));
    MRTX_CUDA(cudaMalloc(&ctx->beam_l, n));
    MRTX_CUDA(cudaMemsetAsync(ctx->accum, 0, n * sizeof(float4), ctx->stream));
    MRTX_CUDA(cudaMemsetAsync(ctx->rgba8, 0, n * sizeof(uchar4), ctx->stream));
    MRTX_CUDA(cudaMemsetAsync(ctx->hit, 0, n * sizeof(float4), ctx->stream));
    return MRTX_OK;
}

// ---- pipelined frames ---------------------------------------------------------------------------
// The synchronous sequence (set overlay texture, render, resolve, read back) leaves the GPU idle while 33 MB go in
// and 33 MB come out of every 4K frame and while the host gets round to the next one.  mrtx_frame_submit() queues a
// whole frame - overlay upload, trace, resolve, read-back into the caller's pinned buffer - and returns; the copies
// run on a second stream against the NEXT / PREVIOUS frame's tracing.  Scene parameters are captured at submit.
static void pipe_release(mrtx_ctx* ctx) {
    for (int k = 0; k < 2; ++k) {
        cudaFree(ctx->pipe_overlay[k]); cudaFree(ctx->pipe_rgba8[k]);
        ctx->pipe_overlay[k] = nullptr; ctx->pipe_rgba8[k] = nullptr;
        if (ctx->pipe_ev_upload[k]) cudaEventDestroy(ctx->pipe_ev_upload[k]);
        if (ctx->pipe_ev_resolve[k]) cudaEventDestroy(ctx->pipe_ev_resolve[k]);
        if (ctx->pipe_ev_d2h[k]) cudaEventDestroy(ctx->pipe_ev_d2h[k]);
        ctx->pipe_ev_upload[k] = ctx->pipe_ev_resolve[k] = ctx->pipe_ev_d2h[k] = nullptr;
        ctx->pipe_busy[k] = 0;
    }
    ctx->pipe_w = ctx->pipe_h = 0;
}

static int pipe_ensure(mrtx_ctx* ctx) {
    if (ctx->pipe_w == ctx->width && ctx->pipe_h == ctx->height && ctx->pipe_rgba8[0]) return MRTX_OK;
    MRTX_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->copy_stream) MRTX_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    else MRTX_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    pipe_release(ctx);
    const size_t bytes = (size_t)ctx->width * ctx->height * sizeof(uchar4);
    for (int k = 0; k < 2; ++k) {
        MRTX_CUDA(cudaMalloc(&ctx->pipe_overlay[k], bytes));
        MRTX_CUDA(cudaMalloc(&ctx->pipe_rgba8[k], bytes));
        MRTX_CUDA(cudaEventCreateWithFlags(&ctx->pipe_ev_upload[k], cudaEventDisableTiming));
        MRTX_CUDA(cudaEventCreateWithFlags(&ctx->pipe_ev_resolve[k], cudaEventDisableTiming));
        MRTX_CUDA(cudaEventCreateWithFlags(&ctx->pipe_ev_d2h[k], cudaEventDisableTiming));
        // (recorded once so that the first waits on them return at once)
        MRTX_CUDA(cudaEventRecord(ctx->pipe_ev_resolve[k], ctx->stream));
        MRTX_CUDA(cudaEventRecord(ctx->pipe_ev_d2h[k], ctx->copy_stream));
    }
    ctx->pipe_w = ctx->width; ctx->pipe_h = ctx->height; ctx->pipe_slot = 0;
    return MRTX_OK;
}

static int frame_submit(mrtx_ctx* ctx, const uint8_t* overlay_rgba_pinned, unsigned nsamples, uint8_t* out_rgba_pinned,
                        int dst_rank, int* ticket);

int mrtx_frame_submit(mrtx_ctx* ctx, const uint8_t* overlay_rgba_pinned, unsigned nsamples, uint8_t* out_rgba_pinned, int* ticket) {
    MRTX_REQUIRE(out_rgba_pinned, "null output buffer");
    return frame_submit(ctx, overlay_rgba_pinned, nsamples, out_rgba_pinned, -1, ticket);
}

int mrtx_frame_submit_to(mrtx_ctx* ctx, const uint8_t* overlay_rgba_pinned, unsigned nsamples, int dst_rank, int* ticket) {
    MRTX_CTX(ctx);
    if (!ctx->nccl_comm && !ctx->p2p_on) { mrtx_set_error("neither mrtx_comm_init nor mrtx_p2p_connect has been called"); return MRTX_ERR_STATE; }
    MRTX_REQUIRE(dst_rank >= 0 && dst_rank < ctx->nranks && dst_rank != ctx->rank, "bad destination rank %d", dst_rank);
    return frame_submit(ctx, overlay_rgba_pinned, nsamples, nullptr, dst_rank, ticket);
}

static int comm_stream_ensure(mrtx_ctx* ctx) {
    if (!ctx->comm_stream) MRTX_CUDA(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
    return MRTX_OK;
}

int mrtx_frame_recv(mrtx_ctx* ctx, int src_rank, uint8_t* out_rgba_pinned, int* ticket) {
    MRTX_CTX(ctx);
    MRTX_REQUIRE(out_rgba_pinned && ticket, "null argument");
    if (!ctx->nccl_comm && !ctx->p2p_on) { mrtx_set_error("neither mrtx_comm_init nor mrtx_p2p_connect has been called"); return MRTX_ERR_STATE; }
    if (!ctx->accum) { mrtx_set_error("mrtx_resize has not been called"); return MRTX_ERR_STATE; }
    MRTX_REQUIRE(src_rank >= 0 && src_rank < ctx->nranks && src_rank != ctx->rank, "bad source rank %d", src_rank);
    int rc = comm_stream_ensure(ctx);
    if (rc) return rc;
    const size_t bytes = (size_t)ctx->width * ctx->height * sizeof(uchar4);
    for (int q = 0; q < 2; ++q)
        if (!ctx->recv_ev[q]) MRTX_CUDA(cudaEventCreateWithFlags(&ctx->recv_ev[q], cudaEventDisableTiming));
    if (!ctx->p2p_on && ctx->recv_bytes != bytes) {
        MRTX_CUDA(cudaStreamSynchronize(ctx->comm_stream));
        for (int q = 0; q < 2; ++q) {
            cudaFree(ctx->recv_buf[q]); ctx->recv_buf[q] = nullptr;
            MRTX_CUDA(cudaMalloc(&ctx->recv_buf[q], bytes));
            ctx->recv_busy[q] = 0;
        }
        ctx->recv_bytes = bytes; ctx->recv_slot = 0;
    }
    const int q = ctx->recv_slot;
    if (ctx->recv_busy[q]) {
        mrtx_set_error("two received frames are pending: mrtx_frame_recv_wait(%d) must be called first", q);
        return MRTX_ERR_STATE;
    }
    if (ctx->p2p_on) {
        // the sender's copy engine has put (or will put) the frame into this rank's mailbox: no staging copy, no kernel
        rc = p2p_recv_frame(ctx, out_rgba_pinned, bytes, src_rank, ctx->comm_stream);
        if (rc) return rc;
    } else {
        rc = comm_recv_bytes(ctx, ctx->recv_buf[q], bytes, src_rank, ctx->comm_stream);
        if (rc) return rc;
        MRTX_CUDA(cudaMemcpyAsync(out_rgba_pinned, ctx->recv_buf[q], bytes, cudaMemcpyDeviceToHost, ctx->comm_stream));
    }
    MRTX_CUDA(cudaEventRecord(ctx->recv_ev[q], ctx->comm_stream));
    ctx->recv_busy[q] = 1; ctx->recv_slot = q ^ 1;
    *ticket = q;
    return MRTX_OK;
}

int mrtx_frame_recv_wait(mrtx_ctx* ctx, int ticket) {
    MRTX_CTX(ctx);
    MRTX_REQUIRE((ticket == 0 || ticket == 1) && ctx->recv_ev[ticket] && ctx->recv_busy[ticket], "no such received frame pending");
    MRTX_CUDA(cudaEventSynchronize(ctx->recv_ev[ticket]));
    ctx->recv_busy[ticket] = 0;
    return MRTX_OK;
}

static int frame_submit(mrtx_ctx* ctx, const uint8_t* overlay_rgba_pinned, unsigned nsamples, uint8_t* out_rgba_pinned,
                        int dst_rank, int* ticket) {
    MRTX_CTX(ctx);
    MRTX_REQUIRE(ticket && nsamples > 0, "null argument / no samples");
    if (!ctx->accum) { mrtx_set_error("mrtx_resize has not been called"); return MRTX_ERR_STATE; }
    if (!ctx->hf.base) { mrtx_set_error("no displacement map set"); return MRTX_ERR_STATE; }
    int rc = pipe_ensure(ctx);
    if (rc) return rc;
    const int k = ctx->pipe_slot;
    if (ctx->pipe_busy[k]) {
        mrtx_set_error("two frames are in flight: mrtx_frame_wait(%d) must be called before the next submit", k);
        return MRTX_ERR_STATE;
    }
    const size_t n = (size_t)ctx->width * ctx->height, bytes = n * sizeof(uchar4);
    if (overlay_rgba_pinned) {
        // slot k's overlay was last read by the resolve of the frame two submits ago
        MRTX_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->pipe_ev_resolve[k], 0));
        MRTX_CUDA(cudaMemcpyAsync(ctx->pipe_overlay[k], overlay_rgba_pinned, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
        MRTX_CUDA(cudaEventRecord(ctx->pipe_ev_upload[k], ctx->copy_stream));
    }
    MRTX_CUDA(cudaMemsetAsync(ctx->accum, 0, n * sizeof(float4), ctx->stream));
    rc = launch_trace(ctx, 0, 0, ctx->width, ctx->height, 0, nsamples);
    if (rc) return rc;
    if (overlay_rgba_pinned) MRTX_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->pipe_ev_upload[k], 0));
    MRTX_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->pipe_ev_d2h[k], 0));      // slot k's output has left the device
    rc = launch_resolve_to(ctx, overlay_rgba_pinned ? ctx->pipe_overlay[k] : nullptr, ctx->pipe_rgba8[k]);
    if (rc) return rc;
    MRTX_CUDA(cudaEventRecord(ctx->pipe_ev_resolve[k], ctx->stream));
    if (dst_rank < 0) {
        MRTX_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->pipe_ev_resolve[k], 0));
        MRTX_CUDA(cudaMemcpyAsync(out_rgba_pinned, ctx->pipe_rgba8[k], bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
        MRTX_CUDA(cudaEventRecord(ctx->pipe_ev_d2h[k], ctx->copy_stream));
    } else {
        // the frame leaves over NVLink instead: ncclSend from the slot, which is free again when the send has completed
        rc = comm_stream_ensure(ctx);
        if (rc) return rc;
        MRTX_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->pipe_ev_resolve[k], 0));
        rc = ctx->p2p_on ? p2p_send_frame(ctx, ctx->pipe_rgba8[k], bytes, dst_rank, ctx->comm_stream)
                         : comm_send_bytes(ctx, ctx->pipe_rgba8[k], bytes, dst_rank, ctx->comm_stream);
        if (rc) return rc;
        MRTX_CUDA(cudaEventRecord(ctx->pipe_ev_d2h[k], ctx->comm_stream));
    }
    ctx->pipe_busy[k] = 1;
    ctx->pipe_slot = k ^ 1;                                 // (only a frame that was queued completely takes its slot)
    *ticket = k;
    return MRTX_OK;
}

int mrtx_frame_wait(mrtx_ctx* ctx, int ticket) {
    MRTX_CTX(ctx);
    MRTX_REQUIRE((ticket == 0 || ticket == 1) && ctx->pipe_ev_d2h[ticket], "no such frame in flight");
    MRTX_CUDA(cudaEventSynchronize(ctx->pipe_ev_d2h[ticket]));
    ctx->pipe_busy[ticket] = 0;
    return MRTX_OK;
}

// ---- render / read-back ----------------------------------------------------------------
int mrtx_render(mrtx_ctx* ctx, int x0, int y0, int x1, int y1, unsigned sample0, unsigned nsamples, int reset) {
    MRTX_CTX(ctx);
    if (!ctx->accum) { mrtx_set_error("mrtx_resize has not been called"); return MRTX_ERR_STATE; }
    if (!ctx->hf.base) { mrtx_set_error("no displacement map set"); return MRTX_ERR_STATE; }
    MRTX_REQUIRE(0 <= x0 && x0 <= x1 && x1 <= ctx->width && 0 <= y0 && y0 <= y1 && y1 <= ctx->height,
                 "render rectangle [%d,%d)x[%d,%d) outside the %d x %d frame", x0, x1, y0, y1, ctx->width, ctx->height);
    const size_t n = (size_t)ctx->width * ctx->height;
    if (reset) MRTX_CUDA(cudaMemsetAsync(ctx->accum, 0, n * sizeof(float4), ctx->stream));
    if (ctx->sp.debug_hits && !ctx->hit64) MRTX_CUDA(cudaMalloc(&ctx->hit64, n * sizeof(double4)));
    if (nsamples == 0 || x0 == x1 || y0 == y1) return MRTX_OK;
    return launch_trace(ctx, x0, y0, x1, y1, sample0, nsamples);
}

int mrtx_render_tiles(mrtx_ctx* ctx, int tile, unsigned sample0, unsigned nsamples, int reset) {
    MRTX_CTX(ctx);
    int tl = 0;
    while ((1 << tl) < tile) ++tl;
    MRTX_REQUIRE(tile >= 8 && tile <= 1024 && (1 << tl) == tile, "tile side must be a power of two in 8..1024");
    if (!ctx->nccl_comm || ctx->nranks < 1) { mrtx_set_error("mrtx_comm_init has not been called"); return MRTX_ERR_STATE; }
    ctx->tile_log2 = tl;
    const int rc = mrtx_render(ctx, 0, 0, ctx->width, ctx->height, sample0, nsamples, reset);
    ctx->tile_log2 = 0;
    return rc;
}

int mrtx_resolve(mrtx_ctx* ctx) {
    MRTX_CTX(ctx);
    if (!ctx->accum) { mrtx_set_error("mrtx_resize has not been called"); return MRTX_ERR_STATE; }
    if (ctx->tex[1].data)
        MRTX_REQUIRE(ctx->tex[1].W == ctx->width && ctx->tex[1].H == ctx->height,
                     "frame_overlay is %d x %d, frame is %d x %d", ctx->tex[1].W, ctx->tex[1].H, ctx->width, ctx->height);
    return launch_resolve(ctx);
}

static int read_back(mrtx_ctx* ctx, void* out, const void* src, size_t elem) {
    MRTX_CTX(ctx);
    MRTX_REQUIRE(out, "null buffer");
    if (!src) { mrtx_set_error("frame buffer not allocated"); return MRTX_ERR_STATE; }
    MRTX_CUDA(cudaMemcpyAsync(out, src, (size_t)ctx->width * ctx->height * elem, cudaMemcpyDeviceToHost, ctx->stream));
    MRTX_CUDA(cudaStreamSynchronize(ctx->stream));
    return MRTX_OK;
}
int mrtx_read_rgba8(mrtx_ctx* ctx, uint8_t* out) { return read_back(ctx, out, ctx ? ctx->rgba8 : nullptr, 4); }
int mrtx_read_accum_f32(mrtx_ctx* ctx, float* out) { return read_back(ctx, out, ctx ? ctx->accum : nullptr, 16); }
int mrtx_read_hit_f32(mrtx_ctx* ctx, float* out) { return read_back(ctx, out, ctx ? ctx->hit : nullptr, 16); }
int mrtx_read_hit_f64(mrtx_ctx* ctx, double* out) { return read_back(ctx, out, ctx ? ctx->hit64 : nullptr, 32); }

int mrtx_hit_at(mrtx_ctx* ctx, int x, int y, float out4[4]) {
    MRTX_CTX(ctx);
    MRTX_REQUIRE(out4, "null buffer");
    if (!ctx->hit) { mrtx_set_error("frame buffer not allocated"); return MRTX_ERR_STATE; }
    MRTX_REQUIRE(x >= 0 && x < ctx->width && y >= 0 && y < ctx->height, "pixel (%d, %d) outside the frame", x, y);
    MRTX_CUDA(cudaMemcpyAsync(out4, ctx->hit + (size_t)y * ctx->width + x, 16, cudaMemcpyDeviceToHost, ctx->stream));
    MRTX_CUDA(cudaStreamSynchronize(ctx->stream));
    return MRTX_OK;
}

int mrtx_frame_buffers_dev(mrtx_ctx* ctx, void** accum_dev, void** rgba8_dev, void** hit_dev) {
    MRTX_CTX(ctx);
    if (accum_dev) *accum_dev = ctx->accum;
    if (rgba8_dev) *rgba8_dev = ctx->rgba8;
    if (hit_dev) *hit_dev = ctx->hit;
    return MRTX_OK;
}

int mrtx_kernel_times(mrtx_ctx* ctx, double out_ms[8], int reset) {
    MRTX_CTX(ctx);
    MRTX_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < ctx->prof_n; ++i) {
        cudaEvent_t* e = ctx->prof_ev + (size_t)i * MRTX_PROF_EVENTS;
        for (int k = 1; k < MRTX_PROF_EVENTS; ++k) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, e[k - 1], e[k]) == cudaSuccess) ctx->prof_ms[k - 1] += (double)ms;
        }
        ctx->prof_launches += 1;
    }
    (void)cudaGetLastError();
    ctx->prof_n = 0;
    if (out_ms) {
        // intervals: 0 cull, 1 beam, 2 trace_kernel_fast, 3 shade_kernel, 4 shadow_kernel, 5 referee, 6 fold
        out_ms[0] = ctx->prof_ms[0]; out_ms[1] = ctx->prof_ms[1]; out_ms[2] = ctx->prof_ms[2];
        out_ms[3] = ctx->prof_ms[4]; out_ms[4] = ctx->prof_ms[5]; out_ms[5] = ctx->prof_ms[6];
        out_ms[6] = (double)ctx->prof_launches; out_ms[7] = ctx->prof_ms[3];
    }
    if (reset) { for (int k = 0; k < 8; ++k) ctx->prof_ms[k] = 0.0; ctx->prof_launches = 0; }
    return MRTX_OK;
}

int mrtx_counters(mrtx_ctx* ctx, uint64_t out[16], int reset) {
    MRTX_CTX(ctx);
    if (out) {
        MRTX_CUDA(cudaMemcpyAsync(out, ctx->d_counters, 128, cudaMemcpyDeviceToHost, ctx->stream));
        MRTX_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    if (reset) MRTX_CUDA(cudaMemsetAsync(ctx->d_counters, 0, 128, ctx->stream));
    return MRTX_OK;
}

int mrtx_defer_stats(mrtx_ctx* ctx, uint64_t out[32], int reset) {
    MRTX_CTX(ctx);
    if (out) {
        MRTX_CUDA(cudaMemcpyAsync(out, ctx->d_defer_stats, 256, cudaMemcpyDeviceToHost, ctx->stream));
        MRTX_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    if (reset) MRTX_CUDA(cudaMemsetAsync(ctx->d_defer_stats, 0, 256, ctx->stream));
    return MRTX_OK;
}

}  // extern "C"
