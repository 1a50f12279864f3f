// Multi-GPU exchange steps: one process per GPU, NCCL over NVLink.
//
// The renderer replicates the height field and shards WORK (frames, samples, screen
// tiles), so the data path needs a collective only at the end of a sharded frame:
//   - progressive-sample split -> ncclAllReduce(sum, f32) of the float4 accumulators;
//   - screen-tile split        -> ncclAllGather of interleaved row tiles of the RGBA8 frame.
// Frame-parallel time-lapse needs no collective at all.
//
// NCCL is opened with dlopen at run time (the torch-bundled libnccl.so.2); the ABI
// stays free of NCCL types: the unique id crosses it as 128 opaque bytes.

#include "common.cuh"

#include <dlfcn.h>

namespace {

typedef struct { char internal[128]; } nccl_uid;
typedef void* nccl_comm_t;
enum { NCCL_UINT8 = 1, NCCL_FLOAT32 = 7, NCCL_SUM = 0 };

struct NcclApi {
    void* lib;
    int (*GetUniqueId)(nccl_uid*);
    int (*CommInitRank)(nccl_comm_t*, int, nccl_uid, int);
    int (*CommDestroy)(nccl_comm_t);
    int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t);
    int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t);
    int (*Send)(const void*, size_t, int, int, nccl_comm_t, cudaStream_t);
    int (*Recv)(void*, size_t, int, int, nccl_comm_t, cudaStream_t);
    const char* (*GetErrorString)(int);
};

NcclApi g_nccl = {};

int load_nccl(const char* path) {
    if (g_nccl.lib) return MRTX_OK;
    void* lib = dlopen(path && *path ? path : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) {
        mrtx_set_error("dlopen(%s): %s", path ? path : "libnccl.so.2", dlerror());
        return MRTX_ERR_NCCL;
    }
#define SYM(field, name)                                                         \
    *(void**)(&g_nccl.field) = dlsym(lib, name);                                 \
    if (!g_nccl.field) { mrtx_set_error("dlsym(%s) failed", name); dlclose(lib); return MRTX_ERR_NCCL; }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllReduce, "ncclAllReduce")
    SYM(AllGather, "ncclAllGather")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.lib = lib;
    return MRTX_OK;
}

#define MRTX_NCCL(call)                                                          \
    do {                                                                         \
        int r_ = (call);                                                         \
        if (r_ != 0) {                                                           \
            mrtx_set_error("%s failed: %s", #call, g_nccl.GetErrorString(r_));   \
            return MRTX_ERR_NCCL;                                                \
        }                                                                        \
    } while (0)

// rows of tile t (global) <-> slot (t / nranks) of rank (t % nranks)
__global__ void pack_rows_kernel(const uchar4* __restrict__ frame, uchar4* __restrict__ send, int W, int H,
                                 int tile_rows, int nranks, int rank, int tiles_per_rank) {
    const size_t n = (size_t)tiles_per_rank * tile_rows * W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        const int lr = (int)(i / W);
        const int y = ((lr / tile_rows) * nranks + rank) * tile_rows + lr % tile_rows;
        send[i] = y < H ? frame[(size_t)y * W + x] : make_uchar4(0, 0, 0, 0);
    }
}
__global__ void unpack_rows_kernel(uchar4* __restrict__ frame, const uchar4* __restrict__ recv, int W, int H,
                                   int tile_rows, int nranks, int tiles_per_rank) {
    const size_t per_rank = (size_t)tiles_per_rank * tile_rows * W;
    const size_t n = per_rank * nranks;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int q = (int)(i / per_rank);
        const size_t j = i - (size_t)q * per_rank;
        const int x = (int)(j % W);
        const int lr = (int)(j / W);
        const int y = ((lr / tile_rows) * nranks + q) * tile_rows + lr % tile_rows;
        if (y < H) frame[(size_t)y * W + x] = recv[i];
    }
}

// Interleaved square tiles (cost is concentrated on the terminator and the limb, so 64 x 64 tiles
// round-robin, not slabs).  Tile t = ty * tiles_x + tx belongs to rank t mod R and is that rank's slot t / R.
// resolve_tiles_kernel tone-maps an owned tile straight into its slot of the send buffer (no packing pass);
// after the all-gather untile_kernel writes every rank's slots back to frame order.
__global__ void resolve_tiles_kernel(const float4* __restrict__ accum, const uchar4* __restrict__ overlay, uchar4* __restrict__ send,
                                     int W, int H, int tl, int R, int rank, int slots, float exposure, float inv_gamma) {
    const int ts = 1 << tl, tiles_x = (W + ts - 1) >> tl, tiles_y = (H + ts - 1) >> tl;
    const size_t n = (size_t)slots << (2 * tl);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(i >> (2 * tl)), w = (int)(i & ((size_t)(ts * ts) - 1));
        const int t = k * R + rank;
        uchar4 px = make_uchar4(0, 0, 0, 0);
        if (t < tiles_x * tiles_y) {
            const int x = (t % tiles_x) * ts + (w & (ts - 1)), y = (t / tiles_x) * ts + (w >> tl);
            if (x < W && y < H) {
                const size_t j = (size_t)y * W + x;
                const float4 a = accum[j];
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
                    const uchar4 o = overlay[j];
                    const unsigned al = o.w, na = 255u - o.w;
                    c8[0] = (o.x * al + c8[0] * na + 127u) / 255u;
                    c8[1] = (o.y * al + c8[1] * na + 127u) / 255u;
                    c8[2] = (o.z * al + c8[2] * na + 127u) / 255u;
                }
                px = make_uchar4((unsigned char)c8[0], (unsigned char)c8[1], (unsigned char)c8[2], 255);
            }
        }
        send[i] = px;
    }
}
__global__ void untile_kernel(uchar4* __restrict__ frame, const uchar4* __restrict__ recv, int W, int H, int tl, int R, int slots) {
    const int ts = 1 << tl, tiles_x = (W + ts - 1) >> tl;
    const size_t n = (size_t)W * H;
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(j % W), y = (int)(j / W);
        const int t = (y >> tl) * tiles_x + (x >> tl);
        const size_t src = (((size_t)(t % R) * slots + (size_t)(t / R)) << (2 * tl)) + (size_t)((y & (ts - 1)) << tl) + (size_t)(x & (ts - 1));
        frame[j] = recv[src];
    }
}

}  // namespace

// point-to-point legs of the frame pipeline (api.cu: mrtx_frame_submit_to / mrtx_frame_recv)
int comm_send_bytes(mrtx_ctx* ctx, const void* buf_dev, size_t bytes, int peer, cudaStream_t st) {
    if (!ctx->nccl_comm) { mrtx_set_error("mrtx_comm_init has not been called"); return MRTX_ERR_STATE; }
    MRTX_NCCL(g_nccl.Send(buf_dev, bytes, NCCL_UINT8, peer, (nccl_comm_t)ctx->nccl_comm, st));
    return MRTX_OK;
}
int comm_recv_bytes(mrtx_ctx* ctx, void* buf_dev, size_t bytes, int peer, cudaStream_t st) {
    if (!ctx->nccl_comm) { mrtx_set_error("mrtx_comm_init has not been called"); return MRTX_ERR_STATE; }
    MRTX_NCCL(g_nccl.Recv(buf_dev, bytes, NCCL_UINT8, peer, (nccl_comm_t)ctx->nccl_comm, st));
    return MRTX_OK;
}

extern "C" {

int mrtx_comm_unique_id(const char* libnccl_path, uint8_t id128[128]) {
    MRTX_REQUIRE(id128, "null argument");
    int rc = load_nccl(libnccl_path);
    if (rc) return rc;
    nccl_uid uid;
    MRTX_NCCL(g_nccl.GetUniqueId(&uid));
    memcpy(id128, uid.internal, 128);
    return MRTX_OK;
}

int mrtx_comm_init(mrtx_ctx* ctx, const char* libnccl_path, int nranks, int rank, const uint8_t id128[128]) {
    MRTX_CTX(ctx);
    MRTX_REQUIRE(id128 && nranks >= 1 && rank >= 0 && rank < nranks, "bad communicator arguments");
    int rc = load_nccl(libnccl_path);
    if (rc) return rc;
    if (ctx->nccl_comm) mrtx_comm_destroy(ctx);
    nccl_uid uid;
    memcpy(uid.internal, id128, 128);
    nccl_comm_t comm = nullptr;
    MRTX_NCCL(g_nccl.CommInitRank(&comm, nranks, uid, rank));
    ctx->nccl_comm = comm;
    ctx->nranks = nranks;
    ctx->rank = rank;
    return MRTX_OK;
}

int mrtx_comm_destroy(mrtx_ctx* ctx) {
    if (!ctx || !ctx->nccl_comm) return MRTX_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->comm_stream) cudaStreamSynchronize(ctx->comm_stream);
    g_nccl.CommDestroy((nccl_comm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    ctx->nranks = 0;
    return MRTX_OK;
}

int mrtx_allreduce_accum(mrtx_ctx* ctx) {
    MRTX_CTX(ctx);
    if (!ctx->nccl_comm) { mrtx_set_error("mrtx_comm_init has not been called"); return MRTX_ERR_STATE; }
    if (!ctx->accum) { mrtx_set_error("frame buffer not allocated"); return MRTX_ERR_STATE; }
    const size_t count = (size_t)ctx->width * ctx->height * 4;
    MRTX_NCCL(g_nccl.AllReduce(ctx->accum, ctx->accum, count, NCCL_FLOAT32, NCCL_SUM,
                               (nccl_comm_t)ctx->nccl_comm, ctx->stream));
    return MRTX_OK;
}

int mrtx_allgather_tiles(mrtx_ctx* ctx, int tile) {
    MRTX_CTX(ctx);
    if (!ctx->nccl_comm) { mrtx_set_error("mrtx_comm_init has not been called"); return MRTX_ERR_STATE; }
    if (!ctx->rgba8) { mrtx_set_error("frame buffer not allocated"); return MRTX_ERR_STATE; }
    int tl = 0;
    while ((1 << tl) < tile) ++tl;
    MRTX_REQUIRE(tile >= 8 && tile <= 1024 && (1 << tl) == tile, "tile side must be a power of two in 8..1024");
    if (ctx->tex[1].data)
        MRTX_REQUIRE(ctx->tex[1].W == ctx->width && ctx->tex[1].H == ctx->height, "frame_overlay does not match the frame size");
    const int W = ctx->width, H = ctx->height, R = ctx->nranks;
    const int tiles = ((W + tile - 1) >> tl) * ((H + tile - 1) >> tl);
    const int slots = (tiles + R - 1) / R;
    const size_t slot_px = (size_t)slots << (2 * tl);               // uchar4 per rank
    const size_t need = slot_px * (size_t)(R + 1) * sizeof(uchar4);
    if (!ctx->gather_buf || ctx->gather_bytes < need) {
        MRTX_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->gather_buf);
        ctx->gather_buf = nullptr; ctx->gather_bytes = 0;
        MRTX_CUDA(cudaMalloc(&ctx->gather_buf, need));
        ctx->gather_bytes = need;
    }
    uchar4* send = (uchar4*)ctx->gather_buf;
    uchar4* recv = send + slot_px;
    const int blocks = ctx->sm_count * 8;
    resolve_tiles_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->accum, ctx->tex[1].data, send, W, H, tl, R, ctx->rank, slots,
                                                         ctx->sp.exposure, ctx->sp.inv_gamma);
    MRTX_CUDA(cudaGetLastError());
    MRTX_NCCL(g_nccl.AllGather(send, recv, slot_px * sizeof(uchar4), NCCL_UINT8, (nccl_comm_t)ctx->nccl_comm, ctx->stream));
    untile_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->rgba8, recv, W, H, tl, R, slots);
    MRTX_CUDA(cudaGetLastError());
    return MRTX_OK;
}

int mrtx_allgather_rows(mrtx_ctx* ctx, int tile_rows) {
    MRTX_CTX(ctx);
    if (!ctx->nccl_comm) { mrtx_set_error("mrtx_comm_init has not been called"); return MRTX_ERR_STATE; }
    if (!ctx->rgba8) { mrtx_set_error("frame buffer not allocated"); return MRTX_ERR_STATE; }
    MRTX_REQUIRE(tile_rows >= 1, "tile_rows must be >= 1");
    const int W = ctx->width, H = ctx->height, R = ctx->nranks;
    const int tiles = (H + tile_rows - 1) / tile_rows;
    const int per_rank = (tiles + R - 1) / R;
    const size_t slot = (size_t)per_rank * tile_rows * W;             // uchar4 per rank
    const size_t need = slot * (size_t)(R + 1) * sizeof(uchar4);
    // gather_buf = [send slot][R receive slots]; reallocated only when the frame grows
    if (!ctx->gather_buf || ctx->gather_bytes < need) {
        MRTX_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->gather_buf);
        ctx->gather_buf = nullptr; ctx->gather_bytes = 0;
        MRTX_CUDA(cudaMalloc(&ctx->gather_buf, need));
        ctx->gather_bytes = need;
    }
    uchar4* send = (uchar4*)ctx->gather_buf;
    uchar4* recv = send + slot;
    const int blocks = ctx->sm_count * 8;
    pack_rows_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->rgba8, send, W, H, tile_rows, R, ctx->rank, per_rank);
    MRTX_CUDA(cudaGetLastError());
    MRTX_NCCL(g_nccl.AllGather(send, recv, slot * sizeof(uchar4), NCCL_UINT8, (nccl_comm_t)ctx->nccl_comm, ctx->stream));
    unpack_rows_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->rgba8, recv, W, H, tile_rows, R, per_rank);
    MRTX_CUDA(cudaGetLastError());
    return MRTX_OK;
}

}  // extern "C"
