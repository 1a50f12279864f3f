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


// ---- frame delivery through peer memory (mailboxes) -----------------------------------------------------------------------
// See mrtx_ctx::p2p_box.  Layout of a mailbox: u32 ready[MAX_RANKS][2] at 0, u32 free[MAX_RANKS][2] at 1024, frame slots
// [src][2] from 4096.  ready[s][q] = k: rank s has written its (2 k - 2 + q)-th frame for this rank into slot [s][q];
// free[d][q] = k (in the SENDER's mailbox): rank d has copied the k-th frame it received through its slot [me][q] out.
namespace {

constexpr size_t P2P_READY_OFF = 0, P2P_FREE_OFF = 1024, P2P_SLOTS_OFF = 4096;
constexpr unsigned P2P_SEQ_N = 1u << 16;                    // frames per (sender, slot): 131 072 frames per pair of ranks
typedef int (*wait_value32_fn)(cudaStream_t, unsigned long long, unsigned, unsigned);

inline unsigned* box_word(void* box, size_t off, int r, int q) { return (unsigned*)((char*)box + off) + 2 * r + q; }
inline char* box_slot(void* box, size_t stride, int src, int q) { return (char*)box + P2P_SLOTS_OFF + (size_t)(2 * src + q) * stride; }

int p2p_wait(mrtx_ctx* ctx, cudaStream_t st, unsigned* word, unsigned value) {
    const int rc = ((wait_value32_fn)ctx->p2p_wait_fn)(st, (unsigned long long)(uintptr_t)word, value, ctx->p2p_wait_flags);
    if (rc != 0) { mrtx_set_error("cuStreamWaitValue32 failed (CUresult %d)", rc); return MRTX_ERR_CUDA; }
    return MRTX_OK;
}

}  // namespace

void p2p_release(mrtx_ctx* ctx) {
    for (int r = 0; r < MRTX_P2P_MAX_RANKS; ++r) {
        if (ctx->p2p_peer[r] && r != ctx->rank) cudaIpcCloseMemHandle(ctx->p2p_peer[r]);
        ctx->p2p_peer[r] = nullptr;
    }
    cudaFree(ctx->p2p_box); cudaFree(ctx->p2p_seq);
    ctx->p2p_box = nullptr; ctx->p2p_seq = nullptr; ctx->p2p_on = 0;
}

int mrtx_p2p_open(mrtx_ctx* ctx, int nranks, int rank, size_t slot_bytes, uint8_t handle_out[64]) {
    MRTX_CTX(ctx);
    MRTX_REQUIRE(handle_out && slot_bytes > 0, "null argument");
    MRTX_REQUIRE(nranks >= 1 && nranks <= MRTX_P2P_MAX_RANKS && rank >= 0 && rank < nranks, "bad rank %d of %d", rank, nranks);
    if (ctx->nccl_comm) MRTX_REQUIRE(nranks == ctx->nranks && rank == ctx->rank, "rank %d of %d differs from the communicator's %d of %d", rank, nranks, ctx->rank, ctx->nranks);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle crosses the ABI as 64 opaque bytes");
    MRTX_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->comm_stream) MRTX_CUDA(cudaStreamSynchronize(ctx->comm_stream));
    p2p_release(ctx);
    // the stream wait on a memory word: a driver entry point, reached through the runtime (no link against libcuda)
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    MRTX_CUDA(cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &qr));
    if (!fn || qr != cudaDriverEntryPointSuccess) { mrtx_set_error("cuStreamWaitValue32 is not available in this driver"); return MRTX_ERR_CUDA; }
    ctx->p2p_wait_fn = fn;
    int dev = 0, can_flush = 0;
    MRTX_CUDA(cudaGetDevice(&dev));
    (void)cudaDeviceGetAttribute(&can_flush, cudaDevAttrCanFlushRemoteWrites, dev);
    (void)cudaGetLastError();
    ctx->p2p_wait_flags = 0x0u /* GEQ */ | (can_flush ? (1u << 30) /* FLUSH */ : 0u);
    ctx->p2p_stride = (slot_bytes + 255) & ~(size_t)255;
    ctx->p2p_slot_bytes = slot_bytes;
    const size_t total = P2P_SLOTS_OFF + (size_t)2 * nranks * ctx->p2p_stride;
    MRTX_CUDA(cudaMalloc(&ctx->p2p_box, total));
    MRTX_CUDA(cudaMemset(ctx->p2p_box, 0, P2P_SLOTS_OFF));
    MRTX_CUDA(cudaMalloc(&ctx->p2p_seq, P2P_SEQ_N * sizeof(unsigned)));
    {
        unsigned* h = (unsigned*)malloc(P2P_SEQ_N * sizeof(unsigned));
        if (!h) { mrtx_set_error("out of host memory"); return MRTX_ERR_INVALID; }
        for (unsigned i = 0; i < P2P_SEQ_N; ++i) h[i] = i;
        const cudaError_t e = cudaMemcpy(ctx->p2p_seq, h, P2P_SEQ_N * sizeof(unsigned), cudaMemcpyHostToDevice);
        free(h);
        MRTX_CUDA(e);
    }
    cudaIpcMemHandle_t h;
    MRTX_CUDA(cudaIpcGetMemHandle(&h, ctx->p2p_box));
    memcpy(handle_out, &h, 64);
    ctx->nranks = nranks; ctx->rank = rank;
    for (int r = 0; r < MRTX_P2P_MAX_RANKS; ++r) { ctx->p2p_sent[r] = 0; ctx->p2p_rcvd[r] = 0; }
    return MRTX_OK;
}

int mrtx_p2p_connect(mrtx_ctx* ctx, const uint8_t* handles) {
    MRTX_CTX(ctx);
    MRTX_REQUIRE(handles, "null argument");
    if (!ctx->p2p_box) { mrtx_set_error("mrtx_p2p_open has not been called"); return MRTX_ERR_STATE; }
    for (int r = 0; r < ctx->nranks; ++r) {
        if (r == ctx->rank) { ctx->p2p_peer[r] = ctx->p2p_box; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)64 * r, 64);
        MRTX_CUDA(cudaIpcOpenMemHandle(&ctx->p2p_peer[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    ctx->p2p_on = 1;
    return MRTX_OK;
}

int mrtx_p2p_close(mrtx_ctx* ctx) {
    MRTX_CTX(ctx);
    MRTX_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->comm_stream) MRTX_CUDA(cudaStreamSynchronize(ctx->comm_stream));
    p2p_release(ctx);
    return MRTX_OK;
}

// queue on `st`: wait until the consumer has emptied the slot, copy the frame into it, publish its sequence number
int p2p_send_frame(mrtx_ctx* ctx, const void* frame_dev, size_t bytes, int dst, cudaStream_t st) {
    MRTX_REQUIRE(ctx->p2p_on && bytes <= ctx->p2p_slot_bytes, "frame of %zu bytes does not fit the %zu-byte mailbox slots", bytes, ctx->p2p_slot_bytes);
    const unsigned j = ctx->p2p_sent[dst], seq = j / 2 + 1;
    const int q = (int)(j & 1u);
    MRTX_REQUIRE(seq < P2P_SEQ_N, "sequence numbers exhausted");
    int rc = MRTX_OK;
    if (seq > 1 && (rc = p2p_wait(ctx, st, box_word(ctx->p2p_box, P2P_FREE_OFF, dst, q), seq - 1))) return rc;
    void* peer = ctx->p2p_peer[dst];
    MRTX_CUDA(cudaMemcpyAsync(box_slot(peer, ctx->p2p_stride, ctx->rank, q), frame_dev, bytes, cudaMemcpyDeviceToDevice, st));
    MRTX_CUDA(cudaMemcpyAsync(box_word(peer, P2P_READY_OFF, ctx->rank, q), ctx->p2p_seq + seq, sizeof(unsigned), cudaMemcpyDeviceToDevice, st));
    ctx->p2p_sent[dst] = j + 1;
    return MRTX_OK;
}

// queue on `st`: wait for the sender's sequence number, copy the slot to pinned host memory, hand the slot back
int p2p_recv_frame(mrtx_ctx* ctx, void* out_pinned, size_t bytes, int src, cudaStream_t st) {
    MRTX_REQUIRE(ctx->p2p_on && bytes <= ctx->p2p_slot_bytes, "frame of %zu bytes does not fit the %zu-byte mailbox slots", bytes, ctx->p2p_slot_bytes);
    const unsigned j = ctx->p2p_rcvd[src], seq = j / 2 + 1;
    const int q = (int)(j & 1u);
    MRTX_REQUIRE(seq < P2P_SEQ_N, "sequence numbers exhausted");
    const int rc = p2p_wait(ctx, st, box_word(ctx->p2p_box, P2P_READY_OFF, src, q), seq);
    if (rc) return rc;
    MRTX_CUDA(cudaMemcpyAsync(out_pinned, box_slot(ctx->p2p_box, ctx->p2p_stride, src, q), bytes, cudaMemcpyDeviceToHost, st));
    MRTX_CUDA(cudaMemcpyAsync(box_word(ctx->p2p_peer[src], P2P_FREE_OFF, ctx->rank, q), ctx->p2p_seq + seq, sizeof(unsigned), cudaMemcpyDeviceToDevice, st));
    ctx->p2p_rcvd[src] = j + 1;
    return MRTX_OK;
}
