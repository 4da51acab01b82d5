// Context, error reporting, NCCL plumbing (dlopen) and the device scan utility.
#include "ab_common.cuh"
#include "ab_peer.cuh"

#include <dlfcn.h>
#include <limits.h>
#include <string.h>

// ---------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";

int ab_fail(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}

extern "C" const char *ab_last_error(void) { return g_err; }

// digest of the sources this library was built from (alona_b200/build.py embeds and checks it)
#ifndef AB_SOURCE_DIGEST
#define AB_SOURCE_DIGEST "unknown"
#endif
extern "C" const char *ab_source_digest(void) {
    static const char tag[] = "AB_SOURCE_DIGEST=" AB_SOURCE_DIGEST;
    return tag + sizeof("AB_SOURCE_DIGEST=") - 1;
}
extern "C" int ab_version(void) { return AB_VERSION; }

// ---------------------------------------------------------------------------------------
// NCCL through dlopen: the library has no link-time dependency on NCCL, so it loads on a
// box without it (single-GPU use) and binds to the torch-bundled libnccl when sharded.
// ---------------------------------------------------------------------------------------
typedef struct { char internal[128]; } ab_nccl_uid;
struct ab_nccl_api {
    void *handle;
    int (*GetUniqueId)(ab_nccl_uid *);
    int (*CommInitRank)(void **, int, ab_nccl_uid, int);
    int (*CommDestroy)(void *);
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t);
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t);
    const char *(*GetErrorString)(int);
};
enum { AB_NCCL_INT8 = 0, AB_NCCL_INT64 = 4, AB_NCCL_F64 = 8, AB_NCCL_SUM = 0 };

static ab_nccl_api *ab_nccl_open(const char *path) {
    static ab_nccl_api api;
    static bool loaded = false;
    if (loaded) return &api;
    void *h = dlopen(path && path[0] ? path : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { ab_fail("dlopen(%s) failed: %s", path ? path : "libnccl.so.2", dlerror()); return nullptr; }
    api.handle = h;
#define AB_SYM(field, name)                                                    \
    *(void **)(&api.field) = dlsym(h, name);                                   \
    if (!api.field) { ab_fail("dlsym(%s) failed", name); return nullptr; }
    AB_SYM(GetUniqueId, "ncclGetUniqueId")
    AB_SYM(CommInitRank, "ncclCommInitRank")
    AB_SYM(CommDestroy, "ncclCommDestroy")
    AB_SYM(AllReduce, "ncclAllReduce")
    AB_SYM(AllGather, "ncclAllGather")
    AB_SYM(GetErrorString, "ncclGetErrorString")
#undef AB_SYM
    loaded = true;
    return &api;
}

#define AB_NCCL(ctx, call)                                                                 \
    do {                                                                                   \
        int _r = (call);                                                                   \
        if (_r != 0) return ab_fail("%s:%d: NCCL %s -> %s", __FILE__, __LINE__, #call,     \
                                    (ctx)->nccl->GetErrorString(_r));                      \
    } while (0)

extern "C" int ab_nccl_unique_id(const char *libnccl_path, void *h_id128) {
    ab_nccl_api *api = ab_nccl_open(libnccl_path);
    if (!api) return 1;
    ab_nccl_uid id;
    int r = api->GetUniqueId(&id);
    if (r != 0) return ab_fail("ncclGetUniqueId -> %s", api->GetErrorString(r));
    memcpy(h_id128, &id, 128);
    return 0;
}

extern "C" int ab_ctx_init_nccl(ab_ctx *ctx, const char *libnccl_path, const void *h_id128) {
    AB_REQUIRE(ctx, "ab_ctx_init_nccl: null ctx");
    if (ctx->world == 1) return 0;
    ab_nccl_api *api = ab_nccl_open(libnccl_path);
    if (!api) return 1;
    ctx->nccl = api;
    ab_nccl_uid id;
    memcpy(&id, h_id128, 128);
    AB_CUDA(cudaSetDevice(ctx->device));
    AB_NCCL(ctx, api->CommInitRank(&ctx->comm, ctx->world, id, ctx->rank));
    return 0;
}

extern "C" int ab_allreduce_f64(ab_ctx *ctx, double *buf, int64_t count, void *stream) {
    if (ctx->world == 1 || count == 0) return 0;
    AB_REQUIRE(ctx->comm, "ab_allreduce_f64: world>1 but NCCL not initialised");
    AB_NCCL(ctx, ctx->nccl->AllReduce(buf, buf, (size_t)count, AB_NCCL_F64, AB_NCCL_SUM, ctx->comm,
                                      (cudaStream_t)stream));
    return 0;
}

extern "C" int ab_allreduce_i64(ab_ctx *ctx, int64_t *buf, int64_t count, void *stream) {
    if (ctx->world == 1 || count == 0) return 0;
    AB_REQUIRE(ctx->comm, "ab_allreduce_i64: world>1 but NCCL not initialised");
    AB_NCCL(ctx, ctx->nccl->AllReduce(buf, buf, (size_t)count, AB_NCCL_INT64, AB_NCCL_SUM, ctx->comm,
                                      (cudaStream_t)stream));
    return 0;
}

extern "C" int ab_allgather_bytes(ab_ctx *ctx, const void *send, void *recv, int64_t bytes_per_rank,
                                  void *stream) {
    if (ctx->world == 1) {
        if (send != recv)
            AB_CUDA(cudaMemcpyAsync(recv, send, (size_t)bytes_per_rank, cudaMemcpyDeviceToDevice,
                                    (cudaStream_t)stream));
        return 0;
    }
    AB_REQUIRE(ctx->comm, "ab_allgather_bytes: world>1 but NCCL not initialised");
    AB_NCCL(ctx, ctx->nccl->AllGather(send, recv, (size_t)bytes_per_rank, AB_NCCL_INT8, ctx->comm,
                                      (cudaStream_t)stream));
    return 0;
}

// ---------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------
// peer-exchange buffers (ab_peer.cuh): allocation, IPC handles, mapping of the other ranks' buffers
// ---------------------------------------------------------------------------------------
extern "C" size_t ab_peer_handle_bytes(void) { return sizeof(cudaIpcMemHandle_t); }

extern "C" int ab_peer_alloc(ab_ctx *ctx, void *h_handle) {
    AB_REQUIRE(ctx && h_handle, "ab_peer_alloc: null argument");
    AB_REQUIRE(ctx->world <= AB_PEER_MAX, "ab_peer_alloc: world %d exceeds %d", ctx->world, AB_PEER_MAX);
    AB_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->peer_local) {
        AB_CUDA(cudaMalloc(&ctx->peer_local, AB_PEER_BYTES));
        AB_CUDA(cudaMemset(ctx->peer_local, 0, AB_PEER_BYTES));
        AB_CUDA(cudaDeviceSynchronize());
    }
    cudaIpcMemHandle_t h;
    AB_CUDA(cudaIpcGetMemHandle(&h, ctx->peer_local));
    memcpy(h_handle, &h, sizeof(h));
    return 0;
}

extern "C" int ab_peer_open(ab_ctx *ctx, const void *h_handles) {
    AB_REQUIRE(ctx && h_handles, "ab_peer_open: null argument");
    AB_REQUIRE(ctx->peer_local, "ab_peer_open: call ab_peer_alloc first");
    AB_CUDA(cudaSetDevice(ctx->device));
    for (int q = 0; q < ctx->world; ++q) {
        if (q == ctx->rank) { ctx->peer_buf[q] = ctx->peer_local; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)h_handles + (size_t)q * sizeof(h), sizeof(h));
        AB_CUDA(cudaIpcOpenMemHandle(&ctx->peer_buf[q], h, cudaIpcMemLazyEnablePeerAccess));
    }
    AbPeer h{};
    h.world = ctx->world;
    h.rank = ctx->rank;
    for (int q = 0; q < ctx->world; ++q) h.buf[q] = (char *)ctx->peer_buf[q];
    if (!ctx->d_peer) AB_CUDA(cudaMalloc(&ctx->d_peer, sizeof(AbPeer)));
    AB_CUDA(cudaMemcpy(ctx->d_peer, &h, sizeof(AbPeer), cudaMemcpyHostToDevice));
    ctx->peer_ready = true;
    return 0;
}

extern "C" int ab_peer_enabled(ab_ctx *ctx) { return ctx && ctx->peer_ready ? 1 : 0; }

extern "C" int ab_ctx_create(ab_ctx **out, int device, int rank, int world) {
    AB_REQUIRE(out, "ab_ctx_create: null out");
    AB_REQUIRE(world >= 1 && rank >= 0 && rank < world, "ab_ctx_create: bad rank/world %d/%d", rank, world);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return ab_fail("ab_ctx_create: no CUDA device (%s); alona_b200 has no CPU fallback",
                       cudaGetErrorString(e));
    AB_REQUIRE(device >= 0 && device < ndev, "ab_ctx_create: device %d out of range (%d)", device, ndev);
    cudaDeviceProp prop;
    AB_CUDA(cudaGetDeviceProperties(&prop, device));
    AB_REQUIRE(prop.major == 10, "ab_ctx_create: device %d is sm_%d%d; kernels are built for sm_100a only",
               device, prop.major, prop.minor);
    AB_CUDA(cudaSetDevice(device));
    ab_ctx *c = new ab_ctx();
    c->device = device;
    c->rank = rank;
    c->world = world;
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    AB_CUDA(cudaMalloc(&c->d_counter, 64 * sizeof(int)));
    AB_CUDA(cudaMemset(c->d_counter, 0, 64 * sizeof(int)));
    *out = c;
    return 0;
}

extern "C" int ab_ctx_destroy(ab_ctx *ctx) {
    if (!ctx) return 0;
    if (ctx->comm && ctx->nccl) ctx->nccl->CommDestroy(ctx->comm);
    for (int q = 0; q < ctx->world && q < AB_PEER_MAX; ++q)
        if (ctx->peer_ready && q != ctx->rank && ctx->peer_buf[q]) cudaIpcCloseMemHandle(ctx->peer_buf[q]);
    if (ctx->peer_local) cudaFree(ctx->peer_local);
    if (ctx->d_peer) cudaFree(ctx->d_peer);
    if (ctx->d_counter) cudaFree(ctx->d_counter);
    for (int i = 0; i < AB_PROF_SLOTS; ++i) {
        for (cudaEvent_t e : ctx->prof[i].begin) cudaEventDestroy(e);
        for (cudaEvent_t e : ctx->prof[i].end) cudaEventDestroy(e);
    }
    for (cudaEvent_t e : ctx->prof_pool) cudaEventDestroy(e);
    delete ctx;
    return 0;
}

extern "C" int64_t ab_launch_count(ab_ctx *ctx) { return ctx ? ctx->launches : 0; }

// ---------------------------------------------------------------------------------------
// CUDA-event profiling slots
// ---------------------------------------------------------------------------------------
static const char *const g_prof_names[AB_PROF_SLOTS] = {
    "norm_moments", "hvg_extract", "irlb_av", "irlb_atw", "irlb_panel", "irlb_small", "knn_prep", "knn_main",
    "knn_rerank", "snn_build", "snn_count", "snn_fill", "irlb_total", "knn_exact", "irlb_ritz", "irlb_svd"};

static bool prof_active(ab_ctx *ctx, int slot) {
    if (!ctx || ctx->prof_level <= 0 || slot < 0 || slot >= AB_PROF_SLOTS) return false;
    if (ctx->prof_level < 2 && ((AB_PROF_FINE_MASK >> slot) & 1)) return false;
    return true;
}
static cudaEvent_t prof_event(ab_ctx *ctx) {
    if (!ctx->prof_pool.empty()) {
        cudaEvent_t e = ctx->prof_pool.back();
        ctx->prof_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
void ab_prof_begin(ab_ctx *ctx, int slot, cudaStream_t st) {
    if (!prof_active(ctx, slot)) return;
    ab_prof_slot &s = ctx->prof[slot];
    if (s.open) return;
    cudaEvent_t e = prof_event(ctx);
    if (!e) return;
    cudaEventRecord(e, st);
    s.begin.push_back(e);
    s.open = true;
}
void ab_prof_end(ab_ctx *ctx, int slot, cudaStream_t st) {
    if (!prof_active(ctx, slot)) return;
    ab_prof_slot &s = ctx->prof[slot];
    if (!s.open) return;
    cudaEvent_t e = prof_event(ctx);
    if (!e) { ctx->prof_pool.push_back(s.begin.back()); s.begin.pop_back(); s.open = false; return; }
    cudaEventRecord(e, st);
    s.end.push_back(e);
    s.open = false;
}

extern "C" int ab_prof_set_level(ab_ctx *ctx, int level) {
    AB_REQUIRE(ctx, "ab_prof_set_level: null ctx");
    AB_REQUIRE(level >= 0 && level <= 2, "ab_prof_set_level: level must be 0, 1 or 2");
    ctx->prof_level = level;
    return 0;
}

extern "C" const char *ab_prof_slot_name(int slot) {
    return (slot >= 0 && slot < AB_PROF_SLOTS) ? g_prof_names[slot] : "";
}

extern "C" int ab_prof_read(ab_ctx *ctx, int slot, double *h_ms_total, int64_t *h_count) {
    AB_REQUIRE(ctx, "ab_prof_read: null ctx");
    AB_REQUIRE(slot >= 0 && slot < AB_PROF_SLOTS, "ab_prof_read: bad slot %d", slot);
    ab_prof_slot &s = ctx->prof[slot];
    double total = 0.0;
    const size_t n = s.end.size();
    for (size_t i = 0; i < n; ++i) {
        AB_CUDA(cudaEventSynchronize(s.end[i]));
        float ms = 0.f;
        AB_CUDA(cudaEventElapsedTime(&ms, s.begin[i], s.end[i]));
        total += ms;
        ctx->prof_pool.push_back(s.begin[i]);
        ctx->prof_pool.push_back(s.end[i]);
    }
    if (s.open) { ctx->prof_pool.push_back(s.begin.back()); s.open = false; }
    s.begin.clear();
    s.end.clear();
    if (h_ms_total) *h_ms_total = total;
    if (h_count) *h_count = (int64_t)n;
    return 0;
}

// ---------------------------------------------------------------------------------------
// exclusive scan: block partials -> scan of partials (single block) -> block scan + offset
// ---------------------------------------------------------------------------------------
static constexpr int SCAN_THREADS = 256;
static constexpr int SCAN_ITEMS = 8;                  // per thread
static constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ long long block_exclusive_scan(long long v, long long *total, long long *smem) {
    // smem: 32 entries.  returns exclusive prefix of v over the block.
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        long long t = __shfl_up_sync(AB_FULL, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        long long w = lane < (blockDim.x >> 5) ? smem[lane] : 0;
        long long winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            long long t = __shfl_up_sync(AB_FULL, winc, d);
            if (lane >= d) winc += t;
        }
        smem[lane] = winc - w;                         // exclusive warp offsets
        if (lane == 31) smem[32] = winc;
    }
    __syncthreads();
    long long res = inc - v + smem[warp];
    *total = smem[32];
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_partials(const int64_t *in, int64_t n, int64_t *partial) {
    __shared__ long long sm[33];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    long long acc = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t p = base + (int64_t)threadIdx.x * SCAN_ITEMS + i;
        if (p < n) acc += in[p];
    }
    long long total;
    block_exclusive_scan(acc, &total, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) scan_block_sums(int64_t *partial, int64_t nblocks, int64_t *grand) {
    __shared__ long long sm[33];
    long long carry = 0;
    for (int64_t base = 0; base < nblocks; base += 1024) {
        int64_t p = base + threadIdx.x;
        long long v = p < nblocks ? partial[p] : 0;
        long long total;
        long long ex = block_exclusive_scan(v, &total, sm);
        if (p < nblocks) partial[p] = ex + carry;
        carry += total;
    }
    if (threadIdx.x == 0) *grand = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_final(const int64_t *in, int64_t *out, int64_t n,
                                                           const int64_t *partial, const int64_t *grand) {
    __shared__ long long sm[33];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    long long v[SCAN_ITEMS];
    long long acc = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t p = base + (int64_t)threadIdx.x * SCAN_ITEMS + i;
        v[i] = p < n ? in[p] : 0;
        acc += v[i];
    }
    long long total;
    long long ex = block_exclusive_scan(acc, &total, sm) + partial[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t p = base + (int64_t)threadIdx.x * SCAN_ITEMS + i;
        if (p < n) out[p] = ex;
        ex += v[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = *grand;
}

size_t ab_scan_ws_bytes(int64_t n) {
    int64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    return (size_t)(nb + 2) * sizeof(int64_t) + 256;
}

int ab_exclusive_scan_i64(ab_ctx *ctx, const int64_t *in, int64_t *out, int64_t n, void *ws,
                          size_t ws_bytes, cudaStream_t stream) {
    AB_REQUIRE(ws_bytes >= ab_scan_ws_bytes(n), "scan: workspace too small");
    if (n == 0) {
        AB_CUDA(cudaMemsetAsync(out, 0, sizeof(int64_t), stream));
        return 0;
    }
    int64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    int64_t *partial = (int64_t *)ws;
    int64_t *grand = partial + nb;
    scan_partials<<<(unsigned)nb, SCAN_THREADS, 0, stream>>>(in, n, partial);
    AB_LAUNCH_CHECK(ctx);
    scan_block_sums<<<1, 1024, 0, stream>>>(partial, nb, grand);
    AB_LAUNCH_CHECK(ctx);
    scan_final<<<(unsigned)nb, SCAN_THREADS, 0, stream>>>(in, out, n, partial, grand);
    AB_LAUNCH_CHECK(ctx);
    return 0;
}
