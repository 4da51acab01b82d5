// Internal helpers shared by the alona_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <limits.h>
#include <string>
#include <vector>

#include "../../include/alona_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "alona_b200 kernels are written for sm_100a only"
#endif

// ---------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------
// profiling slots: CUDA-event pairs recorded on the launching stream around the named kernels
// (ab_prof_set_level: 0 = off, 1 = one pair per stage-level kernel, 2 = every launch incl. the
// many small IRLBA launches).  bench.py reads them for the roofline fields.
enum {
    AB_PROF_NORM = 0, AB_PROF_HVG_EXTRACT = 1, AB_PROF_IRLB_AV = 2, AB_PROF_IRLB_ATW = 3, AB_PROF_IRLB_PANEL = 4,
    AB_PROF_IRLB_SMALL = 5, AB_PROF_KNN_PREP = 6, AB_PROF_KNN_MAIN = 7, AB_PROF_KNN_RERANK = 8,
    AB_PROF_SNN_BUILD = 9, AB_PROF_SNN_COUNT = 10, AB_PROF_SNN_FILL = 11, AB_PROF_IRLB_TOTAL = 12,
    AB_PROF_KNN_EXACT = 13, AB_PROF_IRLB_RITZ = 14, AB_PROF_IRLB_SVD = 15   // AB_PROF_SLOTS (16): public header
};
static const int AB_PROF_FINE_MASK = (1 << AB_PROF_IRLB_AV) | (1 << AB_PROF_IRLB_ATW) | (1 << AB_PROF_IRLB_PANEL) |
                                     (1 << AB_PROF_IRLB_SMALL) | (1 << AB_PROF_IRLB_RITZ) | (1 << AB_PROF_IRLB_SVD);

struct ab_prof_slot {
    std::vector<cudaEvent_t> begin, end;     // completed or in-flight pairs
    bool open = false;
};

struct ab_nccl_api;   // resolved with dlopen, see ab_api.cu
struct ab_ctx {
    int device = 0;
    int rank = 0;
    int world = 1;
    int sm_count = 148;
    size_t smem_optin = 0;
    int64_t launches = 0;
    ab_nccl_api *nccl = nullptr;
    void *comm = nullptr;          // ncclComm_t
    void *peer_local = nullptr;    // this rank's peer-exchange buffer (ab_peer.cuh), cudaMalloc'ed
    void *peer_buf[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // all ranks' buffers, mapped
    void *d_peer = nullptr;        // device copy of the AbPeer table the kernels read
    bool peer_ready = false;       // ab_peer_open succeeded on every rank
    uint32_t peer_seq[4] = {0, 0, 0, 0};   // per-channel instance counters (identical on all ranks)
    int *d_counter = nullptr;      // device scratch: "last block done" tickets (64 ints, kept 0)
    int prof_level = 0;
    ab_prof_slot prof[AB_PROF_SLOTS];
    std::vector<cudaEvent_t> prof_pool;
};

// Grid of a grid-stride kernel: whole resident waves.  A fixed "N CTAs per SM" that exceeds what registers / shared memory
// let reside leaves a partial last wave running at a fraction of the machine (measured on the SNN table kernels: 6 per
// SM launched, 5 resident, the leftover 148 CTAs as long as the first 740).
template <class K>
static inline unsigned ab_wave_grid(const ab_ctx *ctx, K kernel, int threads, size_t smem, int64_t useful_blocks) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    int64_t g = (int64_t)ctx->sm_count * per_sm;
    if (g > useful_blocks) g = useful_blocks;
    return (unsigned)(g < 1 ? 1 : g);
}

void ab_prof_begin(ab_ctx *ctx, int slot, cudaStream_t st);
void ab_prof_end(ab_ctx *ctx, int slot, cudaStream_t st);

int ab_fail(const char *fmt, ...);                 // sets the thread-local message, returns 1
const char *ab_cuda_err_name(cudaError_t e);

#define AB_CUDA(call)                                                                      \
    do {                                                                                   \
        cudaError_t _e = (call);                                                           \
        if (_e != cudaSuccess)                                                             \
            return ab_fail("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
    } while (0)

#define AB_LAUNCH_CHECK(ctx)                                                               \
    do {                                                                                   \
        (ctx)->launches++;                                                                 \
        cudaError_t _e = cudaGetLastError();                                               \
        if (_e != cudaSuccess)                                                             \
            return ab_fail("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
    } while (0)

#define AB_REQUIRE(cond, ...)                                                              \
    do {                                                                                   \
        if (!(cond)) return ab_fail(__VA_ARGS__);                                          \
    } while (0)

static inline size_t ab_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// bump allocator over the caller's workspace
struct ab_arena {
    char *base;
    size_t cap;
    size_t off = 0;
    bool ok = true;
    ab_arena(void *p, size_t bytes) : base((char *)p), cap(bytes) {}
    template <typename T>
    T *take(size_t count) {
        size_t a = ab_align_up(off, 256);
        size_t need = count * sizeof(T);
        if (a + need > cap) { ok = false; return nullptr; }
        off = a + need;
        return (T *)(base + a);
    }
};
static inline size_t ab_arena_need(size_t cur, size_t bytes) { return ab_align_up(cur, 256) + bytes; }

// ---------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------
#ifdef __CUDACC__
#define AB_FULL 0xffffffffu

__device__ __forceinline__ double ab_shfl_f64(double v, int src) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(AB_FULL, lo, src);
    hi = __shfl_sync(AB_FULL, hi, src);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double ab_shfl_xor_f64(double v, int m) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(AB_FULL, lo, m);
    hi = __shfl_xor_sync(AB_FULL, hi, m);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double ab_shfl_down_f64(double v, int d) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_down_sync(AB_FULL, lo, d);
    hi = __shfl_down_sync(AB_FULL, hi, d);
    return __hiloint2double(hi, lo);
}
// fixed-order (deterministic) butterfly sum: every lane ends with the same total
__device__ __forceinline__ double ab_warp_sum_f64(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += ab_shfl_xor_f64(v, m);
    return v;
}
__device__ __forceinline__ long long ab_warp_sum_i64(long long v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(AB_FULL, v, m);
    return v;
}
__device__ __forceinline__ int ab_warp_sum_i32(int v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(AB_FULL, v, m);
    return v;
}

// The reference's arithmetic for one normalised value (alona/cell.py:242-243), FP64, no FMA:
//   x = log2( fl( fl(c / S) * scale ) + 1 )
__device__ __forceinline__ double ab_norm_value(double c, double lib, double scale) {
    double q = __ddiv_rn(c, lib);
    q = __dmul_rn(q, scale);
    q = __dadd_rn(q, 1.0);
    return log2(q);
}
#endif

// ---------------------------------------------------------------------------------------
// exclusive scan (int64 in -> int64 out[n+1]) used by the count->fill protocols
// ---------------------------------------------------------------------------------------
size_t ab_scan_ws_bytes(int64_t n);
// out may alias in.  out has n+1 entries, out[n] = total.
int ab_exclusive_scan_i64(ab_ctx *ctx, const int64_t *in, int64_t *out, int64_t n, void *ws,
                          size_t ws_bytes, cudaStream_t stream);
