// Internal helpers shared by the alona_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <limits.h>
#include <string>

#include "../../include/alona_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "alona_b200 kernels are written for sm_100a only"
#endif

// ---------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------
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
    int *d_counter = nullptr;      // device scratch: "last block done" tickets (64 ints, kept 0)
};

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
