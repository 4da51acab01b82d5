// K1: fused library-size normalisation + log2 + per-gene moments over cell-major CSR counts.
// K3: extraction of the HVG-restricted normalised matrix A_H.
//
// Reference arithmetic (alona/cell.py:241-243):  S_c = sum_g c_gc ;  x = log2(c/S_c*1e4 + 1)
// and the per-gene reductions hvg_seurat needs (alona/hvg.py:148,157): sum x, sum x^2.
//
// Design (HBM-bound integer/FP64 streaming, no tensor cores):
//   * one warp per cell row, 128-bit coalesced loads of (col,val);
//   * counts are small integers, so the expensive FP64 log2 is evaluated once per
//     (row, count value): lane l tabulates x(l+1); every nonzero with c <= 32 fetches its
//     value with two warp shuffles.  That is bit-identical to evaluating the expression per
//     nonzero and removes ~98 % of the FP64 transcendental work;
//   * per-gene accumulation uses FP64 reductions in L2 (RED.ADD.F64): the 2*G accumulators
//     (<= 0.5 MB) stay L2-resident.
#include "ab_common.cuh"

static constexpr int NORM_WARPS = 8;

// Sum of one row's counts by the whole warp (exact, int64).
__device__ __forceinline__ long long row_sum_counts(const int32_t *__restrict__ val, int64_t a, int64_t b, int lane) {
    long long acc = 0;
    // head to 16-byte alignment
    int64_t p = a;
    int64_t a4 = a + ((4 - (int)((reinterpret_cast<uintptr_t>(val + a) >> 2) & 3)) & 3);
    if (a4 > b) a4 = b;
    if (p + lane < a4) acc += val[p + lane];
    p = a4;
    int64_t nvec = (b - p) >> 2;
    const int4 *v4 = reinterpret_cast<const int4 *>(val + p);
    for (int64_t i = lane; i < nvec; i += 32) {
        int4 t = __ldg(v4 + i);
        acc += (long long)t.x + t.y + t.z + t.w;
    }
    p += nvec << 2;
    if (p + lane < b) acc += val[p + lane];
    return ab_warp_sum_i64(acc);
}

__global__ void __launch_bounds__(NORM_WARPS * 32)
norm_moments_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                    const int32_t *__restrict__ val, int64_t n_local, int32_t n_genes, double scale,
                    int64_t *__restrict__ libsize, double *__restrict__ gsum, double *__restrict__ gsumsq) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = (int64_t)blockIdx.x * NORM_WARPS + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * NORM_WARPS;
    for (int64_t row = warp_global; row < n_local; row += nwarps) {
        const int64_t a = rowptr[row], b = rowptr[row + 1];
        const long long S = row_sum_counts(val, a, b, lane);
        if (lane == 0) libsize[row] = S;
        const double lib = (double)S;
        // lane l holds x for count l+1
        const double tab = ab_norm_value((double)(lane + 1), lib, scale);
        for (int64_t p0 = a + lane; p0 - lane < b; p0 += 128) {
            int cc[4], gg[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {                                 // four steps in flight before the reductions
                const int64_t p = p0 + 32 * u;
                cc[u] = p < b ? __ldcs(val + p) : 0;
                gg[u] = p < b ? __ldcs(col + p) : -1;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (p0 - lane + 32 * u >= b) break;                       // warp-uniform
                const int c = cc[u], g = gg[u];
                double x = ab_shfl_f64(tab, (c - 1) & 31);
                if (g >= 0) {
                    if (c > 32) x = ab_norm_value((double)c, lib, scale);
                    atomicAdd(gsum + g, x);
                    atomicAdd(gsumsq + g, __dmul_rn(x, x));
                }
            }
        }
    }
}

extern "C" int ab_norm_moments(ab_ctx *ctx, const int64_t *rowptr, const int32_t *col, const int32_t *val,
                               int64_t n_local, int32_t n_genes, double scale, int64_t *libsize,
                               double *gsum, double *gsumsq, void *stream) {
    AB_REQUIRE(ctx, "ab_norm_moments: null ctx");
    AB_REQUIRE(n_local >= 0 && n_genes > 0, "ab_norm_moments: bad shape");
    if (n_local == 0) return 0;
    int64_t blocks = (n_local + NORM_WARPS - 1) / NORM_WARPS;
    int64_t cap = (int64_t)ctx->sm_count * 8 * 4;         // 8 CTAs/SM resident, 4 waves
    if (blocks > cap) blocks = cap;
    ab_prof_begin(ctx, AB_PROF_NORM, (cudaStream_t)stream);
    norm_moments_kernel<<<(unsigned)blocks, NORM_WARPS * 32, 0, (cudaStream_t)stream>>>(
        rowptr, col, val, n_local, n_genes, scale, libsize, gsum, gsumsq);
    AB_LAUNCH_CHECK(ctx);
    ab_prof_end(ctx, AB_PROF_NORM, (cudaStream_t)stream);
    return 0;
}

// ---------------------------------------------------------------------------------------
// K3: HVG extract
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NORM_WARPS * 32)
hvg_count_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col, int64_t n_local,
                 const int32_t *__restrict__ gene2hvg, int64_t *__restrict__ rowcount) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = (int64_t)blockIdx.x * NORM_WARPS + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * NORM_WARPS;
    for (int64_t row = warp_global; row < n_local; row += nwarps) {
        const int64_t a = rowptr[row], b = rowptr[row + 1];
        int cnt = 0;
        // four steps of 32 column ids in flight before the look-ups (a pure stream: bytes in flight set the speed)
        for (int64_t p0 = a + lane; p0 < b; p0 += 128) {
            int g[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) g[u] = p0 + 32 * u < b ? __ldcs(col + p0 + 32 * u) : -1;
#pragma unroll
            for (int u = 0; u < 4; ++u) cnt += g[u] >= 0 && __ldg(gene2hvg + g[u]) >= 0;
        }
        cnt = ab_warp_sum_i32(cnt);
        if (lane == 0) rowcount[row] = cnt;
    }
}

__global__ void __launch_bounds__(NORM_WARPS * 32)
hvg_fill_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                const int32_t *__restrict__ val, const int64_t *__restrict__ libsize, int64_t n_local,
                const int32_t *__restrict__ gene2hvg, double scale, const int64_t *__restrict__ rowptr_h,
                int32_t *__restrict__ col_h, double *__restrict__ val_h) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = (int64_t)blockIdx.x * NORM_WARPS + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * NORM_WARPS;
    for (int64_t row = warp_global; row < n_local; row += nwarps) {
        const int64_t a = rowptr[row], b = rowptr[row + 1];
        const double lib = (double)libsize[row];
        const double tab = ab_norm_value((double)(lane + 1), lib, scale);
        int64_t out = rowptr_h[row];
        for (int64_t base4 = a; base4 < b; base4 += 128) {
            // four steps loaded before any is compacted
            int cc[4], gg[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t p = base4 + 32 * u + lane;
                cc[u] = p < b ? __ldcs(val + p) : 0;
                gg[u] = p < b ? __ldcs(col + p) : -1;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (base4 + 32 * u >= b) break;                               // warp-uniform
                const int c = cc[u];
                const int h = gg[u] >= 0 ? __ldg(gene2hvg + gg[u]) : -1;
                double x = ab_shfl_f64(tab, (c - 1) & 31);
                const bool keep = h >= 0;
                const unsigned m = __ballot_sync(AB_FULL, keep);
                if (keep) {
                    if (c > 32) x = ab_norm_value((double)c, lib, scale);
                    const int64_t q = out + __popc(m & ((1u << lane) - 1));
                    col_h[q] = h;
                    val_h[q] = x;
                }
                out += __popc(m);
            }
        }
    }
}

extern "C" size_t ab_hvg_extract_ws_bytes(int64_t n_local) { return ab_scan_ws_bytes(n_local) + 256; }

extern "C" int ab_hvg_extract_count(ab_ctx *ctx, const int64_t *rowptr, const int32_t *col, int64_t n_local,
                                    const int32_t *gene2hvg, int64_t *rowptr_h, int64_t *h_nnz, void *ws,
                                    size_t ws_bytes, void *stream) {
    AB_REQUIRE(ctx, "ab_hvg_extract_count: null ctx");
    AB_REQUIRE(ws_bytes >= ab_hvg_extract_ws_bytes(n_local), "ab_hvg_extract_count: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_local > 0) {
        int64_t blocks = (n_local + NORM_WARPS - 1) / NORM_WARPS;
        int64_t cap = (int64_t)ctx->sm_count * 8 * 4;
        if (blocks > cap) blocks = cap;
        ab_prof_begin(ctx, AB_PROF_HVG_EXTRACT, st);
        hvg_count_kernel<<<(unsigned)blocks, NORM_WARPS * 32, 0, st>>>(rowptr, col, n_local, gene2hvg, rowptr_h);
        AB_LAUNCH_CHECK(ctx);
        ab_prof_end(ctx, AB_PROF_HVG_EXTRACT, st);
    }
    if (ab_exclusive_scan_i64(ctx, rowptr_h, rowptr_h, n_local, ws, ws_bytes, st)) return 1;
    if (h_nnz) {
        AB_CUDA(cudaMemcpyAsync(h_nnz, rowptr_h + n_local, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        AB_CUDA(cudaStreamSynchronize(st));
    }
    return 0;
}

extern "C" int ab_hvg_extract_fill(ab_ctx *ctx, const int64_t *rowptr, const int32_t *col, const int32_t *val,
                                   const int64_t *libsize, int64_t n_local, const int32_t *gene2hvg,
                                   double scale, const int64_t *rowptr_h, int32_t *col_h, double *val_h,
                                   void *stream) {
    AB_REQUIRE(ctx, "ab_hvg_extract_fill: null ctx");
    if (n_local == 0) return 0;
    int64_t blocks = (n_local + NORM_WARPS - 1) / NORM_WARPS;
    int64_t cap = (int64_t)ctx->sm_count * 8 * 4;
    if (blocks > cap) blocks = cap;
    ab_prof_begin(ctx, AB_PROF_HVG_EXTRACT, (cudaStream_t)stream);
    hvg_fill_kernel<<<(unsigned)blocks, NORM_WARPS * 32, 0, (cudaStream_t)stream>>>(
        rowptr, col, val, libsize, n_local, gene2hvg, scale, rowptr_h, col_h, val_h);
    AB_LAUNCH_CHECK(ctx);
    ab_prof_end(ctx, AB_PROF_HVG_EXTRACT, (cudaStream_t)stream);
    return 0;
}
