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
#include <algorithm>

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

// library sizes: one warp per cell, a pure stream over the counts
__global__ void __launch_bounds__(NORM_WARPS * 32)
libsize_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ val, int64_t n_local,
               int64_t *__restrict__ libsize) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = (int64_t)blockIdx.x * NORM_WARPS + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * NORM_WARPS;
    for (int64_t row = warp_global; row < n_local; row += nwarps) {
        const long long S = row_sum_counts(val, rowptr[row], rowptr[row + 1], lane);
        if (lane == 0) libsize[row] = S;
    }
}

// Per-gene moments, deterministic and privatised.
//   * Accumulation is 64-bit FIXED POINT (x in units of 2^-45, x^2 in units of 2^-46): integer addition is
//     associative, so the sums do not depend on the order of the atomics, on the grid or on the sharding - two runs
//     and 1 / 2 / 8 GPUs give identical bits (the FP64 REDs of the first version did not).  Rounding one value costs
//     <= 2^-46 (x) / 2^-47 (x^2) absolute, far inside the 1e-12 parity bar.
//   * The accumulators live in SHARED memory: the genes are cut into `ntiles` ranges, CTA t of a thread-block
//     cluster owns range t (16 bytes per gene) and walks the SAME cell rows as its siblings, consuming only the
//     nonzeros of its range (rows are sorted by gene, so that is a prefix-skipped contiguous segment).  Measured on
//     B200 (benchmarks/micro_atomics.cu): 64-bit shared-memory adds (CAS loop) 1.9 lanes/clk/SM against 0.43-0.66 for
//     global REDs of any width - the global RED rate, not HBM, bounded the first version at 0.10 of the HBM roofline.
//   * The cluster re-synchronises every NM_SYNC_ROWS rows so that its CTAs read the same rows at the same time (the
//     second to fourth reader of a line hits L2), and flushes the tables every NM_WINDOW rows: a 64-bit table entry
//     holds at most NM_WINDOW values (<= 2^4 x 2^45 resp. 2^8 x 2^46 each: no overflow), and is added to the global
//     accumulators as two 32-bit halves in separate 64-bit words (acc[g][0..3] = x lo, x hi, x^2 lo, x^2 hi), each
//     of which can absorb 2^32 such flushes - and be summed across ranks with a plain int64 all-reduce.
static constexpr int NM_FRAC_X = 45, NM_FRAC_XX = 46;
static constexpr int NM_WINDOW = 1024;          // rows between table flushes (2^10 x 2^8 x 2^46 = 2^64)
static constexpr int NM_SYNC_ROWS = 128;        // rows between cluster barriers
static constexpr int NM_THREADS = 1024;

__device__ __forceinline__ void nm_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(NM_THREADS, 1)
norm_moments_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                    const int32_t *__restrict__ val, const int64_t *__restrict__ libsize, int64_t n_local,
                    int32_t n_genes, int32_t gtile, double scale, unsigned long long *__restrict__ acc) {
    extern __shared__ unsigned long long nm_tab[];               // [gtile][2]: sum x, sum x^2 (fixed point)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = NM_THREADS / 32;
    const int g_lo = blockIdx.x * gtile, g_hi = min(n_genes, g_lo + gtile);
    for (int i = threadIdx.x; i < 2 * gtile; i += NM_THREADS) nm_tab[i] = 0ull;
    __syncthreads();
    const int64_t n_win = (n_local + NM_WINDOW - 1) / NM_WINDOW;
    for (int64_t w = blockIdx.y; w < n_win; w += gridDim.y) {
        const int64_t w0 = w * NM_WINDOW, w1 = min(n_local, w0 + NM_WINDOW);
        for (int64_t s0 = w0; s0 < w1; s0 += NM_SYNC_ROWS) {
            const int64_t s1 = min(w1, s0 + NM_SYNC_ROWS);
            for (int64_t row = s0 + warp; row < s1; row += NW) {
                const int64_t a = rowptr[row], b = rowptr[row + 1];
                const double lib = (double)libsize[row];
                const double tab = ab_norm_value((double)(lane + 1), lib, scale);      // lane l: x of count l+1
                for (int64_t p0 = a; p0 < b; p0 += 128) {
                    int gg[4], cc[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {                                      // four steps of ids in flight
                        const int64_t p = p0 + 32 * u + lane;
                        gg[u] = p < b ? __ldg(col + p) : INT_MAX;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {                                      // counts only inside the range
                        const int64_t p = p0 + 32 * u + lane;
                        cc[u] = (gg[u] >= g_lo && gg[u] < g_hi) ? __ldg(val + p) : 0;
                    }
                    bool done = false;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int c = cc[u];
                        double x = ab_shfl_f64(tab, (c - 1) & 31);
                        if (c > 0) {                                                   // in range, and not a stored zero
                            if (c > 32) x = ab_norm_value((double)c, lib, scale);
                            const unsigned long long fx = __double2ull_rn(x * (double)(1ull << NM_FRAC_X));
                            const unsigned long long fxx = __double2ull_rn(__dmul_rn(x, x) * (double)(1ull << NM_FRAC_XX));
                            unsigned long long *t = nm_tab + 2 * (size_t)(gg[u] - g_lo);
                            atomicAdd(t, fx);
                            atomicAdd(t + 1, fxx);
                        }
                        // genes ascend inside a row: once a step starts at or beyond the range the row is finished
                        done = __shfl_sync(AB_FULL, gg[u], 0) >= g_hi;
                        if (done) break;
                    }
                    if (done) break;
                }
            }
            if (gridDim.x > 1) nm_cluster_sync();          // siblings stay on the same rows (L2 reuse); no data exchanged
        }
        __syncthreads();
        for (int i = threadIdx.x; i < g_hi - g_lo; i += NM_THREADS) {
            const unsigned long long sx = nm_tab[2 * i], sxx = nm_tab[2 * i + 1];
            if (sx | sxx) {
                unsigned long long *o = acc + 4 * (size_t)(g_lo + i);
                atomicAdd(o + 0, sx & 0xFFFFFFFFull);
                atomicAdd(o + 1, sx >> 32);
                atomicAdd(o + 2, sxx & 0xFFFFFFFFull);
                atomicAdd(o + 3, sxx >> 32);
                nm_tab[2 * i] = 0ull;
                nm_tab[2 * i + 1] = 0ull;
            }
        }
        __syncthreads();
    }
}

// acc[g][0..3] -> gsum, gsumsq (FP64): exact integer recombination, one rounding per value
__global__ void moments_finalize_kernel(const unsigned long long *__restrict__ acc, int32_t n_genes,
                                        double *__restrict__ gsum, double *__restrict__ gsumsq) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_genes) return;
    const unsigned long long *a = acc + 4 * (size_t)g;
    const unsigned __int128 sx = ((unsigned __int128)a[1] << 32) + a[0];
    const unsigned __int128 sxx = ((unsigned __int128)a[3] << 32) + a[2];
    // value < 2^100: split at 2^48 so that both pieces convert exactly, then one rounded addition
    const double x = (double)(unsigned long long)(sx >> 48) * 0x1p48 + (double)(unsigned long long)(sx & 0xFFFFFFFFFFFFull);
    const double xx = (double)(unsigned long long)(sxx >> 48) * 0x1p48 + (double)(unsigned long long)(sxx & 0xFFFFFFFFFFFFull);
    gsum[g] = x * (1.0 / (double)(1ull << NM_FRAC_X));
    gsumsq[g] = xx * (1.0 / (double)(1ull << NM_FRAC_XX));
}

static int nm_tiles_for(int32_t n_genes, size_t smem_optin) {
    const size_t cap = (std::min<size_t>(smem_optin, 227 * 1024) - 2048) / 16;       // genes per CTA table
    int t = 1;
    while ((size_t)(n_genes + t - 1) / t > cap) t <<= 1;
    return t;
}

extern "C" int ab_norm_moments(ab_ctx *ctx, const int64_t *rowptr, const int32_t *col, const int32_t *val,
                               int64_t n_local, int32_t n_genes, double scale, int64_t *libsize,
                               uint64_t *acc, void *stream) {
    AB_REQUIRE(ctx, "ab_norm_moments: null ctx");
    AB_REQUIRE(n_local >= 0 && n_genes > 0, "ab_norm_moments: bad shape");
    if (n_local == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int ntiles = nm_tiles_for(n_genes, ctx->smem_optin);
    AB_REQUIRE(ntiles <= 8, "ab_norm_moments: %d genes need more than 8 shared-memory tables", n_genes);
    const int gtile = (n_genes + ntiles - 1) / ntiles;
    const unsigned blocks = ab_wave_grid(ctx, libsize_kernel, NORM_WARPS * 32, 0, (n_local + NORM_WARPS - 1) / NORM_WARPS);
    ab_prof_begin(ctx, AB_PROF_NORM, st);
    libsize_kernel<<<blocks, NORM_WARPS * 32, 0, st>>>(rowptr, val, n_local, libsize);
    AB_LAUNCH_CHECK(ctx);
    const size_t smem = (size_t)gtile * 16;
    AB_CUDA(cudaFuncSetAttribute(norm_moments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t n_win = (n_local + NM_WINDOW - 1) / NM_WINDOW;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)ntiles, (unsigned)std::max<int64_t>(1, std::min<int64_t>(n_win, ctx->sm_count / ntiles)));
    cfg.blockDim = dim3(NM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)ntiles;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    AB_CUDA(cudaLaunchKernelEx(&cfg, norm_moments_kernel, rowptr, col, val, (const int64_t *)libsize, n_local, n_genes,
                               (int32_t)gtile, scale, (unsigned long long *)acc));
    AB_LAUNCH_CHECK(ctx);
    ab_prof_end(ctx, AB_PROF_NORM, st);
    return 0;
}

extern "C" int ab_moments_finalize(ab_ctx *ctx, const uint64_t *acc, int32_t n_genes, double *gsum, double *gsumsq,
                                   void *stream) {
    AB_REQUIRE(ctx && acc && gsum && gsumsq, "ab_moments_finalize: null argument");
    moments_finalize_kernel<<<(n_genes + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        (const unsigned long long *)acc, n_genes, gsum, gsumsq);
    AB_LAUNCH_CHECK(ctx);
    return 0;
}

// ---------------------------------------------------------------------------------------
// N4: per-gene moments of the LINEAR-scale values y = 2^x - 1 that hvg_brennecke reads
// (alona/hvg.py:83-94: `2**data_norm - 1`, then mean / var per gene).  Same row walk as K1; a selector that is
// rarely used, so the sums go to L2-resident FP64 accumulators with plain REDs (about 2e-16 relative run-to-run
// jitter from the summation order; the HVG decision is a chi-square test against an FDR of 0.1).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NORM_WARPS * 32)
lin_moments_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col, const int32_t *__restrict__ val,
                   const int64_t *__restrict__ libsize, int64_t n_local, double scale, double *__restrict__ ysum,
                   double *__restrict__ ysumsq) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = (int64_t)blockIdx.x * NORM_WARPS + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * NORM_WARPS;
    for (int64_t row = warp_global; row < n_local; row += nwarps) {
        const int64_t a = rowptr[row], b = rowptr[row + 1];
        const double lib = (double)libsize[row];
        // lane l: y of count l+1, formed the way the reference does: 2**log2(q + 1) - 1
        const double tab = __dsub_rn(exp2(ab_norm_value((double)(lane + 1), lib, scale)), 1.0);
        for (int64_t p0 = a + lane; p0 - lane < b; p0 += 128) {
            int cc[4], gg[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t p = p0 + 32 * u;
                cc[u] = p < b ? __ldcs(val + p) : 0;
                gg[u] = p < b ? __ldcs(col + p) : -1;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (p0 - lane + 32 * u >= b) break;                       // warp-uniform
                const int c = cc[u], g = gg[u];
                double y = ab_shfl_f64(tab, (c - 1) & 31);
                if (g >= 0 && c > 0) {
                    if (c > 32) y = __dsub_rn(exp2(ab_norm_value((double)c, lib, scale)), 1.0);
                    atomicAdd(ysum + g, y);
                    atomicAdd(ysumsq + g, __dmul_rn(y, y));
                }
            }
        }
    }
}

extern "C" int ab_norm_moments_linear(ab_ctx *ctx, const int64_t *rowptr, const int32_t *col, const int32_t *val,
                                      const int64_t *libsize, int64_t n_local, int32_t n_genes, double scale,
                                      double *ysum, double *ysumsq, void *stream) {
    AB_REQUIRE(ctx, "ab_norm_moments_linear: null ctx");
    AB_REQUIRE(n_local >= 0 && n_genes > 0, "ab_norm_moments_linear: bad shape");
    if (n_local == 0) return 0;
    const unsigned blocks = ab_wave_grid(ctx, lin_moments_kernel, NORM_WARPS * 32, 0, (n_local + NORM_WARPS - 1) / NORM_WARPS);
    lin_moments_kernel<<<blocks, NORM_WARPS * 32, 0, (cudaStream_t)stream>>>(rowptr, col, val, libsize, n_local,
                                                                                       scale, ysum, ysumsq);
    AB_LAUNCH_CHECK(ctx);
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
                if (c <= 0) x = 0.0;                                          // explicitly stored zero: log2(0 + 1)
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
        const unsigned blocks = ab_wave_grid(ctx, hvg_count_kernel, NORM_WARPS * 32, 0, (n_local + NORM_WARPS - 1) / NORM_WARPS);
        ab_prof_begin(ctx, AB_PROF_HVG_EXTRACT, st);
        hvg_count_kernel<<<blocks, NORM_WARPS * 32, 0, st>>>(rowptr, col, n_local, gene2hvg, rowptr_h);
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
    const unsigned blocks = ab_wave_grid(ctx, hvg_fill_kernel, NORM_WARPS * 32, 0, (n_local + NORM_WARPS - 1) / NORM_WARPS);
    ab_prof_begin(ctx, AB_PROF_HVG_EXTRACT, (cudaStream_t)stream);
    hvg_fill_kernel<<<blocks, NORM_WARPS * 32, 0, (cudaStream_t)stream>>>(
        rowptr, col, val, libsize, n_local, gene2hvg, scale, rowptr_h, col_h, val_h);
    AB_LAUNCH_CHECK(ctx);
    ab_prof_end(ctx, AB_PROF_HVG_EXTRACT, (cudaStream_t)stream);
    return 0;
}
