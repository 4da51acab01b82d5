// N2: per-cluster summaries of the normalised expression for the downstream consumers,
// straight from the count CSR - the dense genes x cells `data_norm` frame (291 GB at C3) is never built.
//
// Reference:
//   mean_exp    alona/celltypes.py:56-68   data_norm.groupby(cluster, axis=1).aggregate(np.mean)
//   median_exp  alona/celltypes.py:42-54   data_norm.groupby(cluster, axis=1).aggregate(np.median)
//   fit_lm_tt   alona/find_markers.py:75-111   one-way model ~0 + C(cl): coefficients = cluster means, residual
//                                              variance per gene from sum x^2 - sum_c n_c mean_c^2
// x = log2(count / S_cell * 1e4 + 1) (alona/cell.py:241-243), evaluated by the same routine as K1/K3.
//
// ab_cluster_moments: one streaming pass, three L2 reductions per nonzero into [C][G] tables (sum x, sum x^2,
// number of nonzeros).  ab_cluster_median_*: a (cluster, gene) pair has a nonzero median only if the upper middle
// rank falls past its zeros; the nonzero values of exactly those pairs are gathered into segments, sorted with
// cub::DeviceSegmentedSort and the middle order statistics are read off (np.median: mean of the two middle
// values).  HBM-bound integer/FP64 streaming; no tensor work.
#include "ab_common.cuh"
#include <cub/device/device_segmented_sort.cuh>
#include <vector>

static constexpr int CL_WARPS = 8;

__global__ void __launch_bounds__(CL_WARPS * 32)
cluster_moments_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col, const int32_t *__restrict__ val,
                       const int64_t *__restrict__ libsize, int64_t n_local, int32_t n_genes, double scale,
                       const int32_t *__restrict__ cluster, int32_t n_clusters, double *__restrict__ sum,
                       double *__restrict__ sumsq, unsigned long long *__restrict__ nnz) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t)blockIdx.x * CL_WARPS + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * CL_WARPS;
    for (int64_t row = wid; row < n_local; row += nw) {
        const int cl = cluster[row];
        if (cl < 0 || cl >= n_clusters) continue;
        const int64_t a = rowptr[row], b = rowptr[row + 1];
        const double lib = (double)libsize[row];
        const double tab = ab_norm_value((double)(lane + 1), lib, scale);      // lane l: x for count l+1
        const size_t base_cg = (size_t)cl * n_genes;
        for (int64_t p0 = a; p0 < b; p0 += 32) {
            const int64_t p = p0 + lane;
            const bool live = p < b;
            int c = 0, g = 0;
            if (live) { c = __ldg(val + p); g = __ldg(col + p); }
            double x = ab_shfl_f64(tab, (c - 1) & 31);
            if (live && c > 0) {
                if (c > 32) x = ab_norm_value((double)c, lib, scale);
                atomicAdd(sum + base_cg + g, x);
                if (sumsq) atomicAdd(sumsq + base_cg + g, __dmul_rn(x, x));
                atomicAdd(nnz + base_cg + g, 1ull);
            }
        }
    }
}

extern "C" int ab_cluster_moments(ab_ctx *ctx, const int64_t *rowptr, const int32_t *col, const int32_t *val,
                                  const int64_t *libsize, int64_t n_local, int32_t n_genes, double scale,
                                  const int32_t *cluster, int32_t n_clusters, double *sum, double *sumsq, int64_t *nnz,
                                  void *stream) {
    AB_REQUIRE(ctx, "ab_cluster_moments: null ctx");
    AB_REQUIRE(n_local >= 0 && n_genes > 0 && n_clusters > 0, "ab_cluster_moments: bad shape");
    AB_REQUIRE(cluster && sum && nnz && libsize, "ab_cluster_moments: null argument");
    if (n_local == 0) return 0;
    const unsigned grid = (unsigned)std::min<int64_t>((n_local + CL_WARPS - 1) / CL_WARPS, (int64_t)ctx->sm_count * 32);
    cluster_moments_kernel<<<grid, CL_WARPS * 32, 0, (cudaStream_t)stream>>>(
        rowptr, col, val, libsize, n_local, n_genes, scale, cluster, n_clusters, sum, sumsq,
        reinterpret_cast<unsigned long long *>(nnz));
    AB_LAUNCH_CHECK(ctx);
    return 0;
}

// ---------------------------------------------------------------------------------------
// medians
// ---------------------------------------------------------------------------------------
// seg_len[c*G+g] = nonzeros to gather for the pair (0 unless its median can be nonzero)
__global__ void median_plan_kernel(const unsigned long long *__restrict__ nnz, const int64_t *__restrict__ csize,
                                   int32_t n_genes, int64_t pairs, int64_t *__restrict__ seg_len) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pairs) return;
    const int64_t n = csize[i / n_genes], k = (int64_t)nnz[i];
    const int64_t zeros = n - k, hi = n / 2;                  // upper middle rank (0-based) of n sorted values
    seg_len[i] = (n > 0 && hi >= zeros) ? k : 0;
}

__global__ void __launch_bounds__(CL_WARPS * 32)
median_gather_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col, const int32_t *__restrict__ val,
                     const int64_t *__restrict__ libsize, int64_t n_local, int32_t n_genes, double scale,
                     const int32_t *__restrict__ cluster, int c_begin, int c_end, const int64_t *__restrict__ seg_off,
                     int64_t base, unsigned int *__restrict__ cursor, double *__restrict__ vals) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t)blockIdx.x * CL_WARPS + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * CL_WARPS;
    for (int64_t row = wid; row < n_local; row += nw) {
        const int cl = cluster[row];
        if (cl < c_begin || cl >= c_end) continue;
        const int64_t a = rowptr[row], b = rowptr[row + 1];
        const double lib = (double)libsize[row];
        const double tab = ab_norm_value((double)(lane + 1), lib, scale);
        const size_t base_cg = (size_t)cl * n_genes;
        for (int64_t p0 = a; p0 < b; p0 += 32) {
            const int64_t p = p0 + lane;
            const bool live = p < b;
            int c = 0, g = 0;
            if (live) { c = __ldg(val + p); g = __ldg(col + p); }
            double x = ab_shfl_f64(tab, (c - 1) & 31);
            if (live && c > 0) {
                const size_t pair = base_cg + g;
                const int64_t s0 = seg_off[pair], s1 = seg_off[pair + 1];
                if (s1 > s0) {
                    if (c > 32) x = ab_norm_value((double)c, lib, scale);
                    vals[s0 - base + atomicAdd(cursor + pair, 1u)] = x;
                }
            }
        }
    }
}

// segment offsets of one batch of clusters, relative to the batch's first value
__global__ void median_rebase_kernel(const int64_t *__restrict__ seg_off, int64_t first_pair, int64_t npairs, int64_t base,
                                     int64_t *__restrict__ rel) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= npairs) rel[i] = seg_off[first_pair + i] - base;
}

__global__ void median_pick_kernel(const unsigned long long *__restrict__ nnz, const int64_t *__restrict__ csize,
                                   int32_t n_genes, int64_t first_pair, int64_t npairs, const int64_t *__restrict__ rel,
                                   const double *__restrict__ sorted, double *__restrict__ median) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= npairs) return;
    const int64_t i = first_pair + j;
    const int64_t n = csize[i / n_genes];
    if (n <= 0) { median[i] = __longlong_as_double(0x7FF8000000000000ll); return; }      // empty cluster: NaN (np.median)
    const int64_t k = (int64_t)nnz[i], zeros = n - k;
    const int64_t lo = (n - 1) / 2, hi = n / 2;
    const int64_t s0 = rel[j];
    if (rel[j + 1] == s0) { median[i] = 0.0; return; }
    const double vh = sorted[s0 + (hi - zeros)];
    const double vl = lo >= zeros ? sorted[s0 + (lo - zeros)] : 0.0;
    median[i] = lo == hi ? vh : __dmul_rn(__dadd_rn(vl, vh), 0.5);
}

struct MedianWs {
    int64_t *seg_len, *seg_off, *rel, *csize;
    unsigned int *cursor;
    double *vals, *vals2;
    void *scan, *cub;
    size_t scan_bytes, cub_bytes;
};
static size_t median_layout(int32_t n_genes, int32_t n_clusters, int64_t batch, void *ws, size_t bytes, MedianWs *out) {
    ab_arena a(ws ? ws : (void *)256, ws ? bytes : (size_t)-1 / 2);
    MedianWs w;
    const int64_t pairs = (int64_t)n_genes * n_clusters;
    w.seg_len = a.take<int64_t>((size_t)pairs + 1);
    w.seg_off = a.take<int64_t>((size_t)pairs + 1);
    w.rel = a.take<int64_t>((size_t)pairs + 1);
    w.csize = a.take<int64_t>((size_t)n_clusters);
    w.cursor = a.take<unsigned int>((size_t)pairs);
    w.scan_bytes = ab_scan_ws_bytes(pairs + 1);
    w.scan = a.take<char>(w.scan_bytes);
    w.vals = a.take<double>((size_t)std::max<int64_t>(batch, 1));
    w.vals2 = a.take<double>((size_t)std::max<int64_t>(batch, 1));
    cub::DoubleBuffer<double> db(nullptr, nullptr);
    size_t tb = 0;
    cub::DeviceSegmentedSort::SortKeys(nullptr, tb, db, (int)std::min<int64_t>(std::max<int64_t>(batch, 1), INT_MAX - 1),
                                       (int)std::min<int64_t>(pairs, INT_MAX - 1), (const int64_t *)nullptr,
                                       (const int64_t *)nullptr);
    w.cub_bytes = tb + 256;
    w.cub = a.take<char>(w.cub_bytes);
    if (out) *out = w;
    return a.ok ? a.off : 0;
}

static int median_plan(ab_ctx *ctx, const int64_t *nnz, const int64_t *h_cluster_size, int32_t n_genes, int32_t n_clusters,
                       MedianWs &w, cudaStream_t st) {
    const int64_t pairs = (int64_t)n_genes * n_clusters;
    AB_CUDA(cudaMemcpyAsync(w.csize, h_cluster_size, (size_t)n_clusters * 8, cudaMemcpyHostToDevice, st));
    median_plan_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, st>>>(reinterpret_cast<const unsigned long long *>(nnz),
                                                                        w.csize, n_genes, pairs, w.seg_len);
    AB_LAUNCH_CHECK(ctx);
    return ab_exclusive_scan_i64(ctx, w.seg_len, w.seg_off, pairs, w.scan, w.scan_bytes, st);
}

extern "C" size_t ab_cluster_median_ws_bytes(int32_t n_genes, int32_t n_clusters, int64_t batch_values) {
    return median_layout(n_genes, n_clusters, batch_values, nullptr, 0, nullptr) + 256;
}

// step 1: how many values each cluster has to gather (host int64[C]); sizes the batches of step 2
extern "C" int ab_cluster_median_count(ab_ctx *ctx, const int64_t *nnz, const int64_t *h_cluster_size, int32_t n_genes,
                                       int32_t n_clusters, int64_t *h_values_per_cluster, void *ws, size_t ws_bytes,
                                       void *stream) {
    AB_REQUIRE(ctx && nnz && h_cluster_size && h_values_per_cluster, "ab_cluster_median_count: null argument");
    MedianWs w;
    AB_REQUIRE(median_layout(n_genes, n_clusters, 0, ws, ws_bytes, &w) != 0, "ab_cluster_median_count: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    if (median_plan(ctx, nnz, h_cluster_size, n_genes, n_clusters, w, st)) return 1;
    std::vector<int64_t> cum((size_t)n_clusters + 1);
    AB_CUDA(cudaMemcpy2DAsync(cum.data(), 8, w.seg_off, (size_t)n_genes * 8, 8, (size_t)n_clusters + 1, cudaMemcpyDeviceToHost, st));
    AB_CUDA(cudaStreamSynchronize(st));
    for (int c = 0; c < n_clusters; ++c) h_values_per_cluster[c] = cum[c + 1] - cum[c];
    return 0;
}

// step 2: clusters are processed in batches of at most `batch_values` gathered values (the capacity the workspace
// was sized for): gather -> cub::DeviceSegmentedSort -> pick
extern "C" int ab_cluster_median_fill(ab_ctx *ctx, const int64_t *rowptr, const int32_t *col, const int32_t *val,
                                      const int64_t *libsize, int64_t n_local, int32_t n_genes, double scale,
                                      const int32_t *cluster, int32_t n_clusters, const int64_t *nnz,
                                      const int64_t *h_cluster_size, int64_t batch_values, double *median, void *ws,
                                      size_t ws_bytes, void *stream) {
    AB_REQUIRE(ctx && median && nnz && h_cluster_size, "ab_cluster_median_fill: null argument");
    AB_REQUIRE(batch_values >= 0 && batch_values < (int64_t)INT_MAX, "ab_cluster_median_fill: batch of %lld values exceeds the segmented sort's 2^31 limit",
               (long long)batch_values);
    MedianWs w;
    AB_REQUIRE(median_layout(n_genes, n_clusters, batch_values, ws, ws_bytes, &w) != 0, "ab_cluster_median_fill: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t pairs = (int64_t)n_genes * n_clusters;
    if (median_plan(ctx, nnz, h_cluster_size, n_genes, n_clusters, w, st)) return 1;
    std::vector<int64_t> cum((size_t)n_clusters + 1);
    AB_CUDA(cudaMemcpy2DAsync(cum.data(), 8, w.seg_off, (size_t)n_genes * 8, 8, (size_t)n_clusters + 1, cudaMemcpyDeviceToHost, st));
    AB_CUDA(cudaMemsetAsync(w.cursor, 0, (size_t)pairs * sizeof(unsigned int), st));
    AB_CUDA(cudaStreamSynchronize(st));
    const unsigned g_rows = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n_local + CL_WARPS - 1) / CL_WARPS, (int64_t)ctx->sm_count * 32));
    int c0 = 0;
    while (c0 < n_clusters) {
        int c1 = c0 + 1;
        AB_REQUIRE(cum[c1] - cum[c0] <= batch_values, "ab_cluster_median_fill: cluster %d needs %lld values, batch capacity is %lld",
                   c0, (long long)(cum[c1] - cum[c0]), (long long)batch_values);
        while (c1 < n_clusters && cum[c1 + 1] - cum[c0] <= batch_values) ++c1;
        const int64_t base = cum[c0], count = cum[c1] - cum[c0];
        const int64_t first_pair = (int64_t)c0 * n_genes, npairs = (int64_t)(c1 - c0) * n_genes;
        median_rebase_kernel<<<(unsigned)((npairs + 256) / 256), 256, 0, st>>>(w.seg_off, first_pair, npairs, base, w.rel);
        AB_LAUNCH_CHECK(ctx);
        cub::DoubleBuffer<double> db(w.vals, w.vals2);
        if (count > 0 && n_local > 0) {
            median_gather_kernel<<<g_rows, CL_WARPS * 32, 0, st>>>(rowptr, col, val, libsize, n_local, n_genes, scale, cluster, c0,
                                                                   c1, w.seg_off, base, w.cursor, w.vals);
            AB_LAUNCH_CHECK(ctx);
            size_t tb = w.cub_bytes;
            AB_CUDA(cub::DeviceSegmentedSort::SortKeys(w.cub, tb, db, (int)count, (int)npairs, w.rel, w.rel + 1, st));
        }
        median_pick_kernel<<<(unsigned)((npairs + 255) / 256), 256, 0, st>>>(reinterpret_cast<const unsigned long long *>(nnz),
                                                                             w.csize, n_genes, first_pair, npairs, w.rel,
                                                                             db.Current(), median);
        AB_LAUNCH_CHECK(ctx);
        c0 = c1;
    }
    return 0;
}
