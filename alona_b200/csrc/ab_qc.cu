// N3: the upstream QC metrics and filters of the reference on the cell-major count CSR.
//
// Reference (dense genes x cells DataFrame):
//   remove_empty      alona/cell.py:64-81    cells / genes whose counts are all zero
//   simple_filters    alona/cell.py:141-173  cells with sum(counts) > minreads; genes expressed (count > 0) in
//                                            more than `minexpgenes` cells (integer) or fraction of cells (float)
//   QC metrics        alona/cell.py:306-312  reads per cell, genes detected per cell, reads falling on a gene
//                                            set (mitochondrial / rRNA) per cell
// Here: ONE streaming pass produces every per-cell and per-gene count those rules need (ab_qc_metrics); the
// keep/drop decisions are evaluated by the caller with the reference's own comparisons (they are O(N + G));
// ab_filter_count / ab_filter_fill then compact the CSR by a row mask and a gene map.
//
// All integer work, HBM-bound: 8 bytes per nonzero read once by the metrics pass; read twice + written once by
// the compaction (count, fill).  Per-gene counters are privatised per CTA in shared memory (4 bytes per gene,
// native 32-bit shared atomics) and flushed once, so the pass is not bound by L2 reductions the way the FP64
// moments of K1 are.
#include "ab_common.cuh"

static constexpr int QC_THREADS = 1024;

template <bool SMEM_GENES>
__global__ void __launch_bounds__(QC_THREADS, 1)
qc_metrics_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col, const int32_t *__restrict__ val,
                  int64_t n_local, int32_t n_genes, const uint8_t *__restrict__ gene_flags,
                  int64_t *__restrict__ reads, int32_t *__restrict__ genes_det, int64_t *__restrict__ set_reads,
                  unsigned long long *__restrict__ gene_cells) {
    extern __shared__ unsigned int gc_sm[];
    if (SMEM_GENES) {
        for (int i = threadIdx.x; i < n_genes; i += blockDim.x) gc_sm[i] = 0u;
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    const int64_t wid = (int64_t)blockIdx.x * wpc + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * wpc;
    for (int64_t row = wid; row < n_local; row += nw) {
        const int64_t a = rowptr[row], b = rowptr[row + 1];
        long long s = 0, s0 = 0, s1 = 0;
        int det = 0;
        // four steps of 32 nonzeros loaded before any is consumed: the pass is a pure stream, the bytes a warp keeps
        // in flight decide its speed
        for (int64_t p0 = a + lane; p0 < b; p0 += 128) {
            int cc[4], gg[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t p = p0 + 32 * u;
                cc[u] = p < b ? __ldcs(val + p) : 0;
                gg[u] = p < b ? __ldcs(col + p) : 0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = cc[u], g = gg[u];
                if (c > 0) {
                    s += c;
                    ++det;
                    if (gene_flags) {
                        const unsigned f = __ldg(gene_flags + g);
                        if (f & 1u) s0 += c;
                        if (f & 2u) s1 += c;
                    }
                    if (SMEM_GENES) atomicAdd(gc_sm + g, 1u);
                    else atomicAdd(gene_cells + g, 1ull);
                }
            }
        }
        s = ab_warp_sum_i64(s);
        det = ab_warp_sum_i32(det);
        if (gene_flags) {
            s0 = ab_warp_sum_i64(s0);
            s1 = ab_warp_sum_i64(s1);
        }
        if (lane == 0) {
            reads[row] = s;
            genes_det[row] = det;
            if (set_reads) {
                set_reads[row] = s0;
                set_reads[n_local + row] = s1;
            }
        }
    }
    if (SMEM_GENES) {
        __syncthreads();
        for (int i = threadIdx.x; i < n_genes; i += blockDim.x) {
            const unsigned int v = gc_sm[i];
            if (v) atomicAdd(gene_cells + i, (unsigned long long)v);
        }
    }
}

extern "C" int ab_qc_metrics(ab_ctx *ctx, const int64_t *rowptr, const int32_t *col, const int32_t *val,
                             int64_t n_local, int32_t n_genes, const uint8_t *gene_flags, int64_t *reads_per_cell,
                             int32_t *genes_det, int64_t *set_reads, int64_t *gene_cells, void *stream) {
    AB_REQUIRE(ctx, "ab_qc_metrics: null ctx");
    AB_REQUIRE(n_local >= 0 && n_genes > 0, "ab_qc_metrics: bad shape");
    AB_REQUIRE(reads_per_cell && genes_det && gene_cells, "ab_qc_metrics: null output");
    AB_REQUIRE(!set_reads || gene_flags, "ab_qc_metrics: set_reads needs gene_flags");
    if (n_local == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)n_genes * sizeof(unsigned int);
    const int64_t rows_per_cta = QC_THREADS / 32;
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n_local + rows_per_cta - 1) / rows_per_cta, ctx->sm_count));
    unsigned long long *gc = reinterpret_cast<unsigned long long *>(gene_cells);
    if (smem <= (size_t)ctx->smem_optin - 1024) {
        AB_CUDA(cudaFuncSetAttribute(qc_metrics_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        qc_metrics_kernel<true><<<grid, QC_THREADS, smem, st>>>(rowptr, col, val, n_local, n_genes, gene_flags,
                                                                reads_per_cell, genes_det, set_reads, gc);
    } else {
        qc_metrics_kernel<false><<<grid * 2, QC_THREADS, 0, st>>>(rowptr, col, val, n_local, n_genes, gene_flags,
                                                                  reads_per_cell, genes_det, set_reads, gc);
    }
    AB_LAUNCH_CHECK(ctx);
    return 0;
}

// ---------------------------------------------------------------------------------------
// compaction by row mask and gene map (new gene id, or -1 to drop); explicit zeros are dropped
// ---------------------------------------------------------------------------------------
static constexpr int FL_WARPS = 8;

__global__ void __launch_bounds__(FL_WARPS * 32)
filter_count_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col, const int32_t *__restrict__ val,
                    int64_t n_local, const uint8_t *__restrict__ keep_row, const int32_t *__restrict__ gene_map,
                    int64_t *__restrict__ cnt /*[n+1]*/, int64_t *__restrict__ kept /*[n+1]*/) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t)blockIdx.x * FL_WARPS + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * FL_WARPS;
    for (int64_t row = wid; row < n_local; row += nw) {
        const bool keep = keep_row ? keep_row[row] != 0 : true;
        int c = 0;
        if (keep) {
            const int64_t a = rowptr[row], b = rowptr[row + 1];
            for (int64_t p = a + lane; p < b; p += 32)
                c += (__ldg(val + p) > 0) && (gene_map ? __ldg(gene_map + __ldg(col + p)) >= 0 : true);
            c = ab_warp_sum_i32(c);
        }
        if (lane == 0) {
            cnt[row] = c;
            kept[row] = keep ? 1 : 0;
        }
    }
}

__global__ void __launch_bounds__(FL_WARPS * 32)
filter_fill_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col, const int32_t *__restrict__ val,
                   int64_t n_local, const uint8_t *__restrict__ keep_row, const int32_t *__restrict__ gene_map,
                   const int64_t *__restrict__ off /*[n+1] exclusive scan of cnt*/,
                   const int64_t *__restrict__ slot /*[n+1] exclusive scan of kept*/, int64_t *__restrict__ rowptr_out,
                   int32_t *__restrict__ col_out, int32_t *__restrict__ val_out, int64_t *__restrict__ row_src) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t)blockIdx.x * FL_WARPS + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * FL_WARPS;
    if (blockIdx.x == 0 && threadIdx.x == 0) rowptr_out[slot[n_local]] = off[n_local];
    for (int64_t row = wid; row < n_local; row += nw) {
        if (keep_row && keep_row[row] == 0) continue;
        const int64_t a = rowptr[row], b = rowptr[row + 1];
        int64_t out = off[row];
        if (lane == 0) {
            rowptr_out[slot[row]] = out;
            if (row_src) row_src[slot[row]] = row;
        }
        for (int64_t base = a; base < b; base += 32) {
            const int64_t p = base + lane;
            int c = 0, g = -1;
            if (p < b) {
                c = __ldg(val + p);
                g = __ldg(col + p);
                if (gene_map) g = __ldg(gene_map + g);
            }
            const bool k = c > 0 && g >= 0;
            const unsigned m = __ballot_sync(AB_FULL, k);
            if (k) {
                const int64_t q = out + __popc(m & ((1u << lane) - 1));
                col_out[q] = g;
                val_out[q] = c;
            }
            out += __popc(m);
        }
    }
}

struct FilterWs {
    int64_t *off, *slot;
    void *scan;
    size_t scan_bytes;
};
static size_t filter_layout(int64_t n, void *ws, size_t bytes, FilterWs *out) {
    ab_arena a(ws ? ws : (void *)256, ws ? bytes : (size_t)-1 / 2);
    FilterWs w;
    w.off = a.take<int64_t>((size_t)n + 1);
    w.slot = a.take<int64_t>((size_t)n + 1);
    w.scan_bytes = ab_scan_ws_bytes(n + 1);
    w.scan = a.take<char>(w.scan_bytes);
    if (out) *out = w;
    return a.ok ? a.off : 0;
}

extern "C" size_t ab_filter_ws_bytes(int64_t n_local) { return filter_layout(n_local, nullptr, 0, nullptr) + 256; }

extern "C" int ab_filter_count(ab_ctx *ctx, const int64_t *rowptr, const int32_t *col, const int32_t *val,
                               int64_t n_local, const uint8_t *keep_row, const int32_t *gene_map, int64_t *h_rows,
                               int64_t *h_nnz, void *ws, size_t ws_bytes, void *stream) {
    AB_REQUIRE(ctx, "ab_filter_count: null ctx");
    AB_REQUIRE(n_local >= 0 && h_rows && h_nnz, "ab_filter_count: bad arguments");
    FilterWs w;
    AB_REQUIRE(filter_layout(n_local, ws, ws_bytes, &w) != 0, "ab_filter_count: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_local > 0) {
        const unsigned grid = (unsigned)std::min<int64_t>((n_local + FL_WARPS - 1) / FL_WARPS, (int64_t)ctx->sm_count * 32);
        filter_count_kernel<<<grid, FL_WARPS * 32, 0, st>>>(rowptr, col, val, n_local, keep_row, gene_map, w.off, w.slot);
        AB_LAUNCH_CHECK(ctx);
    }
    if (ab_exclusive_scan_i64(ctx, w.off, w.off, n_local, w.scan, w.scan_bytes, st)) return 1;
    if (ab_exclusive_scan_i64(ctx, w.slot, w.slot, n_local, w.scan, w.scan_bytes, st)) return 1;
    AB_CUDA(cudaMemcpyAsync(h_nnz, w.off + n_local, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    AB_CUDA(cudaMemcpyAsync(h_rows, w.slot + n_local, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    AB_CUDA(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int ab_filter_fill(ab_ctx *ctx, const int64_t *rowptr, const int32_t *col, const int32_t *val,
                              int64_t n_local, const uint8_t *keep_row, const int32_t *gene_map, int64_t *rowptr_out,
                              int32_t *col_out, int32_t *val_out, int64_t *row_src, void *ws, size_t ws_bytes,
                              void *stream) {
    AB_REQUIRE(ctx, "ab_filter_fill: null ctx");
    FilterWs w;
    AB_REQUIRE(filter_layout(n_local, ws, ws_bytes, &w) != 0, "ab_filter_fill: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n_local + FL_WARPS - 1) / FL_WARPS, (int64_t)ctx->sm_count * 32));
    filter_fill_kernel<<<grid, FL_WARPS * 32, 0, st>>>(rowptr, col, val, n_local, keep_row, gene_map, w.off, w.slot,
                                                       rowptr_out, col_out, val_out, row_src);
    AB_LAUNCH_CHECK(ctx);
    return 0;
}
