/* alona_b200 — C-ABI of the B200-native clustering front end
 * (normalise -> HVG -> truncated PCA (IRLBA) -> exact kNN -> SNN edge list).
 *
 * The reference (oscar-franzen/alona v0.3.5) is pure Python and has no FFI; its boundary is
 * the method surface of its mixin chain.  Each entry point below replaces the arithmetic of
 * one reference method body (cited per function); the Python package alona_b200/ keeps the
 * reference's class/method/flag names and calls these through ctypes (INTEGRATION.md shows
 * the stub a maintainer would add to the reference itself).
 *
 * Conventions
 *   - plain C; every function returns 0 on success, non-zero on error;
 *     ab_last_error() returns a thread-local message.  No exceptions cross the boundary.
 *   - there is NO CPU fallback: unsupported arguments and a missing GPU are errors.
 *   - all array arguments are DEVICE pointers owned by the caller unless named h_*;
 *     `stream` is a cudaStream_t passed as void*; calls are asynchronous on it unless
 *     documented otherwise (functions that return sizes synchronise the stream).
 *   - matrices: counts are cell-major CSR (rows = cells): rowptr int64 [n+1] (local to the
 *     shard, rowptr[0]==0), col int32 (gene id, ascending within a row), val int32 (count>0).
 *   - sharding: a rank owns a contiguous range of cell rows; `n_local` is the shard height,
 *     `n_total` the global number of cells.
 *   - workspace: functions that need scratch take (ws, ws_bytes); ab_*_ws_bytes() gives the
 *     required size.  Scratch contents are undefined on return.
 */
#ifndef ALONA_B200_H
#define ALONA_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AB_VERSION 100

typedef struct ab_ctx ab_ctx;

/* ---- context ------------------------------------------------------------------------ */
int ab_version(void);
const char *ab_last_error(void);
/* hex digest of the sources (csrc/, this header, compiler flags) the library was built from, "unknown" for a build
 * outside alona_b200/build.py; the Python loader refuses a library whose digest differs from the tree next to it
 * when the sources are present. */
const char *ab_source_digest(void);
/* Binds to CUDA device `device`; fails if it is not an sm_100 part. rank/world describe the
 * row sharding (world==1: single GPU). */
int ab_ctx_create(ab_ctx **out, int device, int rank, int world);
int ab_ctx_destroy(ab_ctx *ctx);
/* Multi-GPU: dlopen()s NCCL at `libnccl_path` and joins the communicator described by the
 * 128-byte ncclUniqueId (rank 0 obtains it with ab_nccl_unique_id and broadcasts it). */
int ab_nccl_unique_id(const char *libnccl_path, void *h_id128);
int ab_ctx_init_nccl(ab_ctx *ctx, const char *libnccl_path, const void *h_id128);
/* In-place sum all-reduce of `count` float64 / int64 on the ctx communicator (no-op at world 1). */
int ab_allreduce_f64(ab_ctx *ctx, double *buf, int64_t count, void *stream);
int ab_allreduce_i64(ab_ctx *ctx, int64_t *buf, int64_t count, void *stream);
/* All-gather of equal-sized byte blocks (bytes_per_rank each) into recv (world*bytes). */
int ab_allgather_bytes(ab_ctx *ctx, const void *send, void *recv, int64_t bytes_per_rank, void *stream);
/* Multi-GPU, optional: peer-memory exchange buffers for the small all-reduces INSIDE ab_irlb (A.v result, CGS
 * coefficients, ||F||^2: three per Lanczos step, latency-bound).  With them the reductions are fused into the
 * producing / consuming kernels as one-shot pushes over NVLink peer memory (fixed rank order: bitwise identical on
 * every rank); without them ab_irlb issues ncclAllReduce between the kernels.  Protocol: every rank calls
 * ab_peer_alloc (returns its cudaIpcMemHandle, ab_peer_handle_bytes() bytes), the caller all-gathers the handles over
 * its own bootstrap (torch.distributed in alona_b200/engine.py) and every rank calls ab_peer_open with the
 * world x handle_bytes array in rank order.  All ranks must be GPUs of one NVLink/NVSwitch domain (world <= 8). */
size_t ab_peer_handle_bytes(void);
int ab_peer_alloc(ab_ctx *ctx, void *h_handle);
int ab_peer_open(ab_ctx *ctx, const void *h_handles);
int ab_peer_enabled(ab_ctx *ctx);

/* ---- K1: library-size normalisation + log2 + per-gene moments -------------------------
 * replaces AlonaCell.normalize raw branch  (alona/cell.py:241-243)
 *      and the per-gene reductions of hvg_seurat (alona/hvg.py:148,157).
 * libsize[i] = sum of the row's counts (exact); x = log2(fl(fl(c/S)*scale)+1) in FP64, the reference's arithmetic;
 * per gene, sum x and sum x^2 over this shard's cells are ACCUMULATED into acc[n_genes][4] (uint64, zeroed by the
 * caller, so shards / tiles can be chained) in 64-bit fixed point: acc[g] = {lo, hi of sum x in units of 2^-45,
 * lo, hi of sum x^2 in units of 2^-46}, value = lo + hi * 2^32.  Integer sums are order independent: the result is
 * bit-identical from run to run and for any sharding.  Across ranks the four words add with ab_allreduce_i64;
 * ab_moments_finalize turns them into the FP64 sums gsum[g], gsumsq[g] (one rounding each; error of the fixed-point
 * representation <= 2^-46 resp. 2^-47 per cell). */
int ab_norm_moments(ab_ctx *ctx, const int64_t *rowptr, const int32_t *col, const int32_t *val,
                    int64_t n_local, int32_t n_genes, double scale,
                    int64_t *libsize, uint64_t *acc, void *stream);
int ab_moments_finalize(ab_ctx *ctx, const uint64_t *acc, int32_t n_genes, double *gsum, double *gsumsq,
                        void *stream);

/* ---- N4 (next row of the scope table): per-gene moments on the LINEAR scale ----------------
 * the device part of AlonaHighlyVariableGenes.hvg_brennecke (alona/hvg.py:83-94): y = 2^x - 1 for every
 * normalised value x, ysum[g] += y, ysumsq[g] += y*y over this shard's cells (FP64, caller-zeroed, summed across
 * ranks with ab_allreduce_f64); libsize as produced by ab_norm_moments.  The selector's remaining steps (Gamma GLM
 * on <= G points, chi-square test, Benjamini-Hochberg) are O(G) host statistics in alona_b200/hvg.py. */
int ab_norm_moments_linear(ab_ctx *ctx, const int64_t *rowptr, const int32_t *col, const int32_t *val,
                           const int64_t *libsize, int64_t n_local, int32_t n_genes, double scale,
                           double *ysum, double *ysumsq, void *stream);

/* ---- K2: Seurat-style HVG selection ----------------------------------------------------
 * replaces AlonaHighlyVariableGenes.hvg_seurat (alona/hvg.py:146-165): mean, ddof-1 variance,
 * `nbins` equal-width mean bins (pandas.cut semantics), dispersion=var/mean, per-bin z-score,
 * top hvg_n by z (descending; ties by gene index; NaN last).
 * Outputs: ranked[hvg_n] gene ids in rank order, z[n_genes], gene2hvg[n_genes] = column of the
 * gene inside the HVG-restricted matrix (HVGs keep ORIGINAL gene order, clustering.py:104-105)
 * or -1, stats[4] (device int32) = {#finite z, #selected, #ties at the cut, 0}. */
int ab_hvg_seurat(ab_ctx *ctx, const double *gsum, const double *gsumsq, int64_t n_total,
                  int32_t n_genes, int32_t nbins, int32_t hvg_n,
                  int32_t *ranked, double *z, int32_t *gene2hvg, int32_t *stats, void *stream);

/* ---- K3: HVG-restricted normalised matrix A_H (cell-major CSR, FP64 values) -------------
 * replaces the row slice `data_norm[index.isin(hvg)]` (alona/clustering.py:104-105).
 * count: rowptr_h[n_local+1] (exclusive scan of per-cell HVG nonzeros); synchronises and
 * returns the shard's nnz_H in *h_nnz.  fill: col_h (0..H-1 ascending), val_h (FP64 x). */
size_t ab_hvg_extract_ws_bytes(int64_t n_local);
int ab_hvg_extract_count(ab_ctx *ctx, const int64_t *rowptr, const int32_t *col, int64_t n_local,
                         const int32_t *gene2hvg, int64_t *rowptr_h, int64_t *h_nnz,
                         void *ws, size_t ws_bytes, void *stream);
int ab_hvg_extract_fill(ab_ctx *ctx, const int64_t *rowptr, const int32_t *col, const int32_t *val,
                        const int64_t *libsize, int64_t n_local, const int32_t *gene2hvg,
                        double scale, const int64_t *rowptr_h, int32_t *col_h, double *val_h,
                        void *stream);

/* ---- N3 (next row of the scope table): QC metrics and filters on the count CSR -----------
 * replaces the dense-DataFrame passes of AlonaCell.remove_empty (alona/cell.py:64-81),
 * AlonaCell.simple_filters (alona/cell.py:141-173) and the per-cell QC metrics of
 * AlonaCell.find_low_quality_cells (alona/cell.py:306-312: reads per cell, genes detected,
 * reads on a gene set).  ab_qc_metrics: one pass; gene_flags (may be NULL) holds per gene
 * bit 0 / bit 1 = member of gene set 0 / 1 (e.g. '^mt-' genes, rRNA genes); outputs
 * reads_per_cell int64[n_local], genes_det int32[n_local] (entries > 0), set_reads
 * int64[2][n_local] (may be NULL), gene_cells int64[G] = number of cells with a count > 0,
 * ACCUMULATED into the caller-zeroed array (sum over shards with ab_allreduce_i64).
 * The keep/drop rules are evaluated by the caller with the reference's comparisons;
 * ab_filter_count / ab_filter_fill compact the CSR: keep_row uint8[n_local] (NULL = all),
 * gene_map int32[G] = new gene id or -1 (NULL = identity); explicit zeros are dropped.
 * count -> h_rows, h_nnz (host) -> caller allocates rowptr_out[h_rows+1], col_out/val_out[h_nnz],
 * optional row_src int64[h_rows] (source row of each kept row) -> fill, same workspace. */
int ab_qc_metrics(ab_ctx *ctx, const int64_t *rowptr, const int32_t *col, const int32_t *val,
                  int64_t n_local, int32_t n_genes, const uint8_t *gene_flags,
                  int64_t *reads_per_cell, int32_t *genes_det, int64_t *set_reads,
                  int64_t *gene_cells, void *stream);
size_t ab_filter_ws_bytes(int64_t n_local);
int ab_filter_count(ab_ctx *ctx, const int64_t *rowptr, const int32_t *col, const int32_t *val,
                    int64_t n_local, const uint8_t *keep_row, const int32_t *gene_map,
                    int64_t *h_rows, int64_t *h_nnz, void *ws, size_t ws_bytes, void *stream);
int ab_filter_fill(ab_ctx *ctx, const int64_t *rowptr, const int32_t *col, const int32_t *val,
                   int64_t n_local, const uint8_t *keep_row, const int32_t *gene_map,
                   int64_t *rowptr_out, int32_t *col_out, int32_t *val_out, int64_t *row_src,
                   void *ws, size_t ws_bytes, void *stream);

/* ---- N2 (next row of the scope table): per-cluster summaries of the normalised expression ---
 * replaces the dense passes of AlonaCellTypePred.mean_exp / median_exp (alona/celltypes.py:42-68:
 * data_norm.groupby(cluster, axis=1).aggregate(np.mean | np.median)) and supplies the sufficient
 * statistics of the one-way model of find_markers (alona/find_markers.py:75-111: coefficients =
 * cluster means, residual variance from sum x^2) without ever building the genes x cells frame.
 * cluster: int32[n_local], cluster id of each cell or -1 (cell ignored).  Tables are [C][G] row-major.
 * ab_cluster_moments ACCUMULATES sum x, sum x^2 (may be NULL) and the number of nonzeros into
 * caller-zeroed tables (sum over shards with ab_allreduce_f64 / _i64).
 * Medians (single shard): ab_cluster_median_count -> host int64[C] = values each cluster has to
 * gather (only (cluster, gene) pairs whose median can be nonzero); the caller picks a batch capacity
 * >= the largest entry (< 2^31), sizes the workspace with ab_cluster_median_ws_bytes and calls
 * ab_cluster_median_fill, which gathers, sorts (cub::DeviceSegmentedSort) and reads off the middle
 * order statistics, np.median convention (mean of the two middle values; NaN for an empty cluster). */
int ab_cluster_moments(ab_ctx *ctx, const int64_t *rowptr, const int32_t *col, const int32_t *val,
                       const int64_t *libsize, int64_t n_local, int32_t n_genes, double scale,
                       const int32_t *cluster, int32_t n_clusters, double *sum, double *sumsq,
                       int64_t *nnz, void *stream);
size_t ab_cluster_median_ws_bytes(int32_t n_genes, int32_t n_clusters, int64_t batch_values);
int ab_cluster_median_count(ab_ctx *ctx, const int64_t *nnz, const int64_t *h_cluster_size,
                            int32_t n_genes, int32_t n_clusters, int64_t *h_values_per_cluster,
                            void *ws, size_t ws_bytes, void *stream);
int ab_cluster_median_fill(ab_ctx *ctx, const int64_t *rowptr, const int32_t *col, const int32_t *val,
                           const int64_t *libsize, int64_t n_local, int32_t n_genes, double scale,
                           const int32_t *cluster, int32_t n_clusters, const int64_t *nnz,
                           const int64_t *h_cluster_size, int64_t batch_values, double *median,
                           void *ws, size_t ws_bytes, void *stream);

/* ---- K4-K8: truncated SVD by augmented implicitly-restarted Lanczos bidiagonalisation ---
 * replaces alona.irlbpy.lanczos (alona/irlbpy/irlb.py:95-258) as called by
 * AlonaClustering.PCA (alona/clustering.py:108-111): A = A_H^T (H x N), no centring/scaling.
 * Same work dimension (irlb.py:138), restart rule (:231-234), invcheck (:84-92), CGS order.
 * v0_local: this shard's slice of the start vector (the reference draws it on the host with
 * np.random.seed(seed); randn(N), irlb.py:151-152; normalised here).
 * center (device, n_hvg doubles, may be NULL): per-gene means for IMPLICIT centring - the operator becomes
 * A - center.1^T without ever forming it (A.x - center.(1^T x),  A^T.w - (center^T w).1); with a tight `tol` this
 * yields the leading components of the centred exact SVD that `--pca regular` computes (alona/clustering.py:114-125).
 * NULL = the uncentred operator of the irlb branch.
 * Outputs: scores_local [n_local x nval] row-major = V.diag(s) (clustering.py:111),
 * v_local [n_local x nval] row-major (may be NULL), s[nval], u[H x nval] row-major (may be
 * NULL); h_info[4] (host int64) = {steps, nmult, sum of CGS panel widths, status flags:
 * bit 0 = not converged within maxit, bit 1 = the compact row operand was used, bit 2 = the small all-reduces
 * ran fused over peer memory (ab_peer_open)}.
 * Synchronises the stream (one small D2H per restart decides convergence).
 * Workspace: ab_irlb_ws_bytes is the minimum; with ab_irlb_ws_bytes_compact (which also needs
 * the shard's nonzero count) ab_irlb re-encodes A_H once as 4-byte (gene | value-slot) words
 * plus per-row tables of the distinct values and runs the A^T.w products from that (same
 * doubles, same order, identical bits; a third of the HBM traffic). */
size_t ab_irlb_ws_bytes(int64_t n_local, int32_t n_hvg, int32_t nval);
size_t ab_irlb_ws_bytes_compact(int64_t n_local, int32_t n_hvg, int32_t nval, int64_t nnz_local);
int ab_irlb(ab_ctx *ctx, const int64_t *rowptr_h, const int32_t *col_h, const double *val_h,
            int64_t n_local, int64_t n_total, int32_t n_hvg, int32_t nval, double tol,
            int32_t maxit, const double *v0_local, const double *center, double *scores_local, double *v_local,
            double *s, double *u, int64_t *h_info, void *ws, size_t ws_bytes, void *stream);

/* ---- K9: exact k-nearest neighbours ----------------------------------------------------
 * replaces AlonaClustering.knn (alona/clustering.py:184-186; scikit-learn BallTree).
 * emb_all: N x d row-major FP64 embedding of ALL cells; queries are rows [q_begin,q_end).
 * Candidate generation: tcgen05 distance GEMM on FP16 operands (database coordinates split into
 * hi + lo halves, queries rounded to FP16: exact distances to the rounded query, whose offset from
 * the true query enters the certificate by the triangle inequality) with per-query top-m selection
 * of candidate groups;
 * then exact re-rank with the reference's own arithmetic (sequential non-fused FP64
 * sum of squared differences) and a certificate that no non-candidate can belong to the
 * top k; uncertified queries are recomputed by an exact FP64 scan.
 * idx [nq x k] int32 0-based sorted by (distance, index) (the Python layer adds 1);
 * rdist [nq x k] FP64 squared distances; stats[8] (device int64) = {rows whose k-th and
 * (k+1)-th distances tie exactly, rows recomputed by the exact scan, rows with a zero-distance
 * duplicate, rows certified by the tensor pass, ...}.
 * mode: 0 = tensor-core path (default), 1 = exact FP64 scan for every query. */
size_t ab_knn_ws_bytes(int64_t n_total, int32_t d, int64_t n_query, int32_t k);
int ab_knn(ab_ctx *ctx, const double *emb_all, int64_t n_total, int32_t d, int64_t q_begin,
           int64_t q_end, int32_t k, int32_t mode, int32_t *idx, double *rdist, int64_t *stats,
           void *ws, size_t ws_bytes, void *stream);

/* Host-only diagnostic, no device work: the schedule of the tensor-core pass for `n_query_blocks` blocks of 256 (128 for
 * wide embeddings) queries on `n_sm` persistent CTAs.  out3 = {whole rounds, tail blocks R, parts S per tail block}: CTA b
 * sweeps blocks b, b + n_sm, ... for the whole rounds, then the part items b, b + grid, ... of the R * S parts of the tail
 * phase (part p of tail block j covers tiles [p T / S, (p + 1) T / S)).  max_parts = 4 (2 for wide embeddings).
 * Exposed so that the CPU tier can check that every (block, tile) pair is swept exactly once. */
int ab_knn_schedule(int64_t n_query_blocks, int32_t n_sm, int32_t max_parts, int32_t *out3);

/* ---- K10: shared-nearest-neighbour graph ------------------------------------------------
 * replaces AlonaClustering.snn (alona/clustering.py:204-231): overlap(i,m)=|N(i) n N(m)| over
 * self-inclusive neighbour sets, keep iff overlap >= v_min (the host evaluates the reference's
 * v/(k+(k-v)) > prune_snn for v=0..k), emit rows ascending and, inside a row, in the order
 * scipy's csr_matmat produces (reverse first-encounter).  Self loops and both directions
 * are kept, as in the reference.
 * knn_all: N x k int32 0-based table of ALL cells; rows [row_begin,row_end) are emitted.
 * count: rowptr [nrows+1] int64 (exclusive scan), returns edges in *h_edges (synchronises);
 * stats (HOST int64[8], may be NULL) = {rows beyond the one-warp path, longest 2-hop walk, rows whose first
 * neighbour is not the cell itself (duplicate points), rows on the dense global path, sum of all 2-hop walk lengths
 * (x4 B = the gather traffic), number of nonzeros of A.A^T in the row range BEFORE pruning (the denominator of the
 * reference's "links pruned" message, clustering.py:231-235), rows in the 4096- / 8192-slot CTA classes, rows in
 * the 16384-slot class};
 * fill: src/dst int32 [E], overlap uint8 [E] (may be NULL). */
size_t ab_snn_ws_bytes(int64_t n_total, int32_t k, int64_t n_rows);
int ab_snn_count(ab_ctx *ctx, const int32_t *knn_all, int64_t n_total, int32_t k,
                 int64_t row_begin, int64_t row_end, int32_t v_min, int64_t *rowptr,
                 int64_t *h_edges, int64_t *stats, void *ws, size_t ws_bytes, void *stream);
int ab_snn_fill(ab_ctx *ctx, const int32_t *knn_all, int64_t n_total, int32_t k,
                int64_t row_begin, int64_t row_end, int32_t v_min, const int64_t *rowptr,
                int32_t *src, int32_t *dst, uint8_t *overlap, void *ws, size_t ws_bytes,
                void *stream);

/* ---- hand-off: the edge list as the bytes of csvs/snn_graph.csv --------------------------------
 * replaces  df.to_csv(snn_path, header=None, index=None)  (alona/clustering.py:237-238): one
 * "source,target\n" line per edge.  count: offsets[n_edges+1] (exclusive scan of the line lengths), total
 * bytes in *h_bytes (synchronises);  fill: out[bytes] (device).  The caller copies `out` to the host and writes
 * it with one write(). */
size_t ab_edges_csv_ws_bytes(int64_t n_edges);
int ab_edges_csv_count(ab_ctx *ctx, const int32_t *src, const int32_t *dst, int64_t n_edges, int64_t *offsets,
                       int64_t *h_bytes, void *ws, size_t ws_bytes, void *stream);
int ab_edges_csv_fill(ab_ctx *ctx, const int32_t *src, const int32_t *dst, int64_t n_edges, const int64_t *offsets,
                      char *out, void *stream);

/* ---- synthetic negative-binomial counts generated on the device (bench / tests only) ----
 * cluster model, Gamma-Poisson counts; data are a pure function of
 * (seed, cell id, gene id), hence identical for any sharding.  count: rowptr[n_local+1];
 * fill: col/val.  prof: n_clusters x n_genes FP32 gene probabilities (device). */
size_t ab_synth_ws_bytes(int64_t n_local);
int ab_synth_nb_count(ab_ctx *ctx, const float *prof, int32_t n_clusters, int32_t n_genes,
                      int64_t cell_begin, int64_t n_local, double lbar, double depth_sigma,
                      double r_disp, uint64_t seed, int64_t *rowptr, int64_t *h_nnz,
                      void *ws, size_t ws_bytes, void *stream);
int ab_synth_nb_fill(ab_ctx *ctx, const float *prof, int32_t n_clusters, int32_t n_genes,
                     int64_t cell_begin, int64_t n_local, double lbar, double depth_sigma,
                     double r_disp, uint64_t seed, const int64_t *rowptr, int32_t *col,
                     int32_t *val, void *stream);

/* number of kernels launched by this library on this thread's context since creation */
int64_t ab_launch_count(ab_ctx *ctx);

/* ---- measurement: CUDA-event pairs recorded on the launching stream around the named kernels ----
 * level 0 = off (default), 1 = one pair per stage-level kernel (norm_moments, hvg_extract, knn_prep,
 * knn_main, knn_rerank, snn_build, snn_count, snn_fill, irlb_total), 2 = additionally every IRLBA launch
 * (irlb_av, irlb_atw, irlb_panel, irlb_small, irlb_ritz, irlb_svd).  ab_prof_read synchronises the slot's events, returns the
 * summed device time in ms and the number of pairs since the last read, and clears the slot.
 * (bench.py's roofline fields come from here; the reference has no counterpart.) */
#define AB_PROF_SLOTS 16
int ab_prof_set_level(ab_ctx *ctx, int level);
const char *ab_prof_slot_name(int slot);
int ab_prof_read(ab_ctx *ctx, int slot, double *h_ms_total, int64_t *h_count);

#ifdef __cplusplus
}
#endif
#endif
