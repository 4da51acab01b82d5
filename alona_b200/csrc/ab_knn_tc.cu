// K9 tensor-core candidate pass (placeholder until the tcgen05 kernel lands): routes to the
// exact FP64 scan so that the pipeline is correct end to end.
#include "ab_common.cuh"
size_t ab_knn_exact_ws_bytes(int64_t nq, int k, int splits);
int ab_knn_exact_pick_splits(ab_ctx *ctx, int64_t nq);
int ab_knn_exact_run(ab_ctx *ctx, const double *emb, int64_t n_total, int d, const int64_t *qlist, int64_t q_begin,
                     int64_t nq, int k, int splits, const int64_t *out_rows, int32_t *idx, double *rdist,
                     unsigned long long *stats, void *ws, size_t ws_bytes, cudaStream_t st);
size_t ab_knn_tc_ws_bytes(int64_t n_total, int d, int64_t nq, int k) { return ab_knn_exact_ws_bytes(nq, k, 64); }
int ab_knn_tc_run(ab_ctx *ctx, const double *emb, int64_t n_total, int d, int64_t q_begin, int64_t nq, int k,
                  int32_t *idx, double *rdist, unsigned long long *stats, void *ws, size_t ws_bytes, cudaStream_t st) {
    return ab_knn_exact_run(ctx, emb, n_total, d, nullptr, q_begin, nq, k, ab_knn_exact_pick_splits(ctx, nq), nullptr,
                            idx, rdist, stats, ws, ws_bytes, st);
}
