// K9 entry point: exact kNN = tensor-core candidate pass + exact FP64 re-rank with certificate
// (ab_knn_tc.cu) with an exact FP64 scan (ab_knn_exact.cu) for uncertified queries / mode 1.
#include "ab_common.cuh"
#include <algorithm>

size_t ab_knn_exact_ws_bytes(int64_t nq, int k, int splits);
int ab_knn_exact_pick_splits(ab_ctx *ctx, int64_t nq);
int ab_knn_exact_run(ab_ctx *ctx, const double *emb, int64_t n_total, int d, const int64_t *qlist, int64_t q_begin,
                     int64_t nq, int k, int splits, const int64_t *out_rows, int32_t *idx, double *rdist,
                     unsigned long long *stats, void *ws, size_t ws_bytes, cudaStream_t st);

size_t ab_knn_tc_ws_bytes(int64_t n_total, int d, int64_t nq, int k);
int ab_knn_tc_run(ab_ctx *ctx, const double *emb, int64_t n_total, int d, int64_t q_begin, int64_t nq, int k,
                  int32_t *idx, double *rdist, unsigned long long *stats, void *ws, size_t ws_bytes, cudaStream_t st);

extern "C" size_t ab_knn_ws_bytes(int64_t n_total, int32_t d, int64_t n_query, int32_t k) {
    size_t exact = ab_knn_exact_ws_bytes(n_query, k, 64);
    size_t tc = ab_knn_tc_ws_bytes(n_total, d, n_query, k);
    return std::max(exact, tc) + 256;
}

extern "C" int ab_knn(ab_ctx *ctx, const double *emb_all, int64_t n_total, int32_t d, int64_t q_begin, int64_t q_end,
                      int32_t k, int32_t mode, int32_t *idx, double *rdist, int64_t *stats, void *ws, size_t ws_bytes,
                      void *stream) {
    AB_REQUIRE(ctx, "ab_knn: null ctx");
    AB_REQUIRE(n_total >= 1 && d >= 1, "ab_knn: bad shape");
    AB_REQUIRE(k >= 1 && k <= n_total, "ab_knn: k=%d must be in 1..n_total (scikit-learn raises the same)", k);
    AB_REQUIRE(k <= 128, "ab_knn: k=%d exceeds the supported maximum of 128", k);
    AB_REQUIRE(q_begin >= 0 && q_end >= q_begin && q_end <= n_total, "ab_knn: bad query range");
    AB_REQUIRE(mode == 0 || mode == 1, "ab_knn: mode must be 0 (tensor-core) or 1 (exact scan)");
    AB_REQUIRE(ws_bytes >= ab_knn_ws_bytes(n_total, d, q_end - q_begin, k), "ab_knn: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nq = q_end - q_begin;
    AB_CUDA(cudaMemsetAsync(stats, 0, 8 * sizeof(int64_t), st));
    if (nq == 0) return 0;
    if (mode == 1) {
        const int splits = ab_knn_exact_pick_splits(ctx, nq);
        return ab_knn_exact_run(ctx, emb_all, n_total, d, nullptr, q_begin, nq, k, splits, nullptr, idx, rdist,
                                (unsigned long long *)stats, ws, ws_bytes, st);
    }
    return ab_knn_tc_run(ctx, emb_all, n_total, d, q_begin, nq, k, idx, rdist, (unsigned long long *)stats, ws,
                         ws_bytes, st);
}
