/* TEST INFRASTRUCTURE - exact brute-force kNN, the checker for alona/clustering.py:179-188.
 *
 * The reference calls scikit-learn NearestNeighbors(algorithm='ball_tree') (un-vendored
 * dependency, scikit-learn 1.9.0 in this image).  Its point-to-point reduced distance is the
 * sequential, non-fused FP64 sum  d += (a_j-b_j)*(a_j-b_j), j ascending
 * (sklearn/metrics/_dist_metrics.pxd.tp:39-49) and an exact kNN under that distance is what
 * the tree returns; rows are ordered by distance.  This file restates that as a plain scan:
 * every query against every database point, keep the k smallest under the total order
 * (rdist, index).  Compile with -ffp-contract=off so no FMA is formed (oracle/Makefile).
 *
 * Only tests/, smoke() and bench.py's CPU-baseline legs may load this library.
 */
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline int before(double da, int64_t ia, double db, int64_t ib) {
    return da < db || (da == db && ia < ib);
}

int oracle_knn_bruteforce(const double *db, int64_t n, const double *q, int64_t nq, int d,
                          int k, int64_t *idx_out, double *rd_out, int threads) {
    if (k > n || k <= 0 || d <= 0) return 1;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    int failed = 0;
#pragma omp parallel
    {
        double *hd = (double *)malloc(sizeof(double) * (size_t)k);
        int64_t *hi = (int64_t *)malloc(sizeof(int64_t) * (size_t)k);
        if (!hd || !hi) {
#pragma omp atomic write
            failed = 1;
        } else {
#pragma omp for schedule(dynamic, 16)
            for (int64_t r = 0; r < nq; ++r) {
                const double *a = q + r * (int64_t)d;
                int cnt = 0;
                for (int64_t c = 0; c < n; ++c) {
                    const double *b = db + c * (int64_t)d;
                    double acc = 0.0;
                    for (int j = 0; j < d; ++j) {
                        double t = a[j] - b[j];
                        acc += t * t;
                    }
                    if (cnt == k && !before(acc, c, hd[k - 1], hi[k - 1])) continue;
                    /* sorted insertion under (rdist, index) */
                    int p = cnt < k ? cnt : k - 1;
                    while (p > 0 && before(acc, c, hd[p - 1], hi[p - 1])) {
                        hd[p] = hd[p - 1];
                        hi[p] = hi[p - 1];
                        --p;
                    }
                    hd[p] = acc;
                    hi[p] = c;
                    if (cnt < k) ++cnt;
                }
                for (int j = 0; j < k; ++j) {
                    idx_out[r * (int64_t)k + j] = hi[j];
                    rd_out[r * (int64_t)k + j] = hd[j];
                }
            }
        }
        free(hd);
        free(hi);
    }
    return failed;
}
