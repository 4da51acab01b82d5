// K9 (exact part): FP64 k-nearest-neighbour scan with the reference's own distance arithmetic.
//
// The reference calls scikit-learn's BallTree (alona/clustering.py:184-186); its reduced
// distance is the sequential non-fused FP64 sum  d += (a_j-b_j)*(a_j-b_j)  (j ascending)
// and the returned rows are the exact k smallest, ordered by distance.
// This kernel reproduces that number bit for bit (__dsub_rn/__dmul_rn/__dadd_rn, no FMA)
// and orders by (distance, index).
//
// Used (a) as the re-rank/fallback engine behind the tensor-core candidate pass and
// (b) as `mode 1` of ab_knn (every query scanned exactly) — the on-device checker at scale.
//
// One thread per query, database streamed through shared memory in tiles; the database
// range can be split across blockIdx.y, partial lists are merged by a second kernel.
#include "ab_common.cuh"
#include <math.h>

static constexpr int KE_THREADS = 128;
static constexpr int KE_TILE = 32;          // database points per shared-memory tile
static constexpr int KE_DREG = 64;          // query coordinates held in registers

__device__ __forceinline__ bool pair_before(double da, int ia, double db, int ib) {
    return da < db || (da == db && ia < ib);
}

// dynamic smem: tile[KE_TILE][d] f64 | qtail[(d-KE_DREG)+][KE_THREADS] f64 | ld[kk][KE_THREADS] f64 | li[kk][KE_THREADS] i32
__global__ void __launch_bounds__(KE_THREADS)
knn_exact_scan_kernel(const double *__restrict__ emb, int64_t n_total, int d, const int64_t *__restrict__ qlist,
                      int64_t q_begin, int64_t nq, int kk, int64_t db_per_split,
                      double *__restrict__ part_d /*[splits][nq][kk]*/, int32_t *__restrict__ part_i) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *tile = reinterpret_cast<double *>(smem_raw);
    const int dtail = d > KE_DREG ? d - KE_DREG : 0;
    double *qtail = tile + (size_t)KE_TILE * d;
    double *ld = qtail + (size_t)dtail * KE_THREADS;
    int *li = reinterpret_cast<int *>(ld + (size_t)kk * KE_THREADS);
    const int t = threadIdx.x;
    const int64_t qpos = (int64_t)blockIdx.x * KE_THREADS + t;
    const bool live = qpos < nq;
    const int64_t qid = live ? (qlist ? qlist[qpos] : q_begin + qpos) : 0;
    double q[KE_DREG];
#pragma unroll
    for (int j = 0; j < KE_DREG; ++j) q[j] = (live && j < d) ? emb[qid * d + j] : 0.0;
    for (int j = 0; j < dtail; ++j) qtail[(size_t)j * KE_THREADS + t] = live ? emb[qid * d + KE_DREG + j] : 0.0;
    for (int s = 0; s < kk; ++s) { ld[(size_t)s * KE_THREADS + t] = INFINITY; li[(size_t)s * KE_THREADS + t] = INT_MAX; }
    const int64_t db0 = (int64_t)blockIdx.y * db_per_split;
    const int64_t db1 = min(n_total, db0 + db_per_split);
    double worst_d = INFINITY;
    int worst_i = INT_MAX;
    for (int64_t base = db0; base < db1; base += KE_TILE) {
        const int cnt = (int)min((int64_t)KE_TILE, db1 - base);
        __syncthreads();
        for (int e = t; e < cnt * d; e += KE_THREADS) tile[e] = emb[base * d + e];
        __syncthreads();
        if (!live) continue;
        for (int c = 0; c < cnt; ++c) {
            const double *x = tile + (size_t)c * d;
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < KE_DREG; ++j) {
                if (j < d) {
                    const double df = __dsub_rn(q[j], x[j]);
                    acc = __dadd_rn(acc, __dmul_rn(df, df));
                }
            }
            for (int j = 0; j < dtail; ++j) {
                const double df = __dsub_rn(qtail[(size_t)j * KE_THREADS + t], x[KE_DREG + j]);
                acc = __dadd_rn(acc, __dmul_rn(df, df));
            }
            const int id = (int)(base + c);
            if (pair_before(acc, id, worst_d, worst_i)) {
                int p = kk - 1;
                while (p > 0 && pair_before(acc, id, ld[(size_t)(p - 1) * KE_THREADS + t], li[(size_t)(p - 1) * KE_THREADS + t])) {
                    ld[(size_t)p * KE_THREADS + t] = ld[(size_t)(p - 1) * KE_THREADS + t];
                    li[(size_t)p * KE_THREADS + t] = li[(size_t)(p - 1) * KE_THREADS + t];
                    --p;
                }
                ld[(size_t)p * KE_THREADS + t] = acc;
                li[(size_t)p * KE_THREADS + t] = id;
                worst_d = ld[(size_t)(kk - 1) * KE_THREADS + t];
                worst_i = li[(size_t)(kk - 1) * KE_THREADS + t];
            }
        }
    }
    if (live) {
        double *od = part_d + ((size_t)blockIdx.y * nq + qpos) * kk;
        int32_t *oi = part_i + ((size_t)blockIdx.y * nq + qpos) * kk;
        for (int s = 0; s < kk; ++s) { od[s] = ld[(size_t)s * KE_THREADS + t]; oi[s] = li[(size_t)s * KE_THREADS + t]; }
    }
}

// merges `splits` sorted partial lists per query; writes k results (+ tie / duplicate statistics)
__global__ void knn_exact_merge_kernel(const double *__restrict__ part_d, const int32_t *__restrict__ part_i,
                                       int splits, int64_t nq, int kk, int k, const int64_t *__restrict__ qlist,
                                       int64_t q_begin, int32_t *__restrict__ idx, double *__restrict__ rdist,
                                       int64_t out_stride_rows /* rows of idx are addressed by qpos or by qlist rank */,
                                       const int64_t *__restrict__ out_rows, unsigned long long *__restrict__ stats) {
    const int64_t qpos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (qpos >= nq) return;
    int head[64];
    for (int s = 0; s < splits; ++s) head[s] = 0;
    const int64_t orow = out_rows ? out_rows[qpos] : qpos;
    double prev_d = -1.0, dk = 0.0, dk1 = 0.0;
    for (int r = 0; r < kk; ++r) {
        int best = -1;
        double bd = INFINITY;
        int bi = INT_MAX;
        for (int s = 0; s < splits; ++s) {
            if (head[s] >= kk) continue;
            const size_t o = ((size_t)s * nq + qpos) * kk + head[s];
            const double dd = part_d[o];
            const int ii = part_i[o];
            if (best < 0 || pair_before(dd, ii, bd, bi)) { best = s; bd = dd; bi = ii; }
        }
        head[best]++;
        if (r < k) { idx[orow * k + r] = bi; rdist[orow * k + r] = bd; }
        if (r == k - 1) dk = bd;
        if (r == k) dk1 = bd;
        prev_d = bd;
    }
    (void)prev_d;
    const int64_t qid = qlist ? qlist[qpos] : q_begin + qpos;
    if (kk > k && dk == dk1) atomicAdd(stats + 0, 1ull);                       // k-th / (k+1)-th exact tie
    if (idx[orow * k] != (int32_t)qid || (k > 1 && rdist[orow * k + 1] == 0.0)) atomicAdd(stats + 2, 1ull);
}

size_t ab_knn_exact_ws_bytes(int64_t nq, int k, int splits) {
    const int kk = k + 1;
    size_t b = 0;
    b = ab_arena_need(b, (size_t)splits * nq * kk * sizeof(double));
    b = ab_arena_need(b, (size_t)splits * nq * kk * sizeof(int32_t));
    return b + 256;
}

int ab_knn_exact_pick_splits(ab_ctx *ctx, int64_t nq) {
    int64_t qblocks = (nq + KE_THREADS - 1) / KE_THREADS;
    int64_t want = ((int64_t)ctx->sm_count * 2 + qblocks - 1) / qblocks;
    if (want < 1) want = 1;
    if (want > 64) want = 64;
    return (int)want;
}

// qlist (device, may be NULL: queries q_begin..q_begin+nq-1); out_rows (device, may be NULL): row of
// idx/rdist each query writes.  stats: device uint64[8].
int ab_knn_exact_run(ab_ctx *ctx, const double *emb, int64_t n_total, int d, const int64_t *qlist, int64_t q_begin,
                     int64_t nq, int k, int splits, const int64_t *out_rows, int32_t *idx, double *rdist,
                     unsigned long long *stats, void *ws, size_t ws_bytes, cudaStream_t st) {
    if (nq == 0) return 0;
    const int kk = (k + 1 <= n_total) ? k + 1 : k;
    AB_REQUIRE(d >= 1 && d <= KE_DREG + 192, "knn: d=%d unsupported (max %d)", d, KE_DREG + 192);
    AB_REQUIRE(splits >= 1 && splits <= 64, "knn: bad split count");
    ab_arena a(ws, ws_bytes);
    double *part_d = a.take<double>((size_t)splits * nq * kk);
    int32_t *part_i = a.take<int32_t>((size_t)splits * nq * kk);
    AB_REQUIRE(a.ok, "knn exact: workspace too small");
    const int dtail = d > KE_DREG ? d - KE_DREG : 0;
    size_t smem = (size_t)KE_TILE * d * 8 + (size_t)dtail * KE_THREADS * 8 + (size_t)kk * KE_THREADS * 12;
    AB_REQUIRE(smem <= ctx->smem_optin, "knn exact: k=%d d=%d needs %zu B shared memory", k, d, smem);
    AB_CUDA(cudaFuncSetAttribute(knn_exact_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t per = ((n_total + splits - 1) / splits + KE_TILE - 1) / KE_TILE * KE_TILE;
    dim3 grid((unsigned)((nq + KE_THREADS - 1) / KE_THREADS), (unsigned)splits);
    knn_exact_scan_kernel<<<grid, KE_THREADS, smem, st>>>(emb, n_total, d, qlist, q_begin, nq, kk, per, part_d, part_i);
    AB_LAUNCH_CHECK(ctx);
    knn_exact_merge_kernel<<<(unsigned)((nq + 127) / 128), 128, 0, st>>>(part_d, part_i, splits, nq, kk, k, qlist,
                                                                          q_begin, idx, rdist, 0, out_rows, stats);
    AB_LAUNCH_CHECK(ctx);
    return 0;
}
