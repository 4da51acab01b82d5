// K2: Seurat-style highly-variable-gene selection on per-gene moments.
// Restates AlonaHighlyVariableGenes.hvg_seurat (alona/hvg.py:146-165) on G scalars:
//   mean (hvg.py:148), pandas.cut equal-width bins (:150), dispersion var/mean (:157),
//   per-bin z-score with ddof-1 std, NaN-skipping (:158), descending sort + head(hvg_n) (:161-165).
// Latency-bound (G ~ 3e4 scalars): one CTA does the statistics with fixed-order reductions
// (bitwise reproducible), an O(G^2) counting rank across the whole GPU replaces the sort
// (ties broken by gene index, NaN last), a one-CTA scan numbers the selected genes in
// ORIGINAL gene order (alona/clustering.py:104-105).
#include "ab_common.cuh"
#include <math.h>

static constexpr int HT = 1024;
static constexpr int MAX_BINS = 64;

__device__ __forceinline__ double block_sum_f64(double v, double *sm /*>=33*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = ab_warp_sum_f64(v);
    __syncthreads();
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    double t = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : 0.0;
    if (warp == 0) {
        t = ab_warp_sum_f64(t);
        if (lane == 0) sm[32] = t;
    }
    __syncthreads();
    return sm[32];
}
__device__ __forceinline__ double block_min_f64(double v, double *sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = fmin(v, ab_shfl_xor_f64(v, m));
    __syncthreads();
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    double t = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : INFINITY;
    if (warp == 0) {
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) t = fmin(t, ab_shfl_xor_f64(t, m));
        if (lane == 0) sm[32] = t;
    }
    __syncthreads();
    return sm[32];
}

// z[] doubles as scratch for the dispersion; bin[] is an int32 scratch (gene2hvg reused).
__global__ void __launch_bounds__(HT)
hvg_stats_kernel(const double *__restrict__ gsum, const double *__restrict__ gsumsq, double n_total,
                 int n_genes, int nbins, double *__restrict__ z, int *__restrict__ bin) {
    __shared__ double sm[33];
    __shared__ double edges[MAX_BINS + 1];
    __shared__ double bmu[MAX_BINS], bsd[MAX_BINS];
    const int tid = threadIdx.x;
    // pass 1: mean range
    double mn = INFINITY, mxneg = INFINITY;
    for (int g = tid; g < n_genes; g += HT) {
        double mean = gsum[g] / n_total;
        mn = fmin(mn, mean);
        mxneg = fmin(mxneg, -mean);
    }
    mn = block_min_f64(mn, sm);
    const double mx = -block_min_f64(mxneg, sm);
    if (tid <= nbins) {
        // numpy.linspace(mn, mx, nbins+1): i*step + mn, last = mx ; pandas.cut: edges[0] -= 0.1 % range
        const double step = (mx - mn) / (double)nbins;
        double e = __dadd_rn(__dmul_rn((double)tid, step), mn);
        if (tid == nbins) e = mx;
        if (tid == 0) e = e - (mx - mn) * 0.001;
        edges[tid] = e;
    }
    __syncthreads();
    // pass 2: dispersion + bin id
    for (int g = tid; g < n_genes; g += HT) {
        const double s = gsum[g], ss = gsumsq[g];
        const double mean = s / n_total;
        const double var = (ss - mean * s) / (n_total - 1.0);
        z[g] = var / mean;                                   // 0/0 -> NaN like pandas
        int cnt = 0;
        for (int e = 0; e <= nbins; ++e) cnt += edges[e] < mean;     // searchsorted(side='left')
        int b = cnt - 1;
        b = b < 0 ? 0 : (b >= nbins ? nbins - 1 : b);
        bin[g] = b;
    }
    __syncthreads();
    // pass 3: per-bin mean / ddof-1 std of the dispersion, NaN-skipping, fixed-order sums
    for (int b = 0; b < nbins; ++b) {
        double c = 0.0, s = 0.0;
        for (int g = tid; g < n_genes; g += HT) {
            if (bin[g] == b) {
                const double d = z[g];
                if (!isnan(d)) { c += 1.0; s += d; }
            }
        }
        c = block_sum_f64(c, sm);
        s = block_sum_f64(s, sm);
        const double mu = s / c;
        double q = 0.0;
        for (int g = tid; g < n_genes; g += HT) {
            if (bin[g] == b) {
                const double d = z[g];
                if (!isnan(d)) q += (d - mu) * (d - mu);
            }
        }
        q = block_sum_f64(q, sm);
        if (tid == 0) {
            bmu[b] = mu;
            bsd[b] = c > 1.0 ? sqrt(q / (c - 1.0)) : NAN;
        }
    }
    __syncthreads();
    for (int g = tid; g < n_genes; g += HT) {
        const int b = bin[g];
        z[g] = (z[g] - bmu[b]) / bsd[b];
    }
}

__device__ __forceinline__ bool ranks_before(double zh, int h, double zg, int g) {
    const bool nh = isnan(zh), ng = isnan(zg);
    if (ng) return !nh || h < g;
    return !nh && (zh > zg || (zh == zg && h < g));
}

// rank of every gene's z among all genes by counting (ties by index).  Four threads per gene, each a quarter of every
// tile of the comparison range: 28k genes are only 110 CTAs of 256 genes - 6 warps per SM, 1.8 ms - but 438 CTAs this way.
__global__ void __launch_bounds__(256)
hvg_rank_kernel(const double *__restrict__ z, int n_genes, int hvg_n, int *__restrict__ ranked,
                int *__restrict__ selected) {
    __shared__ double tile[1024];
    const int g = blockIdx.x * 64 + (threadIdx.x >> 2), sub = threadIdx.x & 3;
    const double zg = g < n_genes ? z[g] : NAN;
    int rank = 0;
    for (int base = 0; base < n_genes; base += 1024) {
        __syncthreads();
        for (int i = threadIdx.x; i < 1024; i += 256) tile[i] = base + i < n_genes ? z[base + i] : NAN;
        __syncthreads();
        const int lim = min(1024, n_genes - base);
        for (int i = sub; i < lim; i += 4) rank += ranks_before(tile[i], base + i, zg, g);
    }
    rank += __shfl_xor_sync(AB_FULL, rank, 1);
    rank += __shfl_xor_sync(AB_FULL, rank, 2);
    if (g < n_genes && sub == 0) {
        const bool sel = rank < hvg_n;
        selected[g] = sel;
        if (sel) ranked[rank] = g;
    }
}

__global__ void __launch_bounds__(HT)
hvg_number_kernel(const double *__restrict__ z, int n_genes, int hvg_n, const int *__restrict__ ranked,
                  int *__restrict__ gene2hvg /* in: selected flag, out: column or -1 */, int *__restrict__ stats) {
    __shared__ int warp_tot[32];
    __shared__ int carry_sm;
    __shared__ int n_finite_sm, n_tie_sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { carry_sm = 0; n_finite_sm = 0; n_tie_sm = 0; }
    __syncthreads();
    const int n_sel = min(hvg_n, n_genes);
    const double zcut = z[ranked[n_sel - 1]];
    int finite = 0, tie = 0;
    for (int base = 0; base < n_genes; base += HT) {
        const int g = base + tid;
        const int flag = g < n_genes ? gene2hvg[g] : 0;
        if (g < n_genes) {
            const double v = z[g];
            finite += !isnan(v);
            tie += (!flag && v == zcut);
        }
        int inc = flag;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(AB_FULL, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < warp; ++w) woff += warp_tot[w];
        const int carry = carry_sm;
        if (g < n_genes) gene2hvg[g] = flag ? carry + woff + inc - 1 : -1;
        __syncthreads();
        if (tid == HT - 1) carry_sm = carry + woff + inc;
        __syncthreads();
    }
    atomicAdd(&n_finite_sm, finite);
    atomicAdd(&n_tie_sm, tie);
    __syncthreads();
    if (tid == 0) {
        stats[0] = n_finite_sm;
        stats[1] = n_sel;
        stats[2] = n_tie_sm;
        stats[3] = 0;
    }
}

extern "C" int ab_hvg_seurat(ab_ctx *ctx, const double *gsum, const double *gsumsq, int64_t n_total,
                             int32_t n_genes, int32_t nbins, int32_t hvg_n, int32_t *ranked, double *z,
                             int32_t *gene2hvg, int32_t *stats, void *stream) {
    AB_REQUIRE(ctx, "ab_hvg_seurat: null ctx");
    AB_REQUIRE(n_genes > 0 && n_total > 1, "ab_hvg_seurat: need n_genes>0 and n_total>1");
    AB_REQUIRE(nbins >= 1 && nbins <= MAX_BINS, "ab_hvg_seurat: nbins must be in 1..%d", MAX_BINS);
    AB_REQUIRE(hvg_n > 0, "ab_hvg_seurat: hvg_n must be positive (alona/alonabase.py:80-81)");
    cudaStream_t st = (cudaStream_t)stream;
    hvg_stats_kernel<<<1, HT, 0, st>>>(gsum, gsumsq, (double)n_total, n_genes, nbins, z, gene2hvg);
    AB_LAUNCH_CHECK(ctx);
    hvg_rank_kernel<<<(n_genes + 63) / 64, 256, 0, st>>>(z, n_genes, hvg_n, ranked, gene2hvg);
    AB_LAUNCH_CHECK(ctx);
    hvg_number_kernel<<<1, HT, 0, st>>>(z, n_genes, hvg_n, ranked, gene2hvg, stats);
    AB_LAUNCH_CHECK(ctx);
    return 0;
}
