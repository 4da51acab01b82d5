// Synthetic negative-binomial count matrices generated on the device (bench / tests only).
// Model of SURVEY.md §8(d): cell -> cluster, depth L_i ~ LogNormal(log Lbar, sigma);
// count(i,g) ~ NB(mean L_i * prof[cluster][g], inverse dispersion r) sampled by inversion of
// the NB CDF from one counter-based uniform per (seed, cell, gene).  Every value is a pure
// function of (seed, cell id, gene id): the matrix is identical for any sharding / GPU count.
#include "ab_common.cuh"
#include <math.h>
#include <algorithm>

static constexpr int SY_WARPS = 8;

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}

struct CellDraw { int cluster; float depth; uint64_t key; };

__device__ __forceinline__ CellDraw cell_draw(uint64_t seed, int64_t cell, int n_clusters, double lbar, double sigma) {
    CellDraw c;
    const uint64_t h = mix64(seed * 0x9e3779b97f4a7c15ull + (uint64_t)cell * 0xd1b54a32d192ed03ull + 0x1234567ull);
    c.key = mix64(h ^ 0xabcdef12345ull);
    c.cluster = (int)(h % (uint64_t)n_clusters);
    const uint64_t h2 = mix64(h + 0x632be59bd9b4e019ull);
    const double u1 = ((double)(h2 >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    const double u2 = ((double)(mix64(h2) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    const double z = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    c.depth = (float)exp(log(lbar) + sigma * z);
    return c;
}

// one NB draw by CDF inversion; mu = mean, r = inverse dispersion
__device__ __forceinline__ int nb_draw(float mu, float r, uint64_t key, int gene) {
    const uint64_t h = mix64(key + (uint64_t)gene * 0x9e3779b97f4a7c15ull);
    const double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);          // [0,1)
    const float lg = log1pf(mu / r);
    const float qnz = -expm1f(-r * lg);                                         // P(count > 0)
    if (u >= (double)qnz) return 0;
    float p = expf(-r * lg);                                                    // P(0)
    const float ratio = mu / (mu + r);
    double cdf = 0.0;
    int c = 0;
    while (c < 100000) {
        p *= (float)(c) + r;
        p /= (float)(c + 1);
        p *= ratio;
        ++c;
        cdf += (double)p;
        if (u < cdf) break;
    }
    return c;
}

template <bool FILL>
__global__ void __launch_bounds__(SY_WARPS * 32)
synth_kernel(const float *__restrict__ prof, int n_clusters, int n_genes, int64_t cell_begin, int64_t n_local,
             double lbar, double sigma, float r, uint64_t seed, int64_t *__restrict__ rowcount,
             const int64_t *__restrict__ rowptr, int32_t *__restrict__ col, int32_t *__restrict__ val) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t)blockIdx.x * SY_WARPS + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * SY_WARPS;
    for (int64_t row = wid; row < n_local; row += nw) {
        const CellDraw cd = cell_draw(seed, cell_begin + row, n_clusters, lbar, sigma);
        const float *pr = prof + (size_t)cd.cluster * n_genes;
        int64_t out = FILL ? rowptr[row] : 0;
        int total = 0;
        for (int g0 = 0; g0 < n_genes; g0 += 32) {
            const int g = g0 + lane;
            int c = 0;
            if (g < n_genes) c = nb_draw(cd.depth * __ldg(pr + g), r, cd.key, g);
            const unsigned m = __ballot_sync(AB_FULL, c > 0);
            if (FILL && c > 0) {
                const int64_t q = out + __popc(m & ((1u << lane) - 1));
                col[q] = g;
                val[q] = c;
            }
            out += __popc(m);
            total += __popc(m);
        }
        if (total == 0) {                    // every cell must have S_c >= 1
            if (FILL && lane == 0) { col[out] = (int)(cd.key % (uint64_t)n_genes); val[out] = 1; }
            total = 1;
        }
        if (!FILL && lane == 0) rowcount[row] = total;
    }
}

extern "C" size_t ab_synth_ws_bytes(int64_t n_local) { return ab_scan_ws_bytes(n_local) + 256; }

extern "C" int ab_synth_nb_count(ab_ctx *ctx, const float *prof, int32_t n_clusters, int32_t n_genes,
                                 int64_t cell_begin, int64_t n_local, double lbar, double depth_sigma, double r_disp,
                                 uint64_t seed, int64_t *rowptr, int64_t *h_nnz, void *ws, size_t ws_bytes,
                                 void *stream) {
    AB_REQUIRE(ctx, "ab_synth_nb_count: null ctx");
    AB_REQUIRE(n_clusters > 0 && n_genes > 0 && n_local >= 0, "ab_synth_nb_count: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_local > 0) {
        int64_t blocks = std::min<int64_t>((n_local + SY_WARPS - 1) / SY_WARPS, (int64_t)ctx->sm_count * 16);
        synth_kernel<false><<<(unsigned)blocks, SY_WARPS * 32, 0, st>>>(prof, n_clusters, n_genes, cell_begin, n_local, lbar,
                                                                         depth_sigma, (float)r_disp, seed, rowptr, nullptr,
                                                                         nullptr, nullptr);
        AB_LAUNCH_CHECK(ctx);
    }
    if (ab_exclusive_scan_i64(ctx, rowptr, rowptr, n_local, ws, ws_bytes, st)) return 1;
    if (h_nnz) {
        AB_CUDA(cudaMemcpyAsync(h_nnz, rowptr + n_local, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        AB_CUDA(cudaStreamSynchronize(st));
    }
    return 0;
}

extern "C" int ab_synth_nb_fill(ab_ctx *ctx, const float *prof, int32_t n_clusters, int32_t n_genes, int64_t cell_begin,
                                int64_t n_local, double lbar, double depth_sigma, double r_disp, uint64_t seed,
                                const int64_t *rowptr, int32_t *col, int32_t *val, void *stream) {
    AB_REQUIRE(ctx, "ab_synth_nb_fill: null ctx");
    if (n_local == 0) return 0;
    int64_t blocks = std::min<int64_t>((n_local + SY_WARPS - 1) / SY_WARPS, (int64_t)ctx->sm_count * 16);
    synth_kernel<true><<<(unsigned)blocks, SY_WARPS * 32, 0, (cudaStream_t)stream>>>(
        prof, n_clusters, n_genes, cell_begin, n_local, lbar, depth_sigma, (float)r_disp, seed, nullptr, rowptr, col, val);
    AB_LAUNCH_CHECK(ctx);
    return 0;
}
