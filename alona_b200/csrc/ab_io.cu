// Edge-list hand-off: formats the SNN edge list as the bytes alona writes to
// csvs/snn_graph.csv  (alona/clustering.py:237-238:  df.to_csv(snn_path, header=None, index=None), i.e. one
// "source,target\n" line per edge, decimal, no padding).  At C3 the list has 2.7e7 edges; pandas needs tens of
// seconds for what is a few milliseconds of device work plus one D2H copy.
//
// count: bytes of every line -> exclusive scan -> total;  fill: digits written back to front, one thread per edge.
#include "ab_common.cuh"

__device__ __forceinline__ int dec_digits(unsigned v) {
    int n = 1;
    n += v >= 10u; n += v >= 100u; n += v >= 1000u; n += v >= 10000u; n += v >= 100000u;
    n += v >= 1000000u; n += v >= 10000000u; n += v >= 100000000u; n += v >= 1000000000u;
    return n;
}

__global__ void csv_len_kernel(const int32_t *__restrict__ src, const int32_t *__restrict__ dst, int64_t n,
                               int64_t *__restrict__ len, int *__restrict__ bad) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int s = src[e], d = dst[e];
    if (s < 0 || d < 0) { atomicAdd(bad, 1); len[e] = 0; return; }
    len[e] = dec_digits((unsigned)s) + dec_digits((unsigned)d) + 2;
}

__device__ __forceinline__ void put_dec(char *end, unsigned v) {      // writes the digits ending just before `end`
    do {
        *--end = (char)('0' + v % 10u);
        v /= 10u;
    } while (v);
}

__global__ void csv_fill_kernel(const int32_t *__restrict__ src, const int32_t *__restrict__ dst, int64_t n,
                                const int64_t *__restrict__ off, char *__restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const unsigned s = (unsigned)src[e], d = (unsigned)dst[e];
    char *p = out + off[e];
    const int ns = dec_digits(s), nd = dec_digits(d);
    put_dec(p + ns, s);
    p[ns] = ',';
    put_dec(p + ns + 1 + nd, d);
    p[ns + 1 + nd] = '\n';
}

extern "C" size_t ab_edges_csv_ws_bytes(int64_t n_edges) { return ab_scan_ws_bytes(n_edges + 1) + 512; }

extern "C" int ab_edges_csv_count(ab_ctx *ctx, const int32_t *src, const int32_t *dst, int64_t n_edges, int64_t *offsets,
                                  int64_t *h_bytes, void *ws, size_t ws_bytes, void *stream) {
    AB_REQUIRE(ctx, "ab_edges_csv_count: null ctx");
    AB_REQUIRE(n_edges >= 0, "ab_edges_csv_count: bad edge count");
    AB_REQUIRE(ws_bytes >= ab_edges_csv_ws_bytes(n_edges), "ab_edges_csv_count: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    int *bad = (int *)ws;
    void *scan_ws = (char *)ws + 256;
    AB_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), st));
    if (n_edges > 0) {
        csv_len_kernel<<<(unsigned)((n_edges + 255) / 256), 256, 0, st>>>(src, dst, n_edges, offsets, bad);
        AB_LAUNCH_CHECK(ctx);
    }
    if (ab_exclusive_scan_i64(ctx, offsets, offsets, n_edges, scan_ws, ws_bytes - 256, st)) return 1;
    int h_bad = 0;
    AB_CUDA(cudaMemcpyAsync(&h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    if (h_bytes) AB_CUDA(cudaMemcpyAsync(h_bytes, offsets + n_edges, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    AB_CUDA(cudaStreamSynchronize(st));
    AB_REQUIRE(h_bad == 0, "ab_edges_csv_count: %d negative node ids", h_bad);
    return 0;
}

extern "C" int ab_edges_csv_fill(ab_ctx *ctx, const int32_t *src, const int32_t *dst, int64_t n_edges,
                                 const int64_t *offsets, char *out, void *stream) {
    AB_REQUIRE(ctx, "ab_edges_csv_fill: null ctx");
    if (n_edges == 0) return 0;
    csv_fill_kernel<<<(unsigned)((n_edges + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, dst, n_edges, offsets, out);
    AB_LAUNCH_CHECK(ctx);
    return 0;
}
