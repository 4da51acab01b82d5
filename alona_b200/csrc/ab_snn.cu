// K10: shared-nearest-neighbour graph from the kNN table.
// Replaces AlonaClustering.snn (alona/clustering.py:204-231):
//   A = kNN incidence (self included), C = A.A^T (integer overlaps, scipy csr_matmat),
//   keep (i,m) iff C[i,m]/(k+(k-C[i,m])) > prune_snn, emit in csr_matmat's order.
//
// C[i,m] = |N(i) n N(m)|.  The nonzeros of row i are the cells m met on the 2-hop walk
//   for j in sorted(N(i)):  for m in sorted(R(j)):      R(j) = { m : j in N(m) }  (reverse lists)
// and scipy's order inside the row is the REVERSE of first-encounter order of that walk.
//
// Round 1 de-duplicated the walk with per-row hash tables in shared memory (five row classes by walk length, up to
// 128 KB of table per row; 68 ms at C3, bound by dependent shared-memory atomics and per-row latency).  This version
// needs no table of the walk at all - it is the sorted-neighbour-list intersection the brief describes:
//   * one warp per cell row i; the row's sorted neighbours go into a small open-addressing hash (cell id -> position
//     a in the sorted row) in shared memory: 2k..4k slots, built once per row;
//   * the walk is enumerated by position p (lane <-> position, coalesced reads of the reverse lists); for the entry
//     (a, m) the lane scans N(m) (k ids, one row of the kNN table, L2 resident) through that hash: the number of hits
//     IS the overlap |N(i) n N(m)| and the smallest hit position a* is where the walk meets m first, so the entry is
//     a first encounter iff a* == a.  No atomics, no inter-lane dependence, any walk length, any in-degree;
//   * positions are visited from the last to the first, so first encounters leave in scipy's order.
// Work: one k-id row read + k hash probes per walk entry (C3: 2.4e9 entries x 80 B of L2 traffic).
#include "ab_common.cuh"
#include <limits.h>
#include <algorithm>

static constexpr int SNN_MAXK = 128;
static constexpr int SORT_CTA_MAX = 4096;
static constexpr int SNN_WARPS = 16;              // warps per CTA of the intersection kernel
// staged edges per row (single-pass protocol); rows with more take a second pass.  Edges per row grow with k
// (C3, k = 20: 21 on average; a C5-shaped run, k = 30: 324), the slot is sized generously: it is only ever read up
// to the row's count.
static inline int snn_stage_cap(int k) { return std::min(1024, std::max(64, 16 * k)); }

// ---------------------------------------------------------------------------------------
// reverse graph
// ---------------------------------------------------------------------------------------
__global__ void snn_degree_kernel(const int32_t *__restrict__ knn, int64_t total, int64_t n_total,
                                  int64_t *__restrict__ deg, int *__restrict__ bad) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < total) {
        int j = knn[e];
        if (j < 0 || j >= n_total) { atomicAdd(bad, 1); return; }
        atomicAdd((unsigned long long *)(deg + j), 1ull);
    }
}

__global__ void snn_scatter_kernel(const int32_t *__restrict__ knn, int64_t total, int k,
                                   const int64_t *__restrict__ rptr, int *__restrict__ cursor,
                                   int32_t *__restrict__ rlist) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < total) {
        int j = knn[e];
        int pos = atomicAdd(cursor + j, 1);
        rlist[rptr[j] + pos] = (int32_t)(e / k);
    }
}

__device__ __forceinline__ int warp_bitonic_sort(int v, int lane) {
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            int other = __shfl_xor_sync(AB_FULL, v, stride);
            bool up = ((lane & size) == 0);
            bool lower = ((lane & stride) == 0);
            int mn = min(v, other), mx = max(v, other);
            v = (lower == up) ? mn : mx;
        }
    }
    return v;
}

// lists of <= 32 entries: one warp each, register bitonic sort.  Longer lists are queued.
__global__ void __launch_bounds__(256)
snn_sort_short_kernel(const int64_t *__restrict__ rptr, int32_t *__restrict__ rlist, int64_t n_total,
                      int32_t *__restrict__ long_ids, int *__restrict__ n_long) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * 8;
    for (int64_t j = w; j < n_total; j += nw) {
        const int64_t a = rptr[j];
        const int len = (int)(rptr[j + 1] - a);
        if (len <= 1) continue;
        if (len > 32) {
            if (lane == 0) long_ids[atomicAdd(n_long, 1)] = (int32_t)j;
            continue;
        }
        int v = lane < len ? rlist[a + lane] : INT_MAX;
        v = warp_bitonic_sort(v, lane);
        if (lane < len) rlist[a + lane] = v;
    }
}

// lists of 33..4096 entries: CTA bitonic sort in shared memory; longer: counting rank via tmp.
__global__ void __launch_bounds__(256)
snn_sort_long_kernel(const int64_t *__restrict__ rptr, int32_t *__restrict__ rlist,
                     const int32_t *__restrict__ long_ids, const int *__restrict__ n_long,
                     int32_t *__restrict__ tmp) {
    __shared__ int sm[SORT_CTA_MAX];
    const int nl = *n_long;
    for (int q = blockIdx.x; q < nl; q += gridDim.x) {
        const int j = long_ids[q];
        const int64_t a = rptr[j];
        const int64_t len = rptr[j + 1] - a;
        if (len <= SORT_CTA_MAX) {
            int n2 = 64;
            while (n2 < len) n2 <<= 1;
            for (int i = threadIdx.x; i < n2; i += 256) sm[i] = i < len ? rlist[a + i] : INT_MAX;
            __syncthreads();
            for (int size = 2; size <= n2; size <<= 1) {
                for (int stride = size >> 1; stride > 0; stride >>= 1) {
                    for (int i = threadIdx.x; i < (n2 >> 1); i += 256) {
                        int lo = ((i & ~(stride - 1)) << 1) | (i & (stride - 1));
                        int hi = lo | stride;
                        bool up = ((lo & size) == 0);
                        int x = sm[lo], y = sm[hi];
                        if ((x > y) == up) { sm[lo] = y; sm[hi] = x; }
                    }
                    __syncthreads();
                }
            }
            for (int i = threadIdx.x; i < len; i += 256) rlist[a + i] = sm[i];
            __syncthreads();
        } else {
            // entries of one reverse list are distinct cells -> rank = #smaller
            for (int64_t i = threadIdx.x; i < len; i += 256) {
                const int v = rlist[a + i];
                int64_t r = 0;
                for (int64_t t = 0; t < len; ++t) r += rlist[a + t] < v;
                tmp[a + r] = v;
            }
            __syncthreads();
            for (int64_t i = threadIdx.x; i < len; i += 256) rlist[a + i] = tmp[a + i];
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------
// per-row state of a warp: sorted neighbours, their reverse-list extents and walk positions, membership hash
// ---------------------------------------------------------------------------------------
struct WarpRowState {
    int nb[SNN_MAXK];            // sorted neighbour ids
    int base[SNN_MAXK + 1];      // walk position where each neighbour's reverse list starts
    long long start[SNN_MAXK];   // rptr of each neighbour
};

// Loads and sorts row `grow`, fills st; returns the walk length.  One warp.
__device__ __forceinline__ long long warp_load_row(const int32_t *__restrict__ knn, int k, int64_t grow,
                                                   const int64_t *__restrict__ rptr, WarpRowState &st, int lane) {
    // sort k <= 128 ids: rank by counting through shared memory
    for (int c = lane; c < k; c += 32) st.base[c] = knn[grow * k + c];       // base[] as scratch
    __syncwarp();
    for (int c = lane; c < k; c += 32) {
        const int v = st.base[c];
        int rank = 0;
        for (int t = 0; t < k; ++t) {
            const int u = st.base[t];
            rank += (u < v) || (u == v && t < c);
        }
        st.nb[rank] = v;
    }
    __syncwarp();
    // degrees and exclusive prefix
    long long run = 0;
    for (int c0 = 0; c0 < k; c0 += 32) {
        const int c = c0 + lane;
        int d = 0;
        if (c < k) {
            const int j = st.nb[c];
            const long long s = rptr[j];
            st.start[c] = s;
            d = (int)(rptr[j + 1] - s);
        }
        long long inc = d;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long t = __shfl_up_sync(AB_FULL, inc, o);
            if (lane >= o) inc += t;
        }
        __syncwarp();
        if (c < k) st.base[c] = (int)min(run + inc - d, (long long)INT_MAX);
        run += __shfl_sync(AB_FULL, inc, 31);
    }
    if (lane == 0) st.base[k] = (int)min(run, (long long)INT_MAX);
    __syncwarp();
    return run;
}

__device__ __forceinline__ unsigned snn_hash(int key, int lg) { return ((unsigned)key * 2654435761u) >> (32 - lg); }

// membership probe: position of cell x in the row's sorted neighbour list, or -1
__device__ __forceinline__ int snn_probe(const int *__restrict__ hkeys, const int *__restrict__ hvals, int S, int lg, int x) {
    unsigned s = snn_hash(x, lg);
    int key = hkeys[s];
    while (key != x && key != -1) { s = (s + 1) & (S - 1); key = hkeys[s]; }
    return key == x ? hvals[s] : -1;
}

// ---------------------------------------------------------------------------------------
// the intersection walk.  MODE 0: stage every row's edges (emission order included) into its slot of `cap`
// entries and write the row's edge count (rows are taken from a global ticket counter: walk lengths differ by two
// orders of magnitude).  MODE 1: write the rows listed in `rows` (those whose edges overflowed their slot) straight
// to their final positions given by rowptr_out.
// GL lanes share one walk entry (a, m): together they read the k ids of N(m) as 16-byte pieces - one contiguous
// 4k-byte segment per entry instead of k scattered 4-byte reads per lane, which is what bounds a lane-per-entry
// version (measured: 163 ms at C3, the L1 processes one 32-byte sector per cycle) - and probe the row's hash; the
// hits are summed inside the group.  An entry is emitted iff overlap >= v_min and it is the first encounter of m
// (the smallest hit position equals a); first encounters leave in descending walk position = scipy's order.
// The number of nonzeros of A.A^T (the reference's "links pruned" statistic) needs no de-duplication either: an
// (i, m) pair with overlap c is met exactly c times, so it is  #{c = 1} + #{c = 2}/2 + #{first encounters with c >= 3}.
// ---------------------------------------------------------------------------------------
template <int MODE, int GL>
__global__ void __launch_bounds__(SNN_WARPS * 32)
snn_isect_kernel(const int32_t *__restrict__ knn, int k, int lg, int64_t row_begin, int64_t n_rows,
                 const int32_t *__restrict__ rows, const int *__restrict__ n_listed, const int64_t *__restrict__ rptr,
                 const int32_t *__restrict__ rlist, int v_min, int64_t *__restrict__ rowcount,
                 const int64_t *__restrict__ rowptr_out, int32_t *__restrict__ src, int32_t *__restrict__ dst,
                 uint8_t *__restrict__ overlap, int cap, unsigned long long *__restrict__ ticket,
                 unsigned long long *__restrict__ walk_stats /* [0] max walk, [1] sum of walks, [2] nonzeros of A.A^T */,
                 int *__restrict__ n_dup) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int EP = 32 / GL;                                  // walk entries per step
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane / GL, gl = lane % GL;
    WarpRowState &st = reinterpret_cast<WarpRowState *>(smem_raw)[warp];
    const int S = 1 << lg;
    int *hkeys = reinterpret_cast<int *>(smem_raw + sizeof(WarpRowState) * SNN_WARPS) + (size_t)warp * 2 * S;
    int *hvals = hkeys + S;
    const int64_t total = MODE == 0 ? n_rows : (int64_t)*n_listed;
    const bool vec = (k & 3) == 0;                               // rows of the table are 16-byte aligned
    unsigned long long wmax = 0, wsum = 0, n1 = 0, n2 = 0, n3 = 0;
    int dups = 0;
    while (true) {
        unsigned long long q = 0;
        if (lane == 0) q = atomicAdd(ticket, 1ull);
        q = __shfl_sync(AB_FULL, q, 0);
        if ((int64_t)q >= total) break;
        const int64_t r = MODE == 0 ? (int64_t)q : (int64_t)rows[q];
        const int64_t grow = row_begin + r;
        const long long walk_ll = warp_load_row(knn, k, grow, rptr, st, lane);
        for (int s = lane; s < S; s += 32) hkeys[s] = -1;
        __syncwarp();
        for (int a = lane; a < k; a += 32) {                             // distinct ids: plain CAS insertion
            const int key = st.nb[a];
            unsigned s = snn_hash(key, lg);
            while (atomicCAS(hkeys + s, -1, key) != -1) s = (s + 1) & (S - 1);
            hvals[s] = a;
        }
        __syncwarp();
        if (MODE == 0) {
            wmax = max(wmax, (unsigned long long)walk_ll);
            wsum += (unsigned long long)walk_ll;
            if (lane == 0 && knn[grow * k] != (int32_t)grow) ++dups;
        }
        const int64_t obase = MODE == 0 ? r * (int64_t)cap : rowptr_out[r];
        int64_t out = obase;
        for (int a = k - 1; a >= 0; --a) {
            const int len = st.base[a + 1] - st.base[a];
            const int32_t *rl = rlist + st.start[a];
            for (int t0 = ((len - 1) / EP) * EP; t0 >= 0; t0 -= EP) {
                const int e = t0 + g;
                const bool valid = e < len;
                const int m = valid ? __ldg(rl + e) : 0;
                // this lane's up to four ids of N(m)
                int x0 = -2, x1 = -2, x2 = -2, x3 = -2;                  // -2 is never a key
                if (valid && 4 * gl < k) {
                    const int32_t *nm = knn + (int64_t)m * k + 4 * gl;
                    if (vec) {
                        const int4 v = __ldg(reinterpret_cast<const int4 *>(nm));
                        x0 = v.x; x1 = v.y; x2 = v.z; x3 = v.w;
                    } else {
                        x0 = __ldg(nm);
                        if (4 * gl + 1 < k) x1 = __ldg(nm + 1);
                        if (4 * gl + 2 < k) x2 = __ldg(nm + 2);
                        if (4 * gl + 3 < k) x3 = __ldg(nm + 3);
                    }
                }
                const int p0 = snn_probe(hkeys, hvals, S, lg, x0), p1 = snn_probe(hkeys, hvals, S, lg, x1);
                const int p2 = snn_probe(hkeys, hvals, S, lg, x2), p3 = snn_probe(hkeys, hvals, S, lg, x3);
                int cnt = (p0 >= 0) + (p1 >= 0) + (p2 >= 0) + (p3 >= 0);
#pragma unroll
                for (int o = GL / 2; o > 0; o >>= 1) cnt += __shfl_xor_sync(AB_FULL, cnt, o);
                bool keep = false;
                if (__any_sync(AB_FULL, cnt >= min(v_min, 3))) {           // rare: first-encounter test
                    int amin = INT_MAX;
                    if (p0 >= 0) amin = p0;
                    if (p1 >= 0) amin = min(amin, p1);
                    if (p2 >= 0) amin = min(amin, p2);
                    if (p3 >= 0) amin = min(amin, p3);
#pragma unroll
                    for (int o = GL / 2; o > 0; o >>= 1) amin = min(amin, __shfl_xor_sync(AB_FULL, amin, o));
                    const bool first = valid && amin == a;
                    keep = first && cnt >= v_min && gl == 0;
                    if (MODE == 0 && first && cnt >= 3 && gl == 0) ++n3;
                }
                if (MODE == 0 && valid && gl == 0) { n1 += cnt == 1; n2 += cnt == 2; }
                const unsigned b = __ballot_sync(AB_FULL, keep);
                if (b) {
                    if (keep) {
                        const int64_t o = out + __popc(b & ~((2u << lane) - 1));     // later walk positions first
                        if (MODE == 1 || o - obase < cap) {
                            dst[o] = m;
                            if (MODE == 1) src[o] = (int32_t)grow;
                            if (overlap) overlap[o] = (uint8_t)min(cnt, 255);
                        }
                    }
                    out += __popc(b);
                }
            }
        }
        if (MODE == 0 && lane == 0) rowcount[r] = out - obase;
        __syncwarp();
    }
    if (MODE == 0) {
        n1 = (unsigned long long)ab_warp_sum_i64((long long)n1);
        n2 = (unsigned long long)ab_warp_sum_i64((long long)n2);
        n3 = (unsigned long long)ab_warp_sum_i64((long long)n3);
        if (lane == 0) {
            atomicMax(walk_stats, wmax);
            atomicAdd(walk_stats + 1, wsum);
            atomicAdd(walk_stats + 2, n1 + n2 / 2 + n3);
            if (dups) atomicAdd(n_dup, dups);
        }
    }
}

// staging -> final edge list: one warp per row; rows whose edges did not fit their staging slot are queued
__global__ void __launch_bounds__(256)
snn_compact_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ stage_dst,
                   const uint8_t *__restrict__ stage_ov, int cap, int64_t n_rows, int64_t row_begin,
                   int32_t *__restrict__ src, int32_t *__restrict__ dst, uint8_t *__restrict__ overlap,
                   int32_t *__restrict__ ovf_rows, int *__restrict__ n_ovf) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * 8;
    for (int64_t r = w; r < n_rows; r += nw) {
        const int64_t a = rowptr[r];
        const int64_t cnt = rowptr[r + 1] - a;
        if (cnt > cap) {
            if (lane == 0) ovf_rows[atomicAdd(n_ovf, 1)] = (int32_t)r;
            continue;
        }
        const int grow = (int)(row_begin + r);
        for (int i = lane; i < (int)cnt; i += 32) {
            dst[a + i] = stage_dst[r * cap + i];
            src[a + i] = grow;
            if (overlap) overlap[a + i] = stage_ov[r * cap + i];
        }
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
struct SnnWs {
    int64_t *deg;        // n_total+1 (becomes rptr after the scan)
    int *cursor;         // n_total
    int32_t *rlist;      // n_total*k
    int32_t *tmp;        // n_total*k
    int32_t *long_ids;   // n_total
    int32_t *stage_dst;  // n_rows * snn_stage_cap(k): edges of each row, staged by the (single) walk pass
    uint8_t *stage_ov;   // n_rows * snn_stage_cap(k)
    int32_t *ovf_rows;   // n_rows: rows whose edges did not fit their staging slot
    int *counters;       // 16 ints: [0] n_long, [2] n_dup, [3] bad, [10] overflow rows
    unsigned long long *walk_stats;   // [0] max walk, [1] sum of walks, [2] nonzeros of A.A^T, [3] row ticket
    int64_t *rowcount;   // n_rows+1
    void *scan_ws;
    size_t scan_bytes;
};

static size_t snn_layout(int64_t n_total, int32_t k, int64_t n_rows, void *ws, size_t ws_bytes, SnnWs *out) {
    ab_arena a(ws ? ws : (void *)256, ws ? ws_bytes : (size_t)-1 / 2);
    SnnWs w;
    const size_t nr = (size_t)(n_rows > 0 ? n_rows : 1);
    w.deg = a.take<int64_t>(n_total + 1);
    w.cursor = a.take<int>(n_total);
    w.rlist = a.take<int32_t>((size_t)n_total * k);
    w.tmp = a.take<int32_t>((size_t)n_total * k);
    w.long_ids = a.take<int32_t>(n_total);
    w.stage_dst = a.take<int32_t>(nr * snn_stage_cap(k));
    w.stage_ov = a.take<uint8_t>(nr * snn_stage_cap(k));
    w.ovf_rows = a.take<int32_t>(nr);
    w.counters = a.take<int>(16);
    w.walk_stats = a.take<unsigned long long>(4);
    w.rowcount = a.take<int64_t>(n_rows + 1);
    int64_t big = n_total > n_rows ? n_total : n_rows;
    w.scan_bytes = ab_scan_ws_bytes(big + 1);
    w.scan_ws = a.take<char>(w.scan_bytes);
    if (out) *out = w;
    return a.ok ? a.off : 0;
}

extern "C" size_t ab_snn_ws_bytes(int64_t n_total, int32_t k, int64_t n_rows) {
    return snn_layout(n_total, k, n_rows, nullptr, 0, nullptr) + 256;
}

static int snn_hash_lg(int k) {
    int lg = 5;
    while ((1 << lg) < 8 * k) ++lg;                        // load factor <= 1/8: an absent id stops at its first slot 7 times in 8
    return lg;
}
static size_t snn_isect_smem(int k) {
    return sizeof(WarpRowState) * SNN_WARPS + (size_t)SNN_WARPS * 2 * ((size_t)1 << snn_hash_lg(k)) * sizeof(int);
}

extern "C" int ab_snn_count(ab_ctx *ctx, const int32_t *knn_all, int64_t n_total, int32_t k, int64_t row_begin,
                            int64_t row_end, int32_t v_min, int64_t *rowptr, int64_t *h_edges, int64_t *stats,
                            void *ws, size_t ws_bytes, void *stream) {
    AB_REQUIRE(ctx, "ab_snn_count: null ctx");
    AB_REQUIRE(k >= 1 && k <= SNN_MAXK, "ab_snn_count: k must be in 1..%d", SNN_MAXK);
    AB_REQUIRE(n_total >= 1 && n_total * (int64_t)k < INT_MAX, "ab_snn_count: n_total * k must stay below 2^31");
    AB_REQUIRE(row_begin >= 0 && row_end >= row_begin && row_end <= n_total, "ab_snn_count: bad row range");
    const int64_t n_rows = row_end - row_begin;
    SnnWs w;
    size_t need = snn_layout(n_total, k, n_rows, ws, ws_bytes, &w);
    AB_REQUIRE(need != 0, "ab_snn_count: workspace too small (need %zu)", ab_snn_ws_bytes(n_total, k, n_rows));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = n_total * k;
    AB_CUDA(cudaMemsetAsync(w.deg, 0, (n_total + 1) * sizeof(int64_t), st));
    AB_CUDA(cudaMemsetAsync(w.cursor, 0, n_total * sizeof(int), st));
    AB_CUDA(cudaMemsetAsync(w.counters, 0, 16 * sizeof(int), st));
    AB_CUDA(cudaMemsetAsync(w.walk_stats, 0, 4 * sizeof(unsigned long long), st));
    const unsigned eb = (unsigned)((total + 255) / 256);
    ab_prof_begin(ctx, AB_PROF_SNN_BUILD, st);
    snn_degree_kernel<<<eb, 256, 0, st>>>(knn_all, total, n_total, w.deg, w.counters + 3);
    AB_LAUNCH_CHECK(ctx);
    if (ab_exclusive_scan_i64(ctx, w.deg, w.deg, n_total, w.scan_ws, w.scan_bytes, st)) return 1;
    int h_bad = 0;
    AB_CUDA(cudaMemcpyAsync(&h_bad, w.counters + 3, sizeof(int), cudaMemcpyDeviceToHost, st));
    AB_CUDA(cudaStreamSynchronize(st));
    AB_REQUIRE(h_bad == 0, "ab_snn_count: kNN table holds %d indices outside [0,%lld)", h_bad, (long long)n_total);
    snn_scatter_kernel<<<eb, 256, 0, st>>>(knn_all, total, k, w.deg, w.cursor, w.rlist);
    AB_LAUNCH_CHECK(ctx);
    snn_sort_short_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(w.deg, w.rlist, n_total, w.long_ids, w.counters + 0);
    AB_LAUNCH_CHECK(ctx);
    snn_sort_long_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(w.deg, w.rlist, w.long_ids, w.counters + 0, w.tmp);
    AB_LAUNCH_CHECK(ctx);
    ab_prof_end(ctx, AB_PROF_SNN_BUILD, st);
    ab_prof_begin(ctx, AB_PROF_SNN_COUNT, st);
    if (n_rows > 0) {
        // ONE walk per row: every row's edges (emission order included) are staged in a per-row slot;
        // ab_snn_fill then only compacts the slots into the final list
        const size_t smem = snn_isect_smem(k);
        const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (size_t)200 * 1024 / smem));
        const unsigned grid = (unsigned)std::min<int64_t>((n_rows + SNN_WARPS - 1) / SNN_WARPS, (int64_t)ctx->sm_count * per_sm);
#define AB_SNN_ISECT0(GL)                                                                                             \
        do {                                                                                                          \
            AB_CUDA(cudaFuncSetAttribute(snn_isect_kernel<0, GL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            snn_isect_kernel<0, GL><<<grid, SNN_WARPS * 32, smem, st>>>(                                              \
                knn_all, k, snn_hash_lg(k), row_begin, n_rows, nullptr, nullptr, w.deg, w.rlist, v_min, w.rowcount, nullptr, \
                nullptr, w.stage_dst, w.stage_ov, snn_stage_cap(k), w.walk_stats + 3, w.walk_stats, w.counters + 2);  \
        } while (0)
        if (k <= 32) AB_SNN_ISECT0(8); else if (k <= 64) AB_SNN_ISECT0(16); else AB_SNN_ISECT0(32);
#undef AB_SNN_ISECT0
        AB_LAUNCH_CHECK(ctx);
    }
    if (ab_exclusive_scan_i64(ctx, w.rowcount, rowptr, n_rows, w.scan_ws, w.scan_bytes, st)) return 1;
    ab_prof_end(ctx, AB_PROF_SNN_COUNT, st);
    int h_cnt[12];
    unsigned long long h_walk[4] = {0, 0, 0, 0};
    AB_CUDA(cudaMemcpyAsync(h_cnt, w.counters, 12 * sizeof(int), cudaMemcpyDeviceToHost, st));
    AB_CUDA(cudaMemcpyAsync(h_walk, w.walk_stats, sizeof(h_walk), cudaMemcpyDeviceToHost, st));
    if (h_edges) AB_CUDA(cudaMemcpyAsync(h_edges, rowptr + n_rows, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    AB_CUDA(cudaStreamSynchronize(st));
    if (stats) {
        // host array, see include/alona_b200.h
        stats[0] = 0;
        stats[1] = (int64_t)h_walk[0];
        stats[2] = h_cnt[2];
        stats[3] = 0;
        stats[4] = (int64_t)h_walk[1];
        stats[5] = (int64_t)h_walk[2];
        stats[6] = 0;
        stats[7] = 0;
    }
    return 0;
}

extern "C" int ab_snn_fill(ab_ctx *ctx, const int32_t *knn_all, int64_t n_total, int32_t k, int64_t row_begin,
                           int64_t row_end, int32_t v_min, const int64_t *rowptr, int32_t *src, int32_t *dst,
                           uint8_t *overlap, void *ws, size_t ws_bytes, void *stream) {
    AB_REQUIRE(ctx, "ab_snn_fill: null ctx");
    const int64_t n_rows = row_end - row_begin;
    SnnWs w;
    size_t need = snn_layout(n_total, k, n_rows, ws, ws_bytes, &w);
    AB_REQUIRE(need != 0, "ab_snn_fill: workspace too small");
    if (n_rows == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    ab_prof_begin(ctx, AB_PROF_SNN_FILL, st);
    AB_CUDA(cudaMemsetAsync(w.counters + 10, 0, sizeof(int), st));
    AB_CUDA(cudaMemsetAsync(w.walk_stats + 3, 0, sizeof(unsigned long long), st));
    snn_compact_kernel<<<(unsigned)std::min<int64_t>((n_rows + 7) / 8, (int64_t)ctx->sm_count * 16), 256, 0, st>>>(
        rowptr, w.stage_dst, w.stage_ov, snn_stage_cap(k), n_rows, row_begin, src, dst, overlap, w.ovf_rows, w.counters + 10);
    AB_LAUNCH_CHECK(ctx);
    // rows whose edges overflowed their staging slot (rare) walk once more, straight into their final positions
    const size_t smem = snn_isect_smem(k);
#define AB_SNN_ISECT1(GL)                                                                                             \
    do {                                                                                                              \
        AB_CUDA(cudaFuncSetAttribute(snn_isect_kernel<1, GL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        snn_isect_kernel<1, GL><<<ctx->sm_count, SNN_WARPS * 32, smem, st>>>(                                         \
            knn_all, k, snn_hash_lg(k), row_begin, n_rows, w.ovf_rows, w.counters + 10, w.deg, w.rlist, v_min, nullptr, rowptr, \
            src, dst, overlap, 0, w.walk_stats + 3, w.walk_stats, nullptr);                                            \
    } while (0)
    if (k <= 32) AB_SNN_ISECT1(8); else if (k <= 64) AB_SNN_ISECT1(16); else AB_SNN_ISECT1(32);
#undef AB_SNN_ISECT1
    AB_LAUNCH_CHECK(ctx);
    ab_prof_end(ctx, AB_PROF_SNN_FILL, st);
    return 0;
}
