// K10: shared-nearest-neighbour graph from the kNN table.
// Replaces AlonaClustering.snn (alona/clustering.py:204-231):
//   A = kNN incidence (self included), C = A.A^T (integer overlaps, scipy csr_matmat),
//   keep (i,m) iff C[i,m]/(k+(k-C[i,m])) > prune_snn, emit in csr_matmat's order.
//
// C[i,m] = |N(i) n N(m)| = #{ j in N(i) : m in R(j) },  R(j) = { m : j in N(m) } (reverse lists).
// scipy's order inside row i is the REVERSE of first-encounter order when scanning j over
// sorted(N(i)) and m over sorted(R(j)).  So, per row, the 2-hop walk is
// enumerated by position p = base(j) + t; a per-warp shared-memory hash keeps for every
// distinct m its overlap count and the position of its first encounter; a second walk emits
// the first encounters from the last position to the first, filtered by overlap >= v_min.
// No sort, no atomics in the common path (duplicates inside a 32-wide step are merged with
// match.any, steps are processed in ascending position order).
//
// Rows are classified by walk length: short walks take one warp with a per-warp table, longer
// ones one CTA with a shared-memory table sized to the walk (the common case at scale: kNN graphs
// in 50 dimensions have hubs), and anything beyond shared memory one CTA with dense count/first
// arrays in global scratch (always correct, any in-degree).
#include "ab_common.cuh"
#include <limits.h>
#include <stdlib.h>
#include <algorithm>

static constexpr int SNN_WARPS = 8;
static constexpr int SNN_SLOTS = 512;             // per warp, power of two (walk <= 350: load factor <= 0.69)
static constexpr int SNN_SLOTS_LG = 9;
static constexpr int SNN_WALK_LIMIT = 350;        // rows with a longer walk take a whole CTA (measured at C3: limit 1400 -> 68.7 ms,
                                                  // 700 -> 52.7, 350 -> 51.9, 0 -> 51.9: the one-warp kernel holds 8 warps per SM)
static constexpr int SNN_MAXK = 128;
static constexpr int SNN_BIG_CTAS = 296;         // two per SM (each owns 2*n_total ints of global scratch)
static constexpr int SNN_BIG_THREADS = 1024;     // (148 x 256 threads left the dense path latency-bound: k = 100 at 1.3M cells took 14 s)
static constexpr int SORT_CTA_MAX = 4096;
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
            // longer lists (hubs: in-degrees reach 1e5 at k = 100): sort runs of SORT_CTA_MAX entries in shared memory,
            // then merge runs pairwise by rank - the entries of one reverse list are distinct cells, so an element's
            // place in the merged run is its index in its own run plus the number of smaller elements in the partner
            // run (binary search); rlist and tmp alternate as source and destination.  (The first version ranked every
            // element against the whole list, O(len^2): 15 s for the k = 100 sweep of C4.)
            for (int64_t c0 = 0; c0 < len; c0 += SORT_CTA_MAX) {
                const int cl = (int)min((int64_t)SORT_CTA_MAX, len - c0);
                int n2 = 64;
                while (n2 < cl) n2 <<= 1;
                for (int i = threadIdx.x; i < n2; i += 256) sm[i] = i < cl ? rlist[a + c0 + i] : INT_MAX;
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
                for (int i = threadIdx.x; i < cl; i += 256) rlist[a + c0 + i] = sm[i];
                __syncthreads();
            }
            int32_t *src = rlist + a, *dst = tmp + a;
            for (int64_t width = SORT_CTA_MAX; width < len; width <<= 1) {
                for (int64_t i = threadIdx.x; i < len; i += 256) {
                    const int64_t run = i / width, mine0 = run * width;
                    const int64_t other0 = (run ^ 1) * width;
                    const int v = src[i];
                    int64_t pos = i;
                    if (other0 < len) {
                        const int64_t other1 = min(len, other0 + width);
                        int64_t lo = other0, hi = other1;                      // first partner element > v
                        while (lo < hi) {
                            const int64_t mid = (lo + hi) >> 1;
                            if (src[mid] < v) lo = mid + 1; else hi = mid;
                        }
                        pos = min(mine0, other0) + (i - mine0) + (lo - other0);
                    }
                    dst[pos] = v;
                }
                __syncthreads();
                int32_t *t = src; src = dst; dst = t;
            }
            if (src != rlist + a) {
                for (int64_t i = threadIdx.x; i < len; i += 256) rlist[a + i] = src[i];
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------
// per-row walk length and classification
//   class 0: walk <= SNN_WALK_LIMIT          one warp per row, 2048-slot table per warp
//   class 1: walk <= 0.7 * SNN_MID_SLOTS     one CTA (256 threads) per row, 4096-slot table
//   class 2: walk <= 0.7 * 8192              one CTA (512 threads) per row, 8192-slot table (3 CTAs/SM for k<=32)
//   class 3: walk <= 0.7 * 16384 (k<=32)     one CTA (512 threads) per row, 16384-slot table
//   class 4: anything longer                 one CTA per row, dense arrays in global scratch
// High-dimensional kNN graphs have hubs (in-degrees in the thousands), so classes 1-2 are the
// common case at scale, not the exception.
// ---------------------------------------------------------------------------------------
static constexpr int SNN_MID_SLOTS = 4096;
static constexpr int SNN_MID_THREADS = 256;
static constexpr int SNN_LARGE_THREADS = 512;
static inline int snn_large_slots(int k) { return k <= 32 ? 16384 : 8192; }
static inline int snn_walk_cap(int slots) { return slots / 10 * 7; }

__global__ void snn_walk_kernel(const int32_t *__restrict__ knn, int k, int64_t row_begin, int64_t n_rows,
                                const int64_t *__restrict__ rptr, int32_t *__restrict__ cls_rows /*[5][n_rows]*/,
                                int *__restrict__ n_cls /*[5]*/, int lim0, int lim1, int lim2, int lim3,
                                unsigned long long *__restrict__ max_walk, unsigned long long *__restrict__ sum_walk,
                                int *__restrict__ n_dup, uint8_t *__restrict__ row_cls, int cls4_mark) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    long long walk = 0;
    if (r < n_rows) {
        const int32_t *row = knn + (row_begin + r) * k;
        for (int c = 0; c < k; ++c) {
            int j = row[c];
            walk += rptr[j + 1] - rptr[j];
        }
        if (row[0] != row_begin + r) atomicAdd(n_dup, 1);
        const int cls = walk <= lim0 ? 0 : (walk <= lim1 ? 1 : (walk <= lim2 ? 2 : (walk <= lim3 ? 3 : 4)));
        cls_rows[(int64_t)cls * n_rows + atomicAdd(n_cls + cls, 1)] = (int32_t)r;
        row_cls[r] = (uint8_t)(cls == 4 ? cls4_mark : cls);        // 5: the row is filled by the dense kernel
    }
    // block-level max / sum before touching the global atomics
    unsigned long long wmax = (unsigned long long)walk, wsum = (unsigned long long)walk;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        wmax = max(wmax, __shfl_xor_sync(AB_FULL, wmax, o));
        wsum += __shfl_xor_sync(AB_FULL, wsum, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(max_walk, wmax);
        atomicAdd(sum_walk, wsum);
    }
}

// ---------------------------------------------------------------------------------------
// shared row state: sorted neighbours, their reverse-list extents and walk positions
// ---------------------------------------------------------------------------------------
template <int KM>
struct WarpRowStateT {
    int nb[KM];                  // sorted neighbour ids
    int base[KM + 1];            // walk position where each neighbour's reverse list starts
    long long start[KM];         // rptr of each neighbour
};
using WarpRowState = WarpRowStateT<SNN_MAXK>;

__device__ __forceinline__ unsigned hash_slot(int key) {
    return ((unsigned)key * 2654435761u) >> (32 - SNN_SLOTS_LG);
}
__device__ __forceinline__ unsigned hash_slot_lg(int key, int lg) { return ((unsigned)key * 2654435761u) >> (32 - lg); }

// Loads and sorts row `r`, fills st; returns the walk length.  One warp.
__device__ __forceinline__ int warp_load_row(const int32_t *__restrict__ knn, int k, int64_t grow,
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
    int run = 0;
    for (int c0 = 0; c0 < k; c0 += 32) {
        const int c = c0 + lane;
        int d = 0;
        if (c < k) {
            const int j = st.nb[c];
            const long long s = rptr[j];
            st.start[c] = s;
            d = (int)(rptr[j + 1] - s);
        }
        int inc = d;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(AB_FULL, inc, o);
            if (lane >= o) inc += t;
        }
        __syncwarp();
        if (c < k) st.base[c] = run + inc - d;
        run += __shfl_sync(AB_FULL, inc, 31);
    }
    if (lane == 0) st.base[k] = run;
    __syncwarp();
    return run;
}

// The same row state from values already in registers (one row ahead, see snn_cta_kernel): lane l holds neighbour
// c = l + 32 i of the UNSORTED row, with its reverse-list start and length.  Shared-memory work only.
template <int KM>
__device__ __forceinline__ int warp_build_row(const int (&nb)[KM / 32], const long long (&rs)[KM / 32], const int (&rd)[KM / 32],
                                              int k, WarpRowStateT<KM> &st, int lane) {
    constexpr int SNN_KR = KM / 32;
#pragma unroll
    for (int i = 0; i < SNN_KR; ++i)
        if (lane + 32 * i < k) st.base[lane + 32 * i] = nb[i];                // base[] as scratch
    __syncwarp();
    int rk[SNN_KR];
#pragma unroll
    for (int i = 0; i < SNN_KR; ++i) {
        const int c = lane + 32 * i;
        rk[i] = 0;
        if (c < k) {
            const int v = nb[i];
            int rank = 0;
            for (int t = 0; t < k; ++t) {
                const int u = st.base[t];
                rank += (u < v) || (u == v && t < c);
            }
            rk[i] = rank;
        }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < SNN_KR; ++i)
        if (lane + 32 * i < k) { st.nb[rk[i]] = nb[i]; st.start[rk[i]] = rs[i]; st.base[rk[i]] = rd[i]; }
    __syncwarp();
    int run = 0;
    for (int c0 = 0; c0 < k; c0 += 32) {
        const int c = c0 + lane;
        const int d = c < k ? st.base[c] : 0;
        int inc = d;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(AB_FULL, inc, o);
            if (lane >= o) inc += t;
        }
        __syncwarp();
        if (c < k) st.base[c] = run + inc - d;
        run += __shfl_sync(AB_FULL, inc, 31);
    }
    if (lane == 0) st.base[k] = run;
    __syncwarp();
    return run;
}

// walk position -> cell id m
__device__ __forceinline__ int walk_fetch(const WarpRowState &st, int k, int p, const int32_t *__restrict__ rlist) {
    int lo = 0, hi = k;                                   // last a with base[a] <= p
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (st.base[mid] <= p) lo = mid; else hi = mid;
    }
    return __ldg(rlist + st.start[lo] + (p - st.base[lo]));
}

// ---------------------------------------------------------------------------------------
// class 0: one warp per row, shared-memory hash  (value = first position << 8 | count)
// ---------------------------------------------------------------------------------------
template <bool FILL>
__global__ void __launch_bounds__(SNN_WARPS * 32)
snn_small_kernel(const int32_t *__restrict__ knn, int k, int64_t row_begin, const int32_t *__restrict__ rows,
                 const int *__restrict__ n_rows_ptr, const int64_t *__restrict__ rptr,
                 const int32_t *__restrict__ rlist, int v_min, int64_t *__restrict__ rowcount,
                 const int64_t *__restrict__ rowptr_out, int32_t *__restrict__ src, int32_t *__restrict__ dst,
                 uint8_t *__restrict__ overlap, int cap, unsigned long long *__restrict__ nz_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned long long nz = 0;                       // nonzeros of A.A^T met by this warp (first encounters)
    WarpRowState &st = reinterpret_cast<WarpRowState *>(smem_raw)[warp];
    int *keys = reinterpret_cast<int *>(smem_raw + sizeof(WarpRowState) * SNN_WARPS) + warp * SNN_SLOTS * 2;
    int *vals = keys + SNN_SLOTS;
    const int64_t w = (int64_t)blockIdx.x * SNN_WARPS + warp;
    const int64_t nw = (int64_t)gridDim.x * SNN_WARPS;
    const int64_t n_rows = *n_rows_ptr;
    for (int64_t q = w; q < n_rows; q += nw) {
        const int64_t r = rows[q];
        const int walk = warp_load_row(knn, k, row_begin + r, rptr, st, lane);
        for (int s = lane; s < SNN_SLOTS; s += 32) keys[s] = -1;
        __syncwarp();
        // walk 1: ascending positions; merge duplicates inside a step with match.any
        for (int p0 = 0; p0 < walk; p0 += 32) {
            const int p = p0 + lane;
            const bool live = p < walk;
            const int m = live ? walk_fetch(st, k, p, rlist) : -1 - lane;
            const unsigned grp = __match_any_sync(AB_FULL, m);
            const bool leader = live && (__ffs(grp) - 1 == lane);
            if (leader) {
                const int add = __popc(grp);
                unsigned s = hash_slot(m);
                while (true) {
                    const int prev = atomicCAS(keys + s, -1, m);
                    if (prev == -1) { vals[s] = (p << 8) | add; break; }
                    if (prev == m) { vals[s] += add; break; }
                    s = (s + 1) & (SNN_SLOTS - 1);
                }
            }
            __syncwarp();
        }
        if (!FILL) {
            int c = 0;
            for (int s = lane; s < SNN_SLOTS; s += 32) c += (keys[s] != -1) && ((vals[s] & 255) >= v_min);
            c = ab_warp_sum_i32(c);
            if (lane == 0) rowcount[r] = c;
        } else {
            // cap > 0: "staging" mode of the single-pass protocol: the row's edges go to its own slot of `cap`
            // entries (dst/overlap are the staging arrays, src is implied) and the row count is written as well
            const int64_t obase = cap ? r * (int64_t)cap : rowptr_out[r];
            int64_t out = obase;
            const int grow = (int)(row_begin + r);
            for (int p0 = ((walk - 1) / 32) * 32; p0 >= 0; p0 -= 32) {
                const int p = p0 + lane;
                bool keep = false;
                int m = 0, cnt = 0;
                if (p < walk) {
                    m = walk_fetch(st, k, p, rlist);
                    unsigned s = hash_slot(m);
                    while (keys[s] != m) s = (s + 1) & (SNN_SLOTS - 1);
                    const int v = vals[s];
                    cnt = v & 255;
                    keep = ((v >> 8) == p) && cnt >= v_min;
                    nz += (v >> 8) == p;
                }
                const unsigned b = __ballot_sync(AB_FULL, keep);
                if (keep) {
                    const int64_t o = out + __popc(b & ~((2u << lane) - 1));     // higher lanes first
                    if (!cap || o - obase < cap) {
                        dst[o] = m;
                        if (!cap) src[o] = grow;
                        if (overlap) overlap[o] = (uint8_t)cnt;
                    }
                }
                out += __popc(b);
            }
            if (cap && lane == 0) rowcount[r] = out - obase;
        }
        __syncwarp();
    }
    if (FILL && nz_out) {
        nz = (unsigned long long)ab_warp_sum_i64((long long)nz);
        if (lane == 0 && nz) atomicAdd(nz_out, nz);
    }
}

// ---------------------------------------------------------------------------------------
// classes 1-2: one CTA per row.  The walk is cut into chunks of <= 32 consecutive entries of ONE
// reverse list (coalesced 128-byte reads, no binary search, no duplicate cell inside a chunk) that
// the warps take round-robin.  Table value = bitmask over the row's sorted neighbours a that
// contain the cell: overlap = popcount, first encounter = lowest set bit (the walk visits the
// lists in ascending a and each list in ascending cell id, so the first encounter of m is at
// (lowest a, rank of m in R(nb[a]))); atomicOr makes the insertion order irrelevant.
// ---------------------------------------------------------------------------------------
template <int KM>
struct CtaRowStateT {            // KM = 32 W in the table kernels (k <= 32 takes 0.7 KB instead of 2.6: six CTAs per SM, not five)
    WarpRowStateT<KM> w;
    int chunk0[KM + 1];          // first chunk of each neighbour's list
    int walk, chunks;
};
using CtaRowState = CtaRowStateT<SNN_MAXK>;

template <int SLOTS, int W, int THREADS, bool FILL>
__global__ void __launch_bounds__(THREADS, W == 1 ? (THREADS == 256 ? 6 : THREADS == 512 ? 3 : 1) : (THREADS == 256 ? 2 : 1))
snn_cta_kernel(const int32_t *__restrict__ knn, int k, int64_t row_begin, const int32_t *__restrict__ rows,
               const int *__restrict__ n_rows_ptr, const int64_t *__restrict__ rptr,
               const int32_t *__restrict__ rlist, int v_min, int64_t *__restrict__ rowcount,
               const int64_t *__restrict__ rowptr_out, int32_t *__restrict__ src, int32_t *__restrict__ dst,
               uint8_t *__restrict__ overlap, int cap, unsigned long long *__restrict__ nz_out) {
    static_assert(SLOTS / 10 * 7 / 32 + 32 * W + 1 <= THREADS, "chunk table must fit one scan pass");
    unsigned long long nz = 0;                       // nonzeros of A.A^T met by this thread (first encounters)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int *keys = reinterpret_cast<int *>(smem_raw);
    unsigned *masks = reinterpret_cast<unsigned *>(keys + SLOTS);
    constexpr int SNN_KR = W;                                          // k <= 32 W
    using RowState = CtaRowStateT<32 * W>;
    RowState &st = *reinterpret_cast<RowState *>(masks + (size_t)SLOTS * W);
    int *chunk_tab = reinterpret_cast<int *>(&st + 1);                 // [THREADS]
    unsigned *ballots = reinterpret_cast<unsigned *>(chunk_tab + THREADS);   // [THREADS]
    int *choff = reinterpret_cast<int *>(ballots + THREADS);           // [THREADS]
    __shared__ int red[32];
    __shared__ int sv_n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NW = THREADS / 32;
    const int n_rows = *n_rows_ptr;
    // Warp 0 keeps the NEXT row's neighbour ids and reverse-list extents in registers: the two dependent global round
    // trips behind them (row -> ids -> extents) are issued while the current row is processed, so building the row
    // state at the top of a row is shared-memory work only (it was ~10 % of the kernel with seven warps waiting).
    int nx_nb[SNN_KR];
    long long nx_s[SNN_KR];
    int nx_d[SNN_KR];
    auto fetch_ids = [&](int qq) {
#pragma unroll
        for (int i = 0; i < SNN_KR; ++i) nx_nb[i] = 0;
        if (qq < n_rows) {
            const int64_t r2 = rows[qq];
#pragma unroll
            for (int i = 0; i < SNN_KR; ++i)
                if (lane + 32 * i < k) nx_nb[i] = knn[(row_begin + r2) * k + lane + 32 * i];
        }
    };
    auto fetch_extents = [&]() {
#pragma unroll
        for (int i = 0; i < SNN_KR; ++i) {
            nx_s[i] = 0; nx_d[i] = 0;
            if (lane + 32 * i < k) {
                const long long s0 = rptr[nx_nb[i]];
                nx_s[i] = s0;
                nx_d[i] = (int)(rptr[nx_nb[i] + 1] - s0);
            }
        }
    };
    if (warp == 0) { fetch_ids(blockIdx.x); fetch_extents(); }
    for (int q = blockIdx.x; q < n_rows; q += gridDim.x) {
        const int64_t r = rows[q];
        __syncthreads();                                               // previous row fully retired
        if (threadIdx.x == 0) sv_n = 0;
        if (warp == 0) {
            const int walk = warp_build_row<32 * W>(nx_nb, nx_s, nx_d, k, st.w, lane);
            fetch_ids(q + (int)gridDim.x);                             // next row: ids now, extents after the insert phase
            int run = 0;
            for (int c0 = 0; c0 < k; c0 += 32) {
                const int c = c0 + lane;
                const int nch = c < k ? (st.w.base[c + 1] - st.w.base[c] + 31) >> 5 : 0;
                int inc = nch;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    int t = __shfl_up_sync(AB_FULL, inc, o);
                    if (lane >= o) inc += t;
                }
                if (c < k) st.chunk0[c] = run + inc - nch;
                run += __shfl_sync(AB_FULL, inc, 31);
            }
            if (lane == 0) { st.chunk0[k] = run; st.walk = walk; st.chunks = run; }
        }
        __syncthreads();
        const int walk = st.walk, C = st.chunks;
        int lg = 6;
        while ((1 << lg) < 2 * walk && (1 << lg) < SLOTS) ++lg;
        const int S = 1 << lg;
        for (int s = threadIdx.x; s < S; s += THREADS) keys[s] = -1;
        for (int s = threadIdx.x; s < S * W; s += THREADS) masks[s] = 0u;
        for (int a = warp; a < k; a += NW) {
            const int c0 = st.chunk0[a], nch = st.chunk0[a + 1] - c0;
            for (int c = lane; c < nch; c += 32) chunk_tab[c0 + c] = (a << 24) | c;
        }
        __syncthreads();
        // ---- walk: insert (four chunks of reverse-list reads in flight per warp: the loop was a chain of exposed
        // global latencies, one per chunk) ----
        constexpr int PF = 4;
        for (int c0 = warp; c0 < C; c0 += NW * PF) {
            int mm[PF], aa[PF];
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                const int c = c0 + u * NW;
                mm[u] = -1; aa[u] = 0;
                if (c < C) {
                    const int e = chunk_tab[c];
                    const int a = e >> 24, t = ((e & 0xFFFFFF) << 5) + lane;
                    aa[u] = a;
                    if (t < st.w.base[a + 1] - st.w.base[a]) mm[u] = __ldg(rlist + st.w.start[a] + t);
                }
            }
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                const int m = mm[u], a = aa[u];
                if (m >= 0) {
                    unsigned s = hash_slot_lg(m, lg);
                    while (true) {
                        const int prev = atomicCAS(keys + s, -1, m);
                        if (prev == -1 || prev == m) break;
                        s = (s + 1) & (S - 1);
                    }
                    atomicOr(masks + (size_t)s * W + (a >> 5), 1u << (a & 31));
                }
            }
        }
        if (warp == 0) fetch_extents();                                 // next row's ids have arrived by now
        __syncthreads();
        if (!FILL) {
            int c = 0;
            for (int s = threadIdx.x; s < S; s += THREADS) {
                if (keys[s] != -1) {
                    int pc = 0;
#pragma unroll
                    for (int w2 = 0; w2 < W; ++w2) pc += __popc(masks[(size_t)s * W + w2]);
                    c += pc >= v_min;
                }
            }
            c = ab_warp_sum_i32(c);
            if (lane == 0) red[warp] = c;
            __syncthreads();
            if (threadIdx.x == 0) {
                int tot = 0;
                for (int w2 = 0; w2 < NW; ++w2) tot += red[w2];
                rowcount[r] = tot;
            }
        } else if (cap) {
            // ---- staging mode (the common path): the kept edges are read off the TABLE, not off a second and third walk.
            // A slot knows its overlap (popcount) and the neighbour a of its first encounter (lowest set bit); the
            // position inside R(nb[a]) - the lists are sorted - is a binary search, done for the ~1 % of the
            // candidates that survive the pruning only.  Emission order = descending walk position (rank by counting).
            int *sv_m = reinterpret_cast<int *>(ballots);                      // [cap] cell ids   (aliases the cap == 0 scratch)
            unsigned *sv_key = reinterpret_cast<unsigned *>(sv_m + cap);       // [cap] walk position << 8 | overlap
            for (int s = threadIdx.x; s < S; s += THREADS) {
                const int m = keys[s];
                if (m != -1) {
                    ++nz;
                    int pc = 0, first = -1;
#pragma unroll
                    for (int w2 = 0; w2 < W; ++w2) {
                        const unsigned mk = masks[(size_t)s * W + w2];
                        if (first < 0 && mk) first = 32 * w2 + __ffs(mk) - 1;
                        pc += __popc(mk);
                    }
                    if (pc >= v_min) {
                        const int i = atomicAdd(&sv_n, 1);
                        if (i < cap) { sv_m[i] = m; sv_key[i] = ((unsigned)first << 8) | (unsigned)pc; }
                    }
                }
            }
            __syncthreads();
            const int total = sv_n;
            if (threadIdx.x == 0) rowcount[r] = total;                         // > cap: the overflow pass redoes the row
            if (total <= cap) {
                for (int i = threadIdx.x; i < total; i += THREADS) {
                    const int m = sv_m[i];
                    const unsigned kx = sv_key[i];
                    const int a = (int)(kx >> 8);
                    const int32_t *__restrict__ L = rlist + st.w.start[a];
                    int lo = 0, hi = st.w.base[a + 1] - st.w.base[a];
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (__ldg(L + mid) < m) lo = mid + 1; else hi = mid;
                    }
                    sv_key[i] = ((unsigned)(st.w.base[a] + lo) << 8) | (kx & 255u);
                }
                __syncthreads();
                const int64_t out = r * (int64_t)cap;
                for (int i = threadIdx.x; i < total; i += THREADS) {
                    const unsigned kx = sv_key[i];
                    int rel = 0;
                    for (int j = 0; j < total; ++j) rel += sv_key[j] > kx;      // walk positions are distinct
                    dst[out + rel] = sv_m[i];
                    if (overlap) overlap[out + rel] = (uint8_t)(kx & 255u);
                }
            }
        } else {
            // ---- pass A: which entries of each chunk are kept first encounters ----
            for (int c = warp; c < C; c += NW) {
                const int e = chunk_tab[c];
                const int a = e >> 24, t = ((e & 0xFFFFFF) << 5) + lane;
                const int len = st.w.base[a + 1] - st.w.base[a];
                bool keep = false;
                if (t < len) {
                    const int m = __ldg(rlist + st.w.start[a] + t);
                    unsigned s = hash_slot_lg(m, lg);
                    while (keys[s] != m) s = (s + 1) & (S - 1);
                    int pc = 0, first = -1;
#pragma unroll
                    for (int w2 = 0; w2 < W; ++w2) {
                        const unsigned mk = masks[(size_t)s * W + w2];
                        if (first < 0 && mk) first = 32 * w2 + __ffs(mk) - 1;
                        pc += __popc(mk);
                    }
                    keep = (first == a) && pc >= v_min;
                    nz += first == a;
                }
                const unsigned b = __ballot_sync(AB_FULL, keep);
                if (lane == 0) ballots[c] = b;
            }
            __syncthreads();
            // ---- suffix sums over chunks: output runs from the LAST walk position to the first ----
            {
                const int cidx = C - 1 - (int)threadIdx.x;                 // thread i owns chunk C-1-i
                const int v = cidx >= 0 ? __popc(ballots[cidx]) : 0;
                int inc = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    int t = __shfl_up_sync(AB_FULL, inc, o);
                    if (lane >= o) inc += t;
                }
                if (lane == 31) red[warp] = inc;
                __syncthreads();
                if (warp == 0) {
                    int x = lane < NW ? red[lane] : 0;
                    int xi = x;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        int t = __shfl_up_sync(AB_FULL, xi, o);
                        if (lane >= o) xi += t;
                    }
                    red[lane] = xi - x;
                }
                __syncthreads();
                if (cidx >= 0) choff[cidx] = red[warp] + inc - v;      // entries emitted before this chunk
                if (cap && cidx == 0) rowcount[r] = red[warp] + inc;   // staging mode: the row's edge count
            }
            __syncthreads();
            // ---- pass B: emit (cap > 0: into the row's staging slot of `cap` entries, see snn_small_kernel) ----
            const int64_t out = cap ? r * (int64_t)cap : rowptr_out[r];
            const int grow = (int)(row_begin + r);
            for (int c = warp; c < C; c += NW) {
                const unsigned b = ballots[c];
                if ((b >> lane) & 1u) {
                    const int e = chunk_tab[c];
                    const int a = e >> 24, t = ((e & 0xFFFFFF) << 5) + lane;
                    const int m = __ldg(rlist + st.w.start[a] + t);
                    unsigned s = hash_slot_lg(m, lg);
                    while (keys[s] != m) s = (s + 1) & (S - 1);
                    int pc = 0;
#pragma unroll
                    for (int w2 = 0; w2 < W; ++w2) pc += __popc(masks[(size_t)s * W + w2]);
                    const int rel = choff[c] + __popc(b & ~((2u << lane) - 1));            // higher lanes first
                    if (!cap || rel < cap) {
                        const int64_t o = out + rel;
                        dst[o] = m;
                        if (!cap) src[o] = grow;
                        if (overlap) overlap[o] = (uint8_t)pc;
                    }
                }
            }
        }
    }
    if (FILL && nz_out) {
        nz = (unsigned long long)ab_warp_sum_i64((long long)nz);
        if (lane == 0 && nz) atomicAdd(nz_out, nz);
    }
}

template <int SLOTS, int W, int THREADS>
static size_t snn_cta_smem(int cap) {
    // [keys][masks][row state][chunk table][scratch: ballots + chunk offsets (cap == 0) or the survivor list (2 * cap ints)]
    return (size_t)SLOTS * 4 * (1 + W) + sizeof(CtaRowStateT<32 * W>) + (size_t)THREADS * 4 +
           (size_t)std::max(2 * THREADS, 2 * cap) * 4 + 64;
}

// ---------------------------------------------------------------------------------------
// class 4: rows whose walk does not fit the largest shared-memory table (hubs; every row at k = 50..100: the walk
// grows with k^2).  Same bitmask table as above, but the candidates are split by a second hash into P partitions
// and the walk is repeated once per partition, so any walk length runs out of shared memory (round 1 used dense
// per-CTA arrays of n_total counters in global memory: 3 GB of scratch, every atomic a DRAM round trip - 15 s for
// the k = 100 sweep of C4).  Kept first encounters (a few hundred per row at most) are collected with their walk
// position in a shared-memory list, sorted by position (descending = scipy's order) and emitted.
// cap > 0: staging mode (row's slot, rowcount written); cap == 0: final positions from rowptr_out.
// Rows with more than PT_LIST kept edges or a partition that overflows its table are handed to the dense kernel
// (queued in `fallback_rows`), which stays as the any-input safety net.
// ---------------------------------------------------------------------------------------
static constexpr int PT_THREADS = 1024;          // (512 until the walk was cut to one pass per partition with batched reads)
static constexpr int PT_LIST = 4096;             // kept edges per row the list can hold

template <int SLOTS, int W>
__global__ void __launch_bounds__(PT_THREADS)
snn_part_kernel(const int32_t *__restrict__ knn, int k, int64_t row_begin, const int32_t *__restrict__ rows,
                const int *__restrict__ n_rows_ptr, const int64_t *__restrict__ rptr,
                const int32_t *__restrict__ rlist, int v_min, int64_t *__restrict__ rowcount,
                const int64_t *__restrict__ rowptr_out, int32_t *__restrict__ src, int32_t *__restrict__ dst,
                uint8_t *__restrict__ overlap, int cap, unsigned long long *__restrict__ nz_out,
                int32_t *__restrict__ fallback_rows, int *__restrict__ n_fallback, uint8_t *__restrict__ row_cls) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int *keys = reinterpret_cast<int *>(smem_raw);
    unsigned *masks = reinterpret_cast<unsigned *>(keys + SLOTS);
    unsigned long long *list = reinterpret_cast<unsigned long long *>(masks + (size_t)SLOTS * W);   // (position << 32) | m
    uint8_t *list_ov = reinterpret_cast<uint8_t *>(list + PT_LIST);
    CtaRowState &st = *reinterpret_cast<CtaRowState *>(list_ov + PT_LIST);
    __shared__ int nkept_sm, nins_sm, bad_sm;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NW = PT_THREADS / 32;
    constexpr int LG = SLOTS == 16384 ? 14 : 13;
    static_assert(SLOTS == (1 << LG), "table size");
    unsigned long long nz = 0;
    const int n_rows = *n_rows_ptr;
    for (int s2 = threadIdx.x; s2 < SLOTS; s2 += PT_THREADS) keys[s2] = -1;
    for (int s2 = threadIdx.x; s2 < SLOTS * W; s2 += PT_THREADS) masks[s2] = 0u;
    for (int q = blockIdx.x; q < n_rows; q += gridDim.x) {
        const int64_t r = rows[q];
        __syncthreads();
        if (warp == 0) {
            const int walk = warp_load_row(knn, k, row_begin + r, rptr, st.w, lane);
            int run = 0;
            for (int c0 = 0; c0 < k; c0 += 32) {
                const int c = c0 + lane;
                const int nch = c < k ? (st.w.base[c + 1] - st.w.base[c] + 31) >> 5 : 0;
                int inc = nch;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    int t = __shfl_up_sync(AB_FULL, inc, o);
                    if (lane >= o) inc += t;
                }
                if (c < k) st.chunk0[c] = run + inc - nch;
                run += __shfl_sync(AB_FULL, inc, 31);
            }
            if (lane == 0) { st.chunk0[k] = run; st.walk = walk; st.chunks = run; nkept_sm = 0; bad_sm = 0; }
        }
        __syncthreads();
        const int walk = st.walk, C = st.chunks;
        unsigned long long nz_row = 0;
        // partitions sized for a load factor <= 0.7 even if every walk entry were a distinct cell (they are not)
        constexpr int PCAP = SLOTS / 10 * 7;
        const unsigned P = (unsigned)((walk + PCAP - 1) / PCAP);
        for (unsigned part = 0; part < P; ++part) {
            // (the table is clean here: cleared once at kernel start, then by the scan below as it reads the slots)
            if (threadIdx.x == 0) nins_sm = 0;
            __syncthreads();
            // ---- insert the walk entries of this partition: PF chunks of reverse-list reads in flight per warp ----
            constexpr int PF = 4;
            for (int c0 = warp; c0 < C; c0 += NW * PF) {
                int mm[PF], aa[PF];
#pragma unroll
                for (int u = 0; u < PF; ++u) {
                    const int c = c0 + u * NW;
                    mm[u] = -1; aa[u] = 0;
                    if (c < C) {
                        int lo = 0, hi = k;                            // list a with chunk0[a] <= c
                        while (hi - lo > 1) {
                            const int mid = (lo + hi) >> 1;
                            if (st.chunk0[mid] <= c) lo = mid; else hi = mid;
                        }
                        const int t = ((c - st.chunk0[lo]) << 5) + lane;
                        aa[u] = lo;
                        if (t < st.w.base[lo + 1] - st.w.base[lo]) mm[u] = __ldg(rlist + st.w.start[lo] + t);
                    }
                }
#pragma unroll
                for (int u = 0; u < PF; ++u) {
                    const int m = mm[u], a = aa[u];
                    if (m < 0) continue;
                    const unsigned h = (unsigned)m * 2654435761u;
                    // partition from the LOW bits, slot from the high bits (a partition taken from the high bits would
                    // squeeze its keys into 1/P of the slots); multiply-shift, no integer division per entry
                    if ((((h & 0x3FFFFu) * P) >> 18) != part) continue;
                    unsigned s2 = h >> (32 - LG);
                    int probes = 0;
                    while (true) {
                        const int prev = atomicCAS(keys + s2, -1, m);
                        if (prev == -1) { atomicAdd(&nins_sm, 1); break; }
                        if (prev == m) break;
                        s2 = (s2 + 1) & (SLOTS - 1);
                        if (++probes >= SLOTS) { bad_sm = 1; break; }      // table full: hand the row over
                    }
                    if (probes < SLOTS) atomicOr(masks + (size_t)s2 * W + (a >> 5), 1u << (a & 31));
                }
            }
            __syncthreads();
            const bool bad = bad_sm || nins_sm > SLOTS - SLOTS / 8;
            // ---- read the table off (and clear it): every occupied slot is a first encounter; the kept ones are listed
            // with the neighbour index of their first encounter, the position inside that list is resolved below ----
            for (int s2 = threadIdx.x; s2 < SLOTS; s2 += PT_THREADS) {
                const int m = keys[s2];
                if (m == -1) continue;
                int pc = 0, first = -1;
#pragma unroll
                for (int w2 = 0; w2 < W; ++w2) {
                    const unsigned mk = masks[(size_t)s2 * W + w2];
                    if (first < 0 && mk) first = 32 * w2 + __ffs(mk) - 1;
                    pc += __popc(mk);
                    masks[(size_t)s2 * W + w2] = 0u;
                }
                keys[s2] = -1;
                if (bad) continue;
                ++nz_row;
                if (pc >= v_min) {
                    const int slot = atomicAdd(&nkept_sm, 1);
                    if (slot < PT_LIST) {
                        list[slot] = ((unsigned long long)(unsigned)first << 32) | (unsigned)m;
                        list_ov[slot] = (uint8_t)min(pc, 255);
                    }
                }
            }
            __syncthreads();
            if (bad) { if (threadIdx.x == 0) bad_sm = 1; __syncthreads(); break; }
        }
        if (!bad_sm && nkept_sm <= PT_LIST) {
            // walk position of every kept cell: binary search in the (sorted) reverse list of its first neighbour;
            // key = (walk position + 1, cell): positions are distinct, padding keys are 0
            const int nk0 = nkept_sm;
            for (int i = threadIdx.x; i < nk0; i += PT_THREADS) {
                const unsigned long long e = list[i];
                const int a = (int)(e >> 32), m = (int)(unsigned)(e & 0xFFFFFFFFull);
                const int32_t *__restrict__ L = rlist + st.w.start[a];
                int lo = 0, hi = st.w.base[a + 1] - st.w.base[a];
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (__ldg(L + mid) < m) lo = mid + 1; else hi = mid;
                }
                list[i] = ((unsigned long long)(unsigned)(st.w.base[a] + lo + 1) << 32) | (unsigned)m;
            }
            __syncthreads();
        }
        const int nkept = nkept_sm;
        if (bad_sm || nkept > PT_LIST) {
            // safety net: the dense kernel recounts and emits this row (count mode runs right after this kernel)
            if (threadIdx.x == 0 && cap) {
                fallback_rows[atomicAdd(n_fallback, 1)] = (int32_t)r;
                rowcount[r] = 0;
                row_cls[r] = 5;                                        // filled by snn_big_kernel<true>, not from staging
            }
            continue;
        }
        nz += nz_row;
        // ---- order: descending walk position.  Bitonic sort of (position, m) keys, overlap carried along ----
        int n2 = 32;
        while (n2 < nkept) n2 <<= 1;
        for (int i = nkept + threadIdx.x; i < n2; i += PT_THREADS) { list[i] = 0ull; list_ov[i] = 0; }
        __syncthreads();
        for (int size = 2; size <= n2; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int i = threadIdx.x; i < (n2 >> 1); i += PT_THREADS) {
                    const int lo = ((i & ~(stride - 1)) << 1) | (i & (stride - 1));
                    const int hi = lo | stride;
                    const bool down = ((lo & size) == 0);              // this block sorts descending
                    const unsigned long long x = list[lo], y = list[hi];
                    if ((x < y) == down) {
                        list[lo] = y; list[hi] = x;
                        const uint8_t ox = list_ov[lo]; list_ov[lo] = list_ov[hi]; list_ov[hi] = ox;
                    }
                }
                __syncthreads();
            }
        }
        const int64_t obase = cap ? r * (int64_t)cap : rowptr_out[r];
        const int grow = (int)(row_begin + r);
        for (int i = threadIdx.x; i < nkept; i += PT_THREADS) {
            if (!cap || i < cap) {
                dst[obase + i] = (int32_t)(unsigned)(list[i] & 0xFFFFFFFFull);
                if (!cap) src[obase + i] = grow;
                if (overlap) overlap[obase + i] = list_ov[i];
            }
        }
        if (cap && threadIdx.x == 0) rowcount[r] = nkept;
    }
    if (nz_out && cap) {
        nz = (unsigned long long)ab_warp_sum_i64((long long)nz);
        if (lane == 0 && nz) atomicAdd(nz_out, nz);
    }
}

template <int SLOTS, int W>
static size_t snn_part_smem() {
    return (size_t)SLOTS * 4 * (1 + W) + (size_t)PT_LIST * 9 + sizeof(CtaRowState) + 64;
}

// ---------------------------------------------------------------------------------------
// class 3: one CTA per row, dense per-CTA arrays in global scratch (cnt, first); any in-degree
// ---------------------------------------------------------------------------------------
template <bool FILL>
__global__ void __launch_bounds__(SNN_BIG_THREADS)
snn_big_kernel(const int32_t *__restrict__ knn, int k, int64_t row_begin, int64_t n_total,
               const int64_t *__restrict__ rptr, const int32_t *__restrict__ rlist, int v_min,
               const int32_t *__restrict__ big_rows, const int *__restrict__ n_big,
               int *__restrict__ dense /* SNN_BIG_CTAS x 2 x n_total, cnt=0 / first=INT_MAX */,
               int64_t *__restrict__ rowcount, const int64_t *__restrict__ rowptr_out,
               int32_t *__restrict__ src, int32_t *__restrict__ dst, uint8_t *__restrict__ overlap,
               unsigned long long *__restrict__ nz_out) {
    __shared__ WarpRowState st;
    unsigned long long nz = 0;
    __shared__ int warp_tot[SNN_BIG_THREADS / 32];
    __shared__ int walk_sm;
    __shared__ int total_sm;
    int *cnt = dense + (int64_t)blockIdx.x * 2 * n_total;
    int *first = cnt + n_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nb = *n_big;
    if (!FILL && (int)blockIdx.x < nb) {
        // a CTA that has rows initialises its own dense arrays (count mode runs first and restores them after
        // every row, so fill mode finds them clean); nothing is touched when no row needs the dense path
        for (int64_t i = threadIdx.x; i < n_total; i += SNN_BIG_THREADS) { cnt[i] = 0; first[i] = INT_MAX; }
    }
    for (int q = blockIdx.x; q < nb; q += gridDim.x) {
        const int64_t r = big_rows[q];
        __syncthreads();
        if (warp == 0) {
            int wl = warp_load_row(knn, k, row_begin + r, rptr, st, lane);
            if (lane == 0) { walk_sm = wl; total_sm = 0; }
        }
        __syncthreads();
        const int walk = walk_sm;
        for (int p = threadIdx.x; p < walk; p += SNN_BIG_THREADS) {
            const int m = walk_fetch(st, k, p, rlist);
            atomicAdd(cnt + m, 1);
            atomicMin(first + m, p);
        }
        __syncthreads();
        if (!FILL) {
            int c = 0;
            for (int p = threadIdx.x; p < walk; p += SNN_BIG_THREADS) {
                const int m = walk_fetch(st, k, p, rlist);
                c += (first[m] == p) && (cnt[m] >= v_min);
                nz += first[m] == p;
            }
            c = ab_warp_sum_i32(c);
            if (lane == 0) atomicAdd(&total_sm, c);
            __syncthreads();
            if (threadIdx.x == 0) rowcount[r] = total_sm;
        } else {
            int64_t out = rowptr_out[r];
            const int grow = (int)(row_begin + r);
            for (int p0 = ((walk - 1) / SNN_BIG_THREADS) * SNN_BIG_THREADS; p0 >= 0; p0 -= SNN_BIG_THREADS) {
                const int p = p0 + threadIdx.x;
                bool keep = false;
                int m = 0, c = 0;
                if (p < walk) {
                    m = walk_fetch(st, k, p, rlist);
                    c = cnt[m];
                    keep = (first[m] == p) && c >= v_min;
                }
                const unsigned b = __ballot_sync(AB_FULL, keep);
                if (lane == 0) warp_tot[warp] = __popc(b);
                __syncthreads();
                int higher = 0, tot = 0;
                for (int w2 = 0; w2 < SNN_BIG_THREADS / 32; ++w2) {
                    tot += warp_tot[w2];
                    if (w2 > warp) higher += warp_tot[w2];
                }
                if (keep) {
                    const int64_t pos = out + higher + __popc(b & ~((2u << lane) - 1));
                    dst[pos] = m;
                    src[pos] = grow;
                    if (overlap) overlap[pos] = (uint8_t)min(c, 255);
                }
                out += tot;
                __syncthreads();
            }
        }
        __syncthreads();
        // restore the dense arrays
        for (int p = threadIdx.x; p < walk; p += SNN_BIG_THREADS) {
            const int m = walk_fetch(st, k, p, rlist);
            cnt[m] = 0;
            first[m] = INT_MAX;
        }
    }
    if (!FILL && nz_out) {
        nz = (unsigned long long)ab_warp_sum_i64((long long)nz);
        if (lane == 0 && nz) atomicAdd(nz_out, nz);
    }
}

// staging -> final edge list: one warp per row; rows whose edges did not fit their staging slot are queued
__global__ void __launch_bounds__(256)
snn_compact_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ stage_dst,
                   const uint8_t *__restrict__ stage_ov, int cap, int64_t n_rows, int64_t row_begin,
                   const uint8_t *__restrict__ row_cls, int32_t *__restrict__ src, int32_t *__restrict__ dst,
                   uint8_t *__restrict__ overlap, int32_t *__restrict__ ovf_rows, int32_t *__restrict__ ovf_part,
                   int *__restrict__ n_ovf /* [0] table rows, [2] partitioned rows */) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * 8;
    for (int64_t r = w; r < n_rows; r += nw) {
        if (row_cls[r] == 5) continue;                       // dense-path rows are filled by snn_big_kernel<true>
        const int64_t a = rowptr[r];
        const int64_t cnt = rowptr[r + 1] - a;
        if (cnt > cap) {                                     // second pass: table kernel, or the partitioned kernel for class 4
            if (lane == 0) {
                if (row_cls[r] == 4) ovf_part[atomicAdd(n_ovf + 2, 1)] = (int32_t)r;
                else ovf_rows[atomicAdd(n_ovf, 1)] = (int32_t)r;
            }
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
    int32_t *cls_rows;   // 5*n_rows
    uint8_t *row_cls;    // n_rows: class of every row
    int32_t *stage_dst;  // n_rows * snn_stage_cap(k): edges of each row, staged by the (single) walk pass
    uint8_t *stage_ov;   // n_rows * snn_stage_cap(k)
    int32_t *ovf_rows;   // n_rows: rows whose edges did not fit their staging slot
    int32_t *ovf_part;   // n_rows: same, rows of the partitioned class
    int32_t *fb_rows;    // n_rows: rows the partitioned kernel handed to the dense kernel
    int *dense;          // SNN_BIG_CTAS*2*n_total
    int *counters;       // 16 ints: [0] n_long, [2] n_dup, [3] bad, [4..8] rows per class, [10] overflow rows (tables),
                         // [11] rows handed to the dense kernel, [12] overflow rows (partitioned class)
    unsigned long long *walk_stats;   // [0] max walk, [1] sum of walks, [2] nonzeros of A.A^T (first encounters)
    int64_t *rowcount;   // n_rows+1
    void *scan_ws;
    size_t scan_bytes;
};

static size_t snn_layout(int64_t n_total, int32_t k, int64_t n_rows, void *ws, size_t ws_bytes, SnnWs *out) {
    ab_arena a(ws ? ws : (void *)256, ws ? ws_bytes : (size_t)-1 / 2);
    SnnWs w;
    w.deg = a.take<int64_t>(n_total + 1);
    w.cursor = a.take<int>(n_total);
    w.rlist = a.take<int32_t>((size_t)n_total * k);
    w.tmp = a.take<int32_t>((size_t)n_total * k);
    w.long_ids = a.take<int32_t>(n_total);
    w.cls_rows = a.take<int32_t>((size_t)5 * (n_rows > 0 ? n_rows : 1));
    w.row_cls = a.take<uint8_t>((size_t)(n_rows > 0 ? n_rows : 1));
    w.stage_dst = a.take<int32_t>((size_t)(n_rows > 0 ? n_rows : 1) * snn_stage_cap(k));
    w.stage_ov = a.take<uint8_t>((size_t)(n_rows > 0 ? n_rows : 1) * snn_stage_cap(k));
    w.ovf_rows = a.take<int32_t>((size_t)(n_rows > 0 ? n_rows : 1));
    w.ovf_part = a.take<int32_t>((size_t)(n_rows > 0 ? n_rows : 1));
    w.fb_rows = a.take<int32_t>((size_t)(n_rows > 0 ? n_rows : 1));
    w.dense = a.take<int>((size_t)SNN_BIG_CTAS * 2 * n_total);
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

static size_t snn_small_smem() { return sizeof(WarpRowState) * SNN_WARPS + (size_t)SNN_WARPS * SNN_SLOTS * 2 * sizeof(int); }

// launches the shared-memory-table kernels of classes 0-3 in FILL mode: cap > 0 stages every row's edges into its own
// slot (dst/overlap = staging arrays, rowcount written), cap == 0 writes final positions given by rowptr
static int snn_launch_tables(ab_ctx *ctx, const SnnWs &w, const int32_t *knn_all, int k, int64_t row_begin, int64_t n_rows,
                             int v_min, int64_t *rowcount, const int64_t *rowptr, int32_t *src, int32_t *dst,
                             uint8_t *overlap, int cap, cudaStream_t st) {
    unsigned long long *nz = cap ? w.walk_stats + 2 : nullptr;       // counted once, in the staging pass
    const int32_t *rows0 = w.cls_rows, *rows1 = w.cls_rows + n_rows, *rows2 = w.cls_rows + 2 * n_rows,
                  *rows3 = w.cls_rows + 3 * n_rows;
    const int *ncls = w.counters + 4;
    {
        const size_t smem = snn_small_smem();
        AB_CUDA(cudaFuncSetAttribute(snn_small_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int64_t blocks = std::min<int64_t>((n_rows + SNN_WARPS - 1) / SNN_WARPS, (int64_t)ctx->sm_count * 4);
        snn_small_kernel<true><<<(unsigned)blocks, SNN_WARPS * 32, smem, st>>>(
            knn_all, k, row_begin, rows0, ncls + 0, w.deg, w.rlist, v_min, rowcount, rowptr, src, dst, overlap, cap, nz);
        AB_LAUNCH_CHECK(ctx);
    }
    const int64_t gcap = std::max<int64_t>(1, n_rows);
#define AB_SNN_CTA(SLOTS, W, THREADS, ROWS, NPTR, PER_SM)                                                        \
    do {                                                                                                       \
        const size_t smem = snn_cta_smem<SLOTS, W, THREADS>(cap);                                              \
        AB_CUDA(cudaFuncSetAttribute(snn_cta_kernel<SLOTS, W, THREADS, true>,                                  \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                 \
        /* rows are dealt round-robin over the CTAs: the grid must be ONE resident wave (a grid of 6 per SM with 5   \
           resident left 148 CTAs for a second, nearly empty wave as long as the first) */                           \
        int resident = 0;                                                                                      \
        AB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, snn_cta_kernel<SLOTS, W, THREADS, true>, \
                                                              THREADS, smem));                                 \
        resident = std::max(1, std::min(resident, (PER_SM)));                                                  \
        const unsigned grid = (unsigned)std::min<int64_t>(gcap, (int64_t)ctx->sm_count * resident);            \
        snn_cta_kernel<SLOTS, W, THREADS, true><<<grid, THREADS, smem, st>>>(                                  \
            knn_all, k, row_begin, ROWS, NPTR, w.deg, w.rlist, v_min, rowcount, rowptr, src, dst, overlap, cap, nz); \
        AB_LAUNCH_CHECK(ctx);                                                                                  \
    } while (0)
    if (k <= 32) {
        AB_SNN_CTA(SNN_MID_SLOTS, 1, SNN_MID_THREADS, rows1, ncls + 1, 6);
        AB_SNN_CTA(8192, 1, SNN_LARGE_THREADS, rows2, ncls + 2, 3);
        AB_SNN_CTA(16384, 1, 1024, rows3, ncls + 3, 1);             // one CTA per SM: all the warps an SM takes
    } else {
        AB_SNN_CTA(SNN_MID_SLOTS, 4, SNN_MID_THREADS, rows1, ncls + 1, 2);
        AB_SNN_CTA(8192, 4, SNN_LARGE_THREADS, rows2, ncls + 2, 1);     // (class 3 stays empty: lim3 == lim2)
    }
    return 0;
}

// second pass for the rows whose edges did not fit their staging slot: the largest table kernel takes any class-0..3 row
static int snn_launch_overflow(ab_ctx *ctx, const SnnWs &w, const int32_t *knn_all, int k, int64_t row_begin, int64_t n_rows,
                               int v_min, const int64_t *rowptr, int32_t *src, int32_t *dst, uint8_t *overlap,
                               cudaStream_t st) {
    const int64_t gcap = std::max<int64_t>(1, n_rows);
    const int32_t *rows = w.ovf_rows;
    const int *nptr = w.counters + 10;
    int64_t *rowcount = nullptr;
    const int cap = 0;
    unsigned long long *nz = nullptr;
    if (k <= 32) AB_SNN_CTA(16384, 1, 1024, rows, nptr, 1);
    else AB_SNN_CTA(8192, 4, SNN_LARGE_THREADS, rows, nptr, 1);
#undef AB_SNN_CTA
    return 0;
}

// class 4 rows (`rows`, `n_listed` on the device) through the partitioned-table kernel; cap as in snn_launch_tables
static int snn_launch_part(ab_ctx *ctx, const SnnWs &w, const int32_t *knn_all, int k, int64_t row_begin, int64_t n_rows,
                           int v_min, const int32_t *rows, const int *n_listed, int64_t *rowcount, const int64_t *rowptr,
                           int32_t *src, int32_t *dst, uint8_t *overlap, int cap, cudaStream_t st) {
    unsigned long long *nz = cap ? w.walk_stats + 2 : nullptr;
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(n_rows, ctx->sm_count));
    if (k <= 32) {
        const size_t smem = snn_part_smem<16384, 1>();
        AB_CUDA(cudaFuncSetAttribute(snn_part_kernel<16384, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        snn_part_kernel<16384, 1><<<grid, PT_THREADS, smem, st>>>(knn_all, k, row_begin, rows, n_listed, w.deg, w.rlist, v_min,
                                                                  rowcount, rowptr, src, dst, overlap, cap, nz, w.fb_rows,
                                                                  w.counters + 11, w.row_cls);
    } else {
        const size_t smem = snn_part_smem<8192, 4>();
        AB_CUDA(cudaFuncSetAttribute(snn_part_kernel<8192, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        snn_part_kernel<8192, 4><<<grid, PT_THREADS, smem, st>>>(knn_all, k, row_begin, rows, n_listed, w.deg, w.rlist, v_min,
                                                                 rowcount, rowptr, src, dst, overlap, cap, nz, w.fb_rows,
                                                                 w.counters + 11, w.row_cls);
    }
    AB_LAUNCH_CHECK(ctx);
    return 0;
}

extern "C" int ab_snn_count(ab_ctx *ctx, const int32_t *knn_all, int64_t n_total, int32_t k, int64_t row_begin,
                            int64_t row_end, int32_t v_min, int64_t *rowptr, int64_t *h_edges, int64_t *stats,
                            void *ws, size_t ws_bytes, void *stream) {
    AB_REQUIRE(ctx, "ab_snn_count: null ctx");
    AB_REQUIRE(k >= 1 && k <= SNN_MAXK, "ab_snn_count: k must be in 1..%d", SNN_MAXK);
    AB_REQUIRE(n_total >= 1 && n_total < INT_MAX, "ab_snn_count: bad n_total");
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
    // rows with a 2-hop walk of at most lim0 entries take the one-warp kernel, longer ones a whole CTA
    const bool force_dense = getenv("AB_SNN_DENSE") != nullptr;        // test knob: class 4 straight to the safety net
    const int lim0 = getenv("AB_SNN_WALK_LIMIT") ? std::min(SNN_WALK_LIMIT, atoi(getenv("AB_SNN_WALK_LIMIT"))) : SNN_WALK_LIMIT;
    if (n_rows > 0) {
        snn_walk_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, st>>>(
            knn_all, k, row_begin, n_rows, w.deg, w.cls_rows, w.counters + 4, lim0, snn_walk_cap(SNN_MID_SLOTS),
            snn_walk_cap(8192), snn_walk_cap(snn_large_slots(k)), w.walk_stats, w.walk_stats + 1, w.counters + 2, w.row_cls,
            force_dense ? 5 : 4);
        AB_LAUNCH_CHECK(ctx);
        // ONE walk per row: the table kernels run in fill mode and stage each row's edges (emission order
        // included) in a per-row slot; ab_snn_fill then only compacts the slots into the final list
        if (snn_launch_tables(ctx, w, knn_all, k, row_begin, n_rows, v_min, w.rowcount, nullptr, nullptr, w.stage_dst,
                              w.stage_ov, snn_stage_cap(k), st))
            return 1;
        // class 4 (walks beyond the largest table): partitioned tables; what even they cannot take -> dense kernel
        if (!force_dense && snn_launch_part(ctx, w, knn_all, k, row_begin, n_rows, v_min, w.cls_rows + 4 * n_rows,
                                            w.counters + 8, w.rowcount, nullptr, nullptr, w.stage_dst, w.stage_ov,
                                            snn_stage_cap(k), st))
            return 1;
        snn_big_kernel<false><<<SNN_BIG_CTAS, SNN_BIG_THREADS, 0, st>>>(
            knn_all, k, row_begin, n_total, w.deg, w.rlist, v_min, force_dense ? w.cls_rows + 4 * n_rows : w.fb_rows,
            force_dense ? w.counters + 8 : w.counters + 11, w.dense,
            w.rowcount, nullptr, nullptr, nullptr, nullptr, w.walk_stats + 2);
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
        // host array: {rows beyond the one-warp path (classes 1-3), longest 2-hop walk, rows whose first
        // neighbour is not the cell itself (duplicate points; alona/clustering.py:204 assumes it is),
        // rows on the dense global path (class 3), sum of all walk lengths (the gather traffic / 4 B)}
        stats[0] = (int64_t)h_cnt[5] + h_cnt[6] + h_cnt[7] + h_cnt[8];
        stats[1] = (int64_t)h_walk[0];
        stats[2] = h_cnt[2];
        stats[3] = force_dense ? h_cnt[8] : h_cnt[11];                     // rows the dense kernel had to take
        stats[4] = (int64_t)h_walk[1];
        stats[5] = (int64_t)h_walk[2];                                     // nonzeros of A.A^T in the row range
        stats[6] = h_cnt[5] + h_cnt[6];                                    // rows of the 4096- / 8192-slot CTA classes
        stats[7] = h_cnt[7];                                               // rows of the 16384-slot class
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
    int rc = 0;
    AB_CUDA(cudaMemsetAsync(w.counters + 10, 0, sizeof(int), st));
    AB_CUDA(cudaMemsetAsync(w.counters + 12, 0, sizeof(int), st));
    snn_compact_kernel<<<(unsigned)std::min<int64_t>((n_rows + 7) / 8, (int64_t)ctx->sm_count * 16), 256, 0, st>>>(
        rowptr, w.stage_dst, w.stage_ov, snn_stage_cap(k), n_rows, row_begin, w.row_cls, src, dst, overlap, w.ovf_rows,
        w.ovf_part, w.counters + 10);
    AB_LAUNCH_CHECK(ctx);
    rc = snn_launch_overflow(ctx, w, knn_all, k, row_begin, n_rows, v_min, rowptr, src, dst, overlap, st);
    if (!rc) rc = snn_launch_part(ctx, w, knn_all, k, row_begin, n_rows, v_min, w.ovf_part, w.counters + 12, nullptr, rowptr,
                                  src, dst, overlap, 0, st);
    if (!rc) {
        const bool force_dense = getenv("AB_SNN_DENSE") != nullptr;
        snn_big_kernel<true><<<SNN_BIG_CTAS, SNN_BIG_THREADS, 0, st>>>(
            knn_all, k, row_begin, n_total, w.deg, w.rlist, v_min, force_dense ? w.cls_rows + 4 * n_rows : w.fb_rows,
            force_dense ? w.counters + 8 : w.counters + 11, w.dense, nullptr, rowptr, src, dst, overlap, nullptr);
        AB_LAUNCH_CHECK(ctx);
    }
    ab_prof_end(ctx, AB_PROF_SNN_FILL, (cudaStream_t)stream);
    return rc;
}
