// K4-K8: truncated SVD of the HVG-restricted matrix by augmented implicitly-restarted Lanczos
// bidiagonalisation, following alona/irlbpy/irlb.py:95-258 step for step (same work dimension,
// same CGS order, same restart rule, same invcheck) so that the trajectory (`steps`, `nmult`)
// and the scores V.diag(s) match the reference's (alona/clustering.py:108-111).
//
// A = A_H^T is H x N (genes x cells); the shard holds A_H as cell-major CSR (n_local x H), so
//   A.v   (irlb.py:165,204)  = contraction over cells  -> "scatter" SpMV with per-warp private
//                              FP64 accumulators in shared memory (no atomics, deterministic),
//                              summed across ranks (all-reduce of H doubles);
//   A^T.w (irlb.py:182)      = per-cell row dot products, local to the shard;
//   CGS panels (irlb.py:190) = tall-skinny GEMV / GEMV-T over the cell dimension, coefficients
//                              summed across ranks (all-reduce of j+1 doubles, then 1 double);
//   SVD of B (irlb.py:224)   = one-sided Jacobi, one CTA, on the device;
//   Ritz updates (:238,:246) = FP64 tall-skinny GEMM.
// Everything is FP64 (HBM-bound; FP64 costs bandwidth only).  All reductions use a fixed
// order, so a run is bitwise reproducible for a given (n_local, world).
#include "ab_common.cuh"
#include "ab_async.cuh"
#include "ab_peer.cuh"
#include <math.h>
#include <stdlib.h>
#include <float.h>
#include <limits.h>
#include <algorithm>

static constexpr int IR_MAXWORK = 128;          // work dimension m_b <= 128
static constexpr int DOT_THREADS = 256;
static constexpr int DOT_ROWS = 4;              // rows per thread
static constexpr int DOT_TILE = DOT_THREADS * DOT_ROWS;

// scalar slots (device doubles)
enum { SC_S = 0, SC_FN2 = 1, SC_SMAX = 2, SC_V0N2 = 3, SC_MUW = 4 /* mu^T w of the current left vector (0 without centring) */,
       SC_SUMX = 5 /* sum over ALL cells of the vector the next A.x multiplies (centring only) */, SC_COUNT = 8 };
// control ints
enum { CT_CONV = 0, CT_KEEP = 1, CT_DONE = 2, CT_ROT = 3, CT_COUNT = 8 };

__device__ __forceinline__ double invcheck(double x) {       // irlb.py:84-92
    return x > 2.0 * DBL_EPSILON ? 1.0 / x : 0.0;
}

// "last block" ticket: returns true in exactly one block, after all others have published.
__device__ __forceinline__ bool last_block_ticket(int *counter, bool sys = false) {
    __shared__ int ticket_sm;
    if (sys) __threadfence_system(); else __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = atomicAdd(counter, 1);
        ticket_sm = (t == (int)gridDim.x - 1);
        if (ticket_sm) *counter = 0;                           // re-arm for the next launch
    }
    __syncthreads();
    return ticket_sm != 0;
}

// Sum of `nparts` per-block partial values in the last block of a reduction kernel: every thread adds a strided
// subset (independent loads), then a fixed butterfly / warp-order combine - a fixed order for a given grid, and a
// dozen dependent steps instead of one thread walking all partials (that serial tail was ~45 us per launch: invisible
// at 1.3 M rows per GPU, a third of the kernel at 163 k).  Every thread of the block must call it.
__device__ __forceinline__ double block_sum_partials(const double *__restrict__ partial, int nparts, double *red /* smem [32] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    double t = 0.0;
    for (int b = threadIdx.x; b < nparts; b += blockDim.x) t += __ldcg(partial + b);
    t = ab_warp_sum_f64(t);
    __syncthreads();
    if (lane == 0) red[warp] = t;
    __syncthreads();
    double tt = 0.0;
    for (int w2 = 0; w2 < nwarp; ++w2) tt += red[w2];
    return tt;
}

// ---------------------------------------------------------------------------------------
// A^T.w - s.V_j   (irlb.py:182,189)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
atw_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col, const double *__restrict__ val,
           int64_t n, const double *__restrict__ w, const double *__restrict__ vj, const double *__restrict__ scal,
           double *__restrict__ F) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * 8;
    const double s = scal[SC_S];
    const double muw = scal[SC_MUW];             // implicit centring: (A - mu 1^T)^T w = A^T w - (mu^T w) 1; 0.0 otherwise
    for (int64_t row = wid; row < n; row += nw) {
        const int64_t a = rowptr[row], b = rowptr[row + 1];
        double acc = 0.0;
        for (int64_t p = a + lane; p < b; p += 32) acc = fma(__ldg(val + p), __ldg(w + __ldg(col + p)), acc);
        acc = ab_warp_sum_f64(acc);
        if (lane == 0) F[row] = __dsub_rn(__dsub_rn(acc, muw), __dmul_rn(s, vj[row]));
    }
}

// ---------------------------------------------------------------------------------------
// Compact A_H for the row-dot products.  The values of one cell row are x(c) = log2(c/S*1e4+1)
// for the handful of distinct counts c the row holds, so a row is stored as its table of
// distinct values plus one 32-bit word per nonzero: gene (low 16 bits) | table slot (high 16).
// 4 bytes per nonzero instead of 12; a row's table (typically 10-30 doubles) is read once from
// HBM and then lives in L1.  The values are the very doubles of val_h, taken in the same order,
// so A^T.w is bit-identical to the plain kernel.
// Built once per ab_irlb call (warp per row; linear search of the row's table in shared memory,
// new values de-duplicated inside each 32-nonzero step with match.any).
// ---------------------------------------------------------------------------------------
static constexpr int CB_WARPS = 8;
static constexpr int CB_CAP = 256;              // distinct values per row the compact form supports

__global__ void __launch_bounds__(CB_WARPS * 32)
ah_compact_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col, const double *__restrict__ val,
                  int64_t n, uint32_t *__restrict__ ent, double *__restrict__ tabv, uint32_t *__restrict__ taboff,
                  unsigned int *__restrict__ cursor /*[0] next table slot, [1] overflow flag*/) {
    __shared__ unsigned long long tab_sm[CB_WARPS][CB_CAP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long *tab = tab_sm[warp];
    const int64_t wid = (int64_t)blockIdx.x * CB_WARPS + warp;
    const int64_t nw = (int64_t)gridDim.x * CB_WARPS;
    for (int64_t row = wid; row < n; row += nw) {
        const int64_t a = rowptr[row], b = rowptr[row + 1];
        int t = 0;                                                      // table entries so far (warp-uniform)
        for (int64_t base = a; base < b; base += 32) {
            const int64_t p = base + lane;
            const bool live = p < b;
            const unsigned long long bits = live ? (unsigned long long)__double_as_longlong(__ldg(val + p)) : 0ull;
            const int c = live ? __ldg(col + p) : 0;
            int slot = -1;
            const int tt = min(t, CB_CAP);
            for (int i = 0; i < tt; ++i)
                if (tab[i] == bits) slot = i;
            const bool miss = live && slot < 0;
            const unsigned need = __ballot_sync(AB_FULL, miss);
            if (need) {
                int leader = lane;
                if (miss) leader = __ffs(__match_any_sync(need, bits)) - 1;
                const bool lead = miss && leader == lane;
                const unsigned leads = __ballot_sync(AB_FULL, lead);
                const int mine = t + __popc(leads & ((1u << lane) - 1));
                if (lead && mine < CB_CAP) tab[mine] = bits;
                const int got = __shfl_sync(AB_FULL, mine, leader);
                if (miss) slot = got;
                t += __popc(leads);
                __syncwarp();
            }
            if (live) ent[p] = (uint32_t)c | ((uint32_t)slot << 16);
        }
        unsigned int off = 0;
        if (t > CB_CAP) {
            if (lane == 0) cursor[1] = 1u;                              // the caller falls back to the plain kernels
            t = CB_CAP;
        }
        if (lane == 0) off = atomicAdd(cursor, (unsigned int)t);
        off = __shfl_sync(AB_FULL, off, 0);
        if (lane == 0) taboff[row] = off;
        for (int i = lane; i < t; i += 32) tabv[(size_t)off + i] = __longlong_as_double((long long)tab[i]);
        __syncwarp();
    }
}

// The A^T.w kernel on it: w (H doubles) is staged in shared memory once per CTA; the row's table is read through
// L1.  With 4 bytes per nonzero the kernel is no longer HBM-bound but bound by the L1/shared data pipe (random
// 8-byte reads of w: ~8 wavefronts per 32 nonzeros).  Measured dead ends: table slots from registers by shuffle
// (the shuffles ride the same pipe: 139 ms against 105), software-pipelined rows with 8 steps of look-ahead at
// half the occupancy (128 ms).
__global__ void __launch_bounds__(256)
atw_compact_kernel(const int64_t *__restrict__ rowptr, const uint32_t *__restrict__ ent, const double *__restrict__ tabv,
                   const uint32_t *__restrict__ taboff, int64_t n, int H, const double *__restrict__ w,
                   const double *__restrict__ vj, const double *__restrict__ scal, double *__restrict__ F) {
    extern __shared__ double w_sm[];
    for (int i = threadIdx.x; i < H; i += blockDim.x) w_sm[i] = w[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * 8;
    const double s = scal[SC_S];
    const double muw = scal[SC_MUW];             // see atw_kernel
    for (int64_t row = wid; row < n; row += nw) {
        const int64_t a = rowptr[row], b = rowptr[row + 1];
        const double *__restrict__ tb = tabv + taboff[row];
        double acc = 0.0;
        for (int64_t p = a + lane; p < b; p += 32) {
            const uint32_t e = __ldcs(ent + p);
            acc = fma(__ldg(tb + (e >> 16)), w_sm[e & 0xFFFFu], acc);
        }
        acc = ab_warp_sum_f64(acc);
        if (lane == 0) F[row] = __dsub_rn(__dsub_rn(acc, muw), __dmul_rn(s, vj[row]));
    }
}

// ---------------------------------------------------------------------------------------
// c[l] = sum_i V[i,l] F[i]  for l < ncols     (irlb.py:74 dotY, contraction over cells)
// Thread <-> cell row inside a 256-row tile, CTAs stride over the tiles; PD_CB columns are
// accumulated in registers per pass, so every thread has PD_CB independent coalesced loads in
// flight and the (shuffle-heavy) reductions happen once per CTA and column, not once per tile.
// Fixed grid and fixed summation order: bitwise reproducible for a given n.
// ---------------------------------------------------------------------------------------
static constexpr int PD_CB = 16;

__global__ void __launch_bounds__(DOT_THREADS)
panel_dot_kernel(const double *__restrict__ V, int64_t ld, int64_t n, int ncols, const double *__restrict__ F,
                 double *__restrict__ partial /*[grid][IR_MAXWORK]*/, double *__restrict__ out, int *counter,
                 const AbPeer *__restrict__ peer_p, uint32_t seq_coef) {
    __shared__ double wsum[DOT_THREADS / 32][PD_CB];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t stride = (int64_t)gridDim.x * DOT_THREADS;
    for (int l0 = 0; l0 < ncols; l0 += PD_CB) {
        double acc[PD_CB];
#pragma unroll
        for (int c = 0; c < PD_CB; ++c) acc[c] = 0.0;
        const double *v0 = V + (int64_t)l0 * ld;
        if (l0 + PD_CB <= ncols) {
            for (int64_t row = (int64_t)blockIdx.x * DOT_THREADS + threadIdx.x; row < n; row += stride) {
                const double f = F[row];
#pragma unroll
                for (int c = 0; c < PD_CB; ++c) acc[c] = fma(v0[(int64_t)c * ld + row], f, acc[c]);
            }
        } else {
            const int nc = ncols - l0;
            for (int64_t row = (int64_t)blockIdx.x * DOT_THREADS + threadIdx.x; row < n; row += stride) {
                const double f = F[row];
#pragma unroll
                for (int c = 0; c < PD_CB; ++c)
                    if (c < nc) acc[c] = fma(v0[(int64_t)c * ld + row], f, acc[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < PD_CB; ++c) {
            const double t = ab_warp_sum_f64(acc[c]);
            if (lane == 0) wsum[warp][c] = t;
        }
        __syncthreads();
        if (threadIdx.x < PD_CB && l0 + threadIdx.x < ncols) {
            double t = 0.0;
#pragma unroll
            for (int w2 = 0; w2 < DOT_THREADS / 32; ++w2) t += wsum[w2][threadIdx.x];
            partial[(int64_t)blockIdx.x * IR_MAXWORK + l0 + threadIdx.x] = t;
        }
        __syncthreads();
    }
    if (last_block_ticket(counter)) {
        // a warp per column, lanes over the per-block partials (independent loads), fixed butterfly
        for (int l = warp; l < ncols; l += DOT_THREADS / 32) {
            double t = 0.0;
            for (unsigned b = lane; b < gridDim.x; b += 32) t += __ldcg(partial + (int64_t)b * IR_MAXWORK + l);
            t = ab_warp_sum_f64(t);
            if (lane == 0) {
                if (ab_peer_world(peer_p) > 1) ab_peer_push(*peer_p, AB_CH_COEF, seq_coef, l, t);      // all-reduce, first half: push
                else out[l] = t;
            }
        }
        if (ab_peer_world(peer_p) > 1) {
            __threadfence_system();
            __syncthreads();
            if (threadIdx.x == 0) ab_peer_publish(*peer_p, AB_CH_COEF, seq_coef);
        }
    }
}

// F -= V[:, :ncols].c ;  fn2 = sum F^2      (irlb.py:79,191).  Thread <-> row, 8 loads in flight.
__global__ void __launch_bounds__(DOT_THREADS)
panel_update_kernel(const double *__restrict__ V, int64_t ld, int64_t n, int ncols, const double *__restrict__ c,
                    double *__restrict__ F, double *__restrict__ partial, double *__restrict__ out_fn2, int *counter,
                    const AbPeer *__restrict__ peer_p, uint32_t seq_coef, uint32_t seq_fn2) {
    __shared__ double cs[IR_MAXWORK];
    __shared__ double red[DOT_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (ab_peer_world(peer_p) > 1) {
        ab_peer_wait(*peer_p, AB_CH_COEF, seq_coef);                                    // all-reduce, second half: sum
        for (int l = threadIdx.x; l < ncols; l += DOT_THREADS) cs[l] = ab_peer_sum(*peer_p, AB_CH_COEF, seq_coef, l);
    } else {
        for (int l = threadIdx.x; l < ncols; l += DOT_THREADS) cs[l] = c[l];
    }
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * DOT_THREADS;
    double acc = 0.0;
    for (int64_t row = (int64_t)blockIdx.x * DOT_THREADS + threadIdx.x; row < n; row += stride) {
        const double *vr = V + row;
        double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
        int l = 0;
        for (; l + 7 < ncols; l += 8) {
            const double a0 = vr[(int64_t)l * ld], a1 = vr[(int64_t)(l + 1) * ld], a2 = vr[(int64_t)(l + 2) * ld],
                         a3 = vr[(int64_t)(l + 3) * ld], a4 = vr[(int64_t)(l + 4) * ld], a5 = vr[(int64_t)(l + 5) * ld],
                         a6 = vr[(int64_t)(l + 6) * ld], a7 = vr[(int64_t)(l + 7) * ld];
            t0 = fma(a0, cs[l], t0); t1 = fma(a1, cs[l + 1], t1); t2 = fma(a2, cs[l + 2], t2); t3 = fma(a3, cs[l + 3], t3);
            t0 = fma(a4, cs[l + 4], t0); t1 = fma(a5, cs[l + 5], t1); t2 = fma(a6, cs[l + 6], t2); t3 = fma(a7, cs[l + 7], t3);
        }
        for (; l < ncols; ++l) t0 = fma(vr[(int64_t)l * ld], cs[l], t0);
        const double nf = __dsub_rn(F[row], (t0 + t1) + (t2 + t3));
        F[row] = nf;
        acc = fma(nf, nf, acc);
    }
    acc = ab_warp_sum_f64(acc);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tt = 0.0;
#pragma unroll
        for (int w2 = 0; w2 < DOT_THREADS / 32; ++w2) tt += red[w2];
        partial[blockIdx.x] = tt;
    }
    if (last_block_ticket(counter)) {
        const double tt = block_sum_partials(partial, (int)gridDim.x, red);
        if (threadIdx.x == 0) {
            if (ab_peer_world(peer_p) > 1) {
                ab_peer_push(*peer_p, AB_CH_SCAL, seq_fn2, 0, tt);
                ab_peer_publish(*peer_p, AB_CH_SCAL, seq_fn2);
            } else {
                *out_fn2 = tt;
            }
        }
    }
}

// ||F||^2 summed over ranks: the consumer kernels (A.v or the final F scaling) finish the all-reduce themselves;
// block 0 also leaves the total in scal[SC_FN2] for the kernels that follow on the stream.
__device__ __forceinline__ double resolve_fn2(const AbPeer *__restrict__ peer_p, uint32_t seq_fn2, double *__restrict__ scal) {
    if (ab_peer_world(peer_p) == 1) return scal[SC_FN2];
    ab_peer_wait(*peer_p, AB_CH_SCAL, seq_fn2);
    const double t = ab_peer_sum(*peer_p, AB_CH_SCAL, seq_fn2, 0);
    if (blockIdx.x == 0 && threadIdx.x == 0) scal[SC_FN2] = t;
    return t;
}

// sum of squares of a vector (start vector normalisation, irlb.py:153)
__global__ void __launch_bounds__(DOT_THREADS)
sumsq_kernel(const double *__restrict__ x, int64_t n, double *__restrict__ partial, double *__restrict__ out,
             int *counter) {
    __shared__ double red[DOT_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * DOT_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * DOT_THREADS)
        acc = fma(x[i], x[i], acc);
    acc = ab_warp_sum_f64(acc);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tt = 0.0;
        for (int w2 = 0; w2 < DOT_THREADS / 32; ++w2) tt += red[w2];
        partial[blockIdx.x] = tt;
    }
    if (last_block_ticket(counter)) {
        const double tt = block_sum_partials(partial, (int)gridDim.x, red);
        if (threadIdx.x == 0) *out = tt;
    }
}

// plain sum of a vector (centring: 1^T x), same fixed-order reduction as sumsq_kernel
__global__ void __launch_bounds__(DOT_THREADS)
vecsum_kernel(const double *__restrict__ x, int64_t n, double *__restrict__ partial, double *__restrict__ out,
              int *counter) {
    __shared__ double red[DOT_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * DOT_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * DOT_THREADS)
        acc += x[i];
    acc = ab_warp_sum_f64(acc);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tt = 0.0;
        for (int w2 = 0; w2 < DOT_THREADS / 32; ++w2) tt += red[w2];
        partial[blockIdx.x] = tt;
    }
    if (last_block_ticket(counter)) {
        const double tt = block_sum_partials(partial, (int)gridDim.x, red);
        if (threadIdx.x == 0) *out = tt;
    }
}

// mu^T w over the H genes (centring), one CTA
__global__ void __launch_bounds__(256)
muw_kernel(const double *__restrict__ mu, const double *__restrict__ w, int H, double *__restrict__ out) {
    __shared__ double red[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc = 0.0;
    for (int g = threadIdx.x; g < H; g += 256) acc = fma(mu[g], w[g], acc);
    acc = ab_warp_sum_f64(acc);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tt = 0.0;
        for (int w2 = 0; w2 < 8; ++w2) tt += red[w2];
        *out = tt;
    }
}

// dst = src * invcheck(sqrt(*sumsq))
__global__ void scale_by_invnorm_kernel(const double *__restrict__ src, double *__restrict__ dst, int64_t n,
                                        const double *__restrict__ sumsq) {
    const double inv = invcheck(sqrt(*sumsq));
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = __dmul_rn(inv, src[i]);
}

// ---------------------------------------------------------------------------------------
// A.x : scatter SpMV.  x_i = scale * src_i (scale = invcheck(sqrt(*sumsq)) or 1), optionally
// stored to vdst (the new Lanczos vector, irlb.py:195).  Each warp owns a private FP64
// accumulator of H entries in shared memory and a contiguous range of rows; it walks the
// nonzeros of that range as ONE flat stream, AV_U steps of 32 nonzeros prefetched into
// registers at a time (the loads of a whole batch are in flight together: with 16 KB of
// accumulator per warp only ~12 warps fit an SM, so the bandwidth has to come from
// memory-level parallelism inside a warp, not from occupancy).  A step is consumed in row
// segments: the genes of one cell row are distinct, so the lanes of a segment never collide
// and plain load/FMA/store is race-free and order-deterministic; __syncwarp separates rows.
// Row metadata (row end, x) travels in registers, 32 rows per refill, read by shuffle.
// ---------------------------------------------------------------------------------------
// steps per register batch, one CTA of 12 warps per SM (measured at C3: 4 -> 210 ms, 6 -> 187, 8 -> 179, 12 -> 196, 16 -> 193:
// fewer bytes in flight below 8, instruction-cache misses of the unrolled consume loop above)
static constexpr int AV_U_WIDE = 8;
// Gene-split form (see ah_split_kernel): the genes are cut in two halves, each warp owns rows x ONE half, so its
// accumulator is H/2 doubles and twice as many warps fit an SM (24 x 8 KB at H = 2000).  The register file then allows
// fewer steps in flight per warp; the bytes in flight per SM stay the same, the warps hiding the shared-memory and
// fixed latencies of the consume loop double.
static constexpr int AV_U_SPLIT = 4;

template <class ColT>                // int32_t: the caller's CSR (halves == 1); uint16_t: the gene-split copy (ids < 65536
struct AvOperand {                   // inside a half: 10 bytes per nonzero instead of 12 for a kernel at 0.78 of its HBM floor)
    const int64_t *rowptr[2];
    const ColT *col[2];              // half 1: gene ids rebased by hs
    const double *val[2];
    int halves, hs;                  // hs: first gene of half 1
};

struct AvRowState {           // row metadata of one warp: 32 rows per refill, read by shuffle
    int64_t blk;              // first row of the register block
    int rp_l;                 // lane l: end offset (relative to the warp's first nonzero) of row blk + l
    double x_l;               // lane l: x of row blk + l
    int ri;                   // current row = blk + ri
    int rend;                 // end offset of the current row
    double xrow;              // x of the current row
};

// Refill of the 32-row metadata block: runs once per 32 rows but is reached from every unrolled step, so it is kept
// out of line (values in, values out - no state by reference) to keep the consume loop inside the instruction cache.
struct AvBlk {
    double x;
    int rp;
};
__device__ __noinline__ AvBlk av_fetch_block(const int64_t *__restrict__ rowptr, int64_t blk, int64_t p0, int64_t r1,
                                             const double *__restrict__ src, double scale, double *__restrict__ vdst,
                                             int lane) {
    const int64_t row = blk + lane;
    const bool ok = row < r1;
    AvBlk b;
    b.rp = ok ? (int)(rowptr[row + 1] - p0) : INT_MAX;           // sentinel: never reached
    b.x = ok ? __dmul_rn(scale, src[row]) : 0.0;
    if (ok && vdst) vdst[row] = b.x;
    return b;
}
__device__ __forceinline__ void av_load_block(AvRowState &r, const int64_t *__restrict__ rowptr, int64_t p0, int64_t r1,
                                              const double *__restrict__ src, double scale, double *__restrict__ vdst,
                                              int lane) {
    const AvBlk b = av_fetch_block(rowptr, r.blk, p0, r1, src, scale, vdst, lane);
    r.rp_l = b.rp;
    r.x_l = b.x;
}

// one step of 32 nonzeros starting at stream offset s0
template <bool FULL>
__device__ __forceinline__ void av_step(AvRowState &r, double *acc, int cgu, double vvu, int s0, int len,
                                        const int64_t *__restrict__ rowptr, int64_t p0, int64_t r1,
                                        const double *__restrict__ src, double scale, double *__restrict__ vdst, int lane) {
    const int s1 = FULL ? s0 + 32 : min(s0 + 32, len);
    const int p = s0 + lane;
    if (r.rend >= s1) {
        // the whole step lies inside the current row
        if (FULL || p < s1) acc[cgu] = fma(vvu, r.xrow, acc[cgu]);
    } else {
        int cur = s0;
        while (cur < s1) {
            if (r.rend <= cur) {
                __syncwarp();                            // row switch: order the two rows' updates
                do {                                     // advance to the row that owns `cur` (skips empty rows)
                    if (++r.ri == 32) {
                        r.blk += 32;
                        av_load_block(r, rowptr, p0, r1, src, scale, vdst, lane);
                        r.ri = 0;
                    }
                    r.rend = __shfl_sync(AB_FULL, r.rp_l, r.ri);
                    r.xrow = ab_shfl_f64(r.x_l, r.ri);
                } while (r.rend <= cur);
            }
            const int seg = min(r.rend, s1);
            if (p >= cur && p < seg) acc[cgu] = fma(vvu, r.xrow, acc[cgu]);
            cur = seg;
        }
    }
}

// consumes one batch of AV_U steps of 32 nonzeros.  FULL: every step of the batch is complete (no bounds checks).
// (Issuing two in-row steps interleaved - their 64 genes are distinct - was measured slower: the kernel is
// sensitive to its own code size, 17 % of its stalls are instruction fetch.)
template <bool FULL, int AV_U>
__device__ __forceinline__ void av_consume(AvRowState &r, double *acc, const int (&cg)[AV_U], const double (&vv)[AV_U],
                                           int base, int len, const int64_t *__restrict__ rowptr, int64_t p0, int64_t r1,
                                           const double *__restrict__ src, double scale, double *__restrict__ vdst, int lane) {
#pragma unroll
    for (int u = 0; u < AV_U; ++u) {
        const int s0 = base + 32 * u;
        if (!FULL && s0 >= len) break;
        av_step<FULL>(r, acc, cg[u], vv[u], s0, len, rowptr, p0, r1, src, scale, vdst, lane);
    }
}

template <int AV_U, int THREADS, class ColT>
__global__ void __launch_bounds__(THREADS, 1)
av_kernel(const AvOperand<ColT> op, int64_t n, int H, const double *__restrict__ src,
          double *__restrict__ scal_fn2 /* scal, or NULL: scale 1 */, double *__restrict__ vdst_all,
          int nwarps_cta /* per gene half */, double *__restrict__ partial /*[grid][H]*/,
          const AbPeer *__restrict__ peer_p, uint32_t seq_fn2) {
    extern __shared__ double acc_sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double scale = scal_fn2 ? invcheck(sqrt(resolve_fn2(peer_p, seq_fn2, scal_fn2))) : 1.0;
    const int Hw = op.halves == 2 ? max(op.hs, H - op.hs) : H;         // accumulator width of one warp
    for (int i = threadIdx.x; i < op.halves * nwarps_cta * Hw; i += blockDim.x) acc_sm[i] = 0.0;
    __syncthreads();
    if (warp < op.halves * nwarps_cta) {
        const int half = warp / nwarps_cta, slot = warp - half * nwarps_cta;
        // (selects, not op.x[half]: a dynamically indexed kernel parameter is copied to local memory first)
        const int64_t *__restrict__ rowptr = half ? op.rowptr[1] : op.rowptr[0];
        const ColT *__restrict__ col = half ? op.col[1] : op.col[0];
        const double *__restrict__ val = half ? op.val[1] : op.val[0];
        double *__restrict__ vdst = half == 0 ? vdst_all : nullptr;   // the new Lanczos vector is stored once
        double *acc = acc_sm + (size_t)warp * Hw;
        const int64_t wid = (int64_t)blockIdx.x * nwarps_cta + slot;
        const int64_t nw = (int64_t)gridDim.x * nwarps_cta;
        const int64_t per = (n + nw - 1) / nw;
        const int64_t r0 = min(n, wid * per), r1 = min(n, r0 + per);
        if (r0 < r1) {
            const int64_t p0 = rowptr[r0];
            const int len = (int)(rowptr[r1] - p0);              // a warp's share is far below 2^31 nonzeros
            const ColT *colw = col + p0;
            const double *valw = val + p0;
            AvRowState r;
            r.blk = r0;
            av_load_block(r, rowptr, p0, r1, src, scale, vdst, lane);
            r.ri = 0;
            r.rend = __shfl_sync(AB_FULL, r.rp_l, 0);
            r.xrow = ab_shfl_f64(r.x_l, 0);
            // two register batches: the loads of batch b+1 are in flight while batch b is consumed
            constexpr int B = 32 * AV_U;
            int cg[2][AV_U];
            double vv[2][AV_U];
#pragma unroll
            for (int u = 0; u < AV_U; ++u) {
                const int p = 32 * u + lane;
                const bool ok = p < len;
                cg[0][u] = ok ? (int)__ldg(colw + p) : 0;
                vv[0][u] = ok ? __ldg(valw + p) : 0.0;
            }
            for (int base2 = 0; base2 < len; base2 += 2 * B) {
#pragma unroll
                for (int bb = 0; bb < 2; ++bb) {
                    const int base = base2 + bb * B;
                    const int nb = base + B;                     // batch fetched now, consumed next
                    if (nb + B <= len) {
#pragma unroll
                        for (int u = 0; u < AV_U; ++u) {
                            cg[bb ^ 1][u] = (int)__ldg(colw + nb + 32 * u + lane);
                            vv[bb ^ 1][u] = __ldg(valw + nb + 32 * u + lane);
                        }
                    } else {
#pragma unroll
                        for (int u = 0; u < AV_U; ++u) {
                            const int p = nb + 32 * u + lane;
                            const bool ok = p < len;
                            cg[bb ^ 1][u] = ok ? (int)__ldg(colw + p) : 0;
                            vv[bb ^ 1][u] = ok ? __ldg(valw + p) : 0.0;
                        }
                    }
                    if (base + B <= len) av_consume<true, AV_U>(r, acc, cg[bb], vv[bb], base, len, rowptr, p0, r1, src, scale, vdst, lane);
                    else av_consume<false, AV_U>(r, acc, cg[bb], vv[bb], base, len, rowptr, p0, r1, src, scale, vdst, lane);
                }
            }
            // rows after the last nonzero (empty tail rows) still need their x stored
            if (vdst) {
                for (int64_t b2 = r.blk + 32; b2 < r1; b2 += 32) {
                    const int64_t row = b2 + lane;
                    if (row < r1) vdst[row] = __dmul_rn(scale, src[row]);
                }
            }
        }
    }
    __syncthreads();
    for (int g = threadIdx.x; g < H; g += blockDim.x) {
        const int half = (op.halves == 2 && g >= op.hs) ? 1 : 0;
        const double *a0 = acc_sm + (size_t)half * nwarps_cta * Hw + (g - half * op.hs);
        double t = 0.0;
        for (int w2 = 0; w2 < nwarps_cta; ++w2) t += a0[(size_t)w2 * Hw];
        partial[(size_t)blockIdx.x * H + g] = t;
    }
}

// ---------------------------------------------------------------------------------------
// Gene-split copy of A_H for A.x, built once per ab_irlb call: row r's nonzeros with gene < hs go to stream 0, the
// others (gene ids rebased by hs) to stream 1, each a CSR of its own (rows stay in order, so a warp's flat walk over
// its row range works unchanged).  rp0 = exclusive scan of the per-row counts of half 0.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ah_split_count_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col, int64_t n, int hs,
                      int64_t *__restrict__ cnt0) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    int64_t lo = rowptr[row], hi = rowptr[row + 1];
    const int64_t a = lo;
    while (lo < hi) {                                   // first entry with gene >= hs (columns ascend inside a row)
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(col + mid) < hs) lo = mid + 1; else hi = mid;
    }
    cnt0[row] = lo - a;
}

__global__ void __launch_bounds__(256)
ah_split_fill_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col, const double *__restrict__ val,
                     int64_t n, int hs, const int64_t *__restrict__ rp0 /* n+1, scanned */, int64_t *__restrict__ rp1 /* n+1 */,
                     uint16_t *__restrict__ col2, double *__restrict__ val2) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * 8;
    const int64_t base0 = rowptr[0], nnz0 = rp0[n];
    for (int64_t row = wid; row <= n; row += nw) {
        const int64_t a = rowptr[row];
        // stream 1 starts at nnz0: what precedes row `row` there = all earlier nonzeros minus those of half 0
        const int64_t o1 = nnz0 + (a - base0) - rp0[row];
        if (lane == 0) rp1[row] = o1;
        if (row == n) break;
        const int64_t b = rowptr[row + 1], o0 = rp0[row], c0 = rp0[row + 1] - o0;
        for (int64_t p = a + lane; p < b; p += 32) {
            const int c = __ldg(col + p);
            const int64_t i = p - a;
            const int64_t dst = i < c0 ? o0 + i : o1 + (i - c0);
            col2[dst] = (uint16_t)(i < c0 ? c : c - hs);
            val2[dst] = __ldg(val + p);
        }
    }
}

// Sum of the per-CTA partial results of av_kernel in CTA order.  One GPU: the vector goes to `out`.  Sharded: it
// is this rank's contribution to the all-reduce of A.v and goes straight into every rank's exchange slot over
// NVLink; the last block raises the flags (the consumer, wstep_kernel, sums the slots in rank order).
__global__ void __launch_bounds__(128)
av_reduce_kernel(const double *__restrict__ partial, int nparts, int H, double *__restrict__ out, int *counter,
                 const AbPeer *__restrict__ peer_p, uint32_t seq_vec, const double *__restrict__ mu /* NULL, or on ONE rank the gene means */,
                 const double *__restrict__ scal, int scaled) {
    // four threads per gene (a quarter of the CTA partials each, combined in a fixed order): the chain of 148
    // dependent loads of one thread per gene was 11 us of every A.x at any GPU count
    const int g = blockIdx.x * (blockDim.x / 4) + (threadIdx.x >> 2), part = threadIdx.x & 3;
    double t = 0.0;
    if (g < H)
        for (int b = part; b < nparts; b += 4) t += partial[(size_t)b * H + g];
    t += ab_shfl_xor_f64(t, 1);
    t += ab_shfl_xor_f64(t, 2);
    if (g < H && part == 0) {
        // implicit centring: (A - mu 1^T) x = A x - mu (1^T x); the correction enters once (on rank 0's contribution)
        if (mu) {
            const double sx = (scaled ? invcheck(sqrt(scal[SC_FN2])) : 1.0) * scal[SC_SUMX];
            t = __dsub_rn(t, __dmul_rn(mu[g], sx));
        }
        if (ab_peer_world(peer_p) > 1) ab_peer_push(*peer_p, AB_CH_VEC, seq_vec, g, t);
        else out[g] = t;
    }
    if (ab_peer_world(peer_p) > 1 && last_block_ticket(counter, true)) {
        if (threadIdx.x == 0) ab_peer_publish(*peer_p, AB_CH_VEC, seq_vec);
    }
}

// ---------------------------------------------------------------------------------------
// left-vector step (irlb.py:173-178 and :214-219):
//   t = wraw - (sub_prev ? fn*W[:,jd-1] : 0);  t -= W[:, :ncols].(W[:, :ncols]^T t);
//   s = ||t||;  W[:,jd] = invcheck(s)*t.   Also records B[jb,jb]=s_old, B[jb,jb+1]=fn (irlb.py:196-197).
// One thread-block cluster of WS_CL CTAs (it was one CTA, 66 us per call, replicated on every rank): CTA c owns the
// gene rows [H*c/WS_CL, H*(c+1)/WS_CL), keeps its slice of the panel W in shared memory (read once for both CGS
// passes), and the two global sums (W^T t, ||t||^2) are exchanged through distributed shared memory: every CTA
// stores its partial into all CTAs' inboxes, one cluster barrier, every CTA adds the WS_CL partials in CTA order.
// The partition is fixed (WS_CL = 8 whatever the GPU count), so the result does not depend on the sharding.
// Sharded runs: `wraw` is the sum over ranks of the exchange slots (the all-reduce of A.v finishes here).
// ---------------------------------------------------------------------------------------
static constexpr int WS_CL = 8;
static constexpr int WS_THREADS = 512;

__global__ void __launch_bounds__(WS_THREADS, 1)
wstep_kernel(const double *__restrict__ wraw, double *__restrict__ W, int H, int64_t ldw, int jd, int ncols, int sub_prev,
             double *__restrict__ scal, double *__restrict__ B, int work, int jb, int cache_w,
             const AbPeer *__restrict__ peer_p, uint32_t seq_vec) {
    extern __shared__ double ws_sm[];
    __shared__ double inbox_c[WS_CL][IR_MAXWORK];        // partial W^T t of every CTA of the cluster
    __shared__ double inbox_n[WS_CL];                    // partial ||t||^2
    __shared__ double c2[IR_MAXWORK];
    __shared__ double red[WS_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = WS_THREADS / 32;
    const uint32_t cr = ab_cluster_ctarank();
    const int g0 = (int)(((int64_t)H * cr) / WS_CL), g1 = (int)(((int64_t)H * (cr + 1)) / WS_CL);
    const int ns = g1 - g0;
    const int nsp = (H + WS_CL - 1) / WS_CL + 1;         // slice capacity
    double *t = ws_sm;                                   // [nsp]
    double *Wc = ws_sm + nsp;                            // [ncols][ns] when cache_w
    if (ab_peer_world(peer_p) > 1) ab_peer_wait(*peer_p, AB_CH_VEC, seq_vec);
    const double fn = sqrt(scal[SC_FN2]);
    if (cr == 0 && threadIdx.x == 0 && jb >= 0) {
        B[jb * work + jb] = scal[SC_S];
        if (jb + 1 < work) B[jb * work + jb + 1] = fn;
    }
    for (int g = threadIdx.x; g < ns; g += WS_THREADS) {
        double v = ab_peer_world(peer_p) > 1 ? ab_peer_sum(*peer_p, AB_CH_VEC, seq_vec, g0 + g) : wraw[g0 + g];
        if (sub_prev) v = __dsub_rn(v, __dmul_rn(fn, W[(size_t)(jd - 1) * ldw + g0 + g]));
        t[g] = v;
    }
    if (cache_w)
        for (int l = warp; l < ncols; l += nwarp)
            for (int g = lane; g < ns; g += 32) Wc[(size_t)l * ns + g] = W[(size_t)l * ldw + g0 + g];
    __syncthreads();
    ab_cluster_sync();                                   // every CTA of the cluster is running: its inboxes may be written
    if (ncols > 0) {
        for (int l = warp; l < ncols; l += nwarp) {
            const double *wl = cache_w ? Wc + (size_t)l * ns : W + (size_t)l * ldw + g0;
            double a0 = 0.0, a1 = 0.0;
            int g = lane;
            for (; g + 32 < ns; g += 64) {
                a0 = fma(wl[g], t[g], a0);
                a1 = fma(wl[g + 32], t[g + 32], a1);
            }
            if (g < ns) a0 = fma(wl[g], t[g], a0);
            const double acc = ab_warp_sum_f64(a0 + a1);
            if (lane < WS_CL) ab_st_cluster_f64(ab_mapa(ab_smem_u32(&inbox_c[cr][l]), lane), acc);
        }
        ab_cluster_sync();
        for (int l = threadIdx.x; l < ncols; l += WS_THREADS) {
            double a = 0.0;
#pragma unroll
            for (int q = 0; q < WS_CL; ++q) a += inbox_c[q][l];
            c2[l] = a;
        }
        __syncthreads();
        for (int g = threadIdx.x; g < ns; g += WS_THREADS) {
            const double *wg = cache_w ? Wc + g : W + g0 + g;
            const size_t st = cache_w ? (size_t)ns : (size_t)ldw;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            int l = 0;
            for (; l + 3 < ncols; l += 4) {
                a0 = fma(wg[(size_t)l * st], c2[l], a0);
                a1 = fma(wg[(size_t)(l + 1) * st], c2[l + 1], a1);
                a2 = fma(wg[(size_t)(l + 2) * st], c2[l + 2], a2);
                a3 = fma(wg[(size_t)(l + 3) * st], c2[l + 3], a3);
            }
            for (; l < ncols; ++l) a0 = fma(wg[(size_t)l * st], c2[l], a0);
            t[g] = __dsub_rn(t[g], (a0 + a1) + (a2 + a3));
        }
        __syncthreads();
    }
    double acc = 0.0;
    for (int g = threadIdx.x; g < ns; g += WS_THREADS) acc = fma(t[g], t[g], acc);
    acc = ab_warp_sum_f64(acc);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (warp == 0) {
        double v = lane < nwarp ? red[lane] : 0.0;
        v = ab_warp_sum_f64(v);
        if (lane < WS_CL) ab_st_cluster_f64(ab_mapa(ab_smem_u32(&inbox_n[cr]), lane), v);
    }
    ab_cluster_sync();
    double tot = 0.0;
#pragma unroll
    for (int q = 0; q < WS_CL; ++q) tot += inbox_n[q];
    const double s = sqrt(tot);
    const double sinv = invcheck(s);
    for (int g = threadIdx.x; g < ns; g += WS_THREADS) W[(size_t)jd * ldw + g0 + g] = __dmul_rn(sinv, t[g]);
    if (cr == 0 && threadIdx.x == 0) scal[SC_S] = s;
}

// last Lanczos step of a pass (j = work-1): F *= invcheck(fn); B[j,j] = s  (irlb.py:193,221)
__global__ void fscale_kernel(double *__restrict__ F, int64_t n, double *__restrict__ scal,
                              double *__restrict__ B, int work, int jb, const AbPeer *__restrict__ peer_p, uint32_t seq_fn2) {
    const double inv = invcheck(sqrt(resolve_fn2(peer_p, seq_fn2, scal)));
    if (blockIdx.x == 0 && threadIdx.x == 0) B[jb * work + jb] = scal[SC_S];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        F[i] = __dmul_rn(inv, F[i]);
}

// ---------------------------------------------------------------------------------------
// SVD of the small matrix B (work x work), one-sided Jacobi, one CTA.   (irlb.py:224)
// Outputs: Ub (columns = left vectors), sb (descending), Vb (columns = right vectors),
// all row-major work x work.  Then residuals / convergence / restart bookkeeping (:225-234).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
svd_restart_kernel(const double *__restrict__ B, int work, double *__restrict__ Ub, double *__restrict__ sb,
                   double *__restrict__ Vb, double *__restrict__ Gws /* 2*work*work scratch */,
                   double *__restrict__ scal, int *__restrict__ ctrl, double *__restrict__ R, int nu, double tol,
                   int it, int keep_prev, int q_in_smem) {
    // G (col-major, work x work) starts as B and converges to U.diag(s) under right rotations.  The right
    // vectors are NOT accumulated through the rotations (that doubles the FP64 work of a kernel that is
    // bound by one SM's FP64 pipe): V = B^T U diag(1/s) is formed once at the end.  A copy of B sits next
    // to G in shared memory when it fits (q_in_smem), else it is read from global memory.
    extern __shared__ double svd_sm[];
    const int m = work;
    const int mp = (m + 1) & ~1;                        // round-robin needs an even player count
    const int gs = m | 1;                               // column stride of G: odd, so the four pairs a warp hosts spread over the banks
    double *G = svd_sm;                                 // [m][gs]
    double *Bs = q_in_smem ? svd_sm + (size_t)m * gs : Gws + (size_t)work * work;   // row-major copy of B
    unsigned short *sched = reinterpret_cast<unsigned short *>(svd_sm + (size_t)m * gs + (q_in_smem ? (size_t)m * m : 0));
    __shared__ double nrm[IR_MAXWORK];
    __shared__ double cn[IR_MAXWORK];                   // running squared column norms
    __shared__ int order[IR_MAXWORK];
    __shared__ int rot_sm;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
        const int r = e % m, c = e / m;
        const double v = B[r * m + c];
        G[(size_t)c * gs + r] = v;
        Bs[r * m + c] = v;
    }
    // pair schedule of a sweep (circle method: player mp-1 fixed, the others rotate), p | q << 8, computed once
    for (int e = threadIdx.x; e < (mp - 1) * (mp / 2); e += blockDim.x) {
        const int round = e / (mp / 2), pp = e - round * (mp / 2);
        const int a = (pp == 0) ? mp - 1 : (round + pp) % (mp - 1);
        const int b = (round + mp - 1 - pp) % (mp - 1);
        sched[e] = (unsigned short)(min(a, b) | (max(a, b) << 8));
    }
    __syncthreads();
    // One-sided Jacobi is a chain of ~(m-1) x 7.5 dependent rounds; what a round costs is its latency chain
    //   dot (m/LP FMAs per lane) -> shuffle-reduce -> rotation parameters -> shuffle -> rotate -> barrier.
    // The CTA is exactly LP lanes per column pair (blockDim = LP * mp/2), so a round is one pass with no idle warps
    // issuing; ONE lane per pair runs the parameter chain, which is two dependent FP64 rsqrt instead of
    // sqrt -> divide -> rsqrt:  with r = sqrt(dd^2 + 4 g^2), cos(2 theta) = |dd| / r,  c = sqrt(h), h = (1 + cos 2theta) / 2,
    // s = sign(dd) g / (r c),  t = s / c:   rs = rsqrt(dd^2 + 4 g^2);  rc = rsqrt(h);  c = h rc;  s = sign g rs rc;  t = s rc.
    // Measured: 2400 cycles per round (42.6 ms per C3 pass; the form with every lane running a three-operation chain took
    // 48) against 1900 for svd_cluster_kernel, which spreads the rows over 8 CTAs - the dependent FP64 latencies of the
    // dot / parameter / rotate chain, not the exchange, are what a round costs, and the cluster shortens the dot and the
    // rotation.  Kept as the A/B reference (AB_IRLB_SVD1).
    constexpr int LP = 8;                               // lanes per column pair (4 pairs per warp)
    const int pr = threadIdx.x / LP, hl = threadIdx.x % LP;
    for (int sweep = 0; sweep < 60; ++sweep) {
        if (threadIdx.x == 0) rot_sm = 0;
        // refresh the running norms once per sweep (they are updated incrementally inside it)
        for (int c = warp; c < m; c += nwarp) {
            double acc = 0.0;
            for (int r = lane; r < m; r += 32) acc = fma(G[(size_t)c * gs + r], G[(size_t)c * gs + r], acc);
            acc = ab_warp_sum_f64(acc);
            if (lane == 0) cn[c] = acc;
        }
        __syncthreads();
        for (int round = 0; round < mp - 1; ++round) {
            const unsigned pq = pr < mp / 2 ? sched[round * (mp / 2) + pr] : 0xFFFFu;
            const int p = pq & 255, q = pq >> 8;
            const bool valid = pr < mp / 2 && q < m;
            double *gp = G + (size_t)(valid ? p : 0) * gs, *gq = G + (size_t)(valid ? q : 0) * gs;
            double g0 = 0.0, g1 = 0.0;
            if (valid) {
                int r = hl;
                for (; r + LP < m; r += 2 * LP) { g0 = fma(gp[r], gq[r], g0); g1 = fma(gp[r + LP], gq[r + LP], g1); }
                if (r < m) g0 = fma(gp[r], gq[r], g0);
            }
            double ga = g0 + g1;
#pragma unroll
            for (int o = LP / 2; o > 0; o >>= 1) ga += ab_shfl_xor_f64(ga, o);    // stays inside the lane group
            double cs = 1.0, sn = 0.0;
            if (valid && hl == 0) {
                const double al = cn[p], be = cn[q];
                const double g2 = ga * ga, ab2 = al * be;
                // rotate down to rounding level, but call the sweep "dirty" only above 1e-14: pairs hovering at
                // 1e-15..1e-16 (pure rounding noise) would otherwise keep the loop alive without changing any vector
                if (g2 > 1e-30 * ab2 && fabs(ga) > 1e-300) {
                    if (g2 > 1e-28 * ab2) atomicMax(&rot_sm, (g2 > 1e-14 * ab2) ? 2 : 1);   // 2: a cosine above 1e-7 was seen
                    const double dd = be - al;
                    const double rs = rsqrt(fma(dd, dd, 4.0 * g2));
                    const double h = fma(0.5 * fabs(dd), rs, 0.5);
                    const double rc = rsqrt(h);
                    cs = h * rc;
                    sn = (dd >= 0.0 ? ga : -ga) * rs * rc;
                    const double tg = sn * rc * ga;                 // tan(theta) g
                    cn[p] = al - tg;
                    cn[q] = be + tg;
                }
            }
            cs = ab_shfl_f64(cs, (threadIdx.x & 31) & ~(LP - 1));
            sn = ab_shfl_f64(sn, (threadIdx.x & 31) & ~(LP - 1));
            if (valid && sn != 0.0) {
                for (int r = hl; r < m; r += LP) {
                    const double x = gp[r], y = gq[r];
                    gp[r] = cs * x - sn * y; gq[r] = sn * x + cs * y;
                }
            }
            __syncthreads();
        }
        const int any = rot_sm;
        __syncthreads();
        if (threadIdx.x == 0) ctrl[CT_ROT] = sweep + 1;     // diagnostics: Jacobi sweeps of this call
        if (any < 2) break;                                 // see svd_cluster_kernel
    }
    // singular values = column norms; sort descending by counting rank (ties by index)
    for (int c = warp; c < m; c += nwarp) {
        double acc = 0.0;
        for (int r = lane; r < m; r += 32) acc = fma(G[(size_t)c * gs + r], G[(size_t)c * gs + r], acc);
        acc = ab_warp_sum_f64(acc);
        if (lane == 0) nrm[c] = sqrt(acc);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < m; c += blockDim.x) {
        int rank = 0;
        for (int o = 0; o < m; ++o) rank += (nrm[o] > nrm[c]) || (nrm[o] == nrm[c] && o < c);
        order[rank] = c;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
        const int r = e / m, i = e % m;
        const int c = order[i];
        const double sv = nrm[c];
        Ub[r * m + i] = sv > 0.0 ? G[(size_t)c * gs + r] / sv : 0.0;
        // right vector i = B^T u_i / s_i = B^T G[:, c] / s_i^2
        double acc = 0.0;
        const double *gc = G + (size_t)c * gs;
        for (int l = 0; l < m; ++l) acc = fma(Bs[l * m + r], gc[l], acc);
        Vb[r * m + i] = sv > 0.0 ? acc / (sv * sv) : 0.0;
    }
    for (int i = threadIdx.x; i < m; i += blockDim.x) sb[i] = nrm[order[i]];
    __syncthreads();
    // residuals and restart bookkeeping
    const double fn = sqrt(scal[SC_FN2]);
    __shared__ int conv_sm;
    if (threadIdx.x == 0) {
        double smax = sb[0];
        if (it > 0) smax = fmax(smax, scal[SC_SMAX]);
        scal[SC_SMAX] = smax;
        conv_sm = 0;
    }
    __syncthreads();
    const double smax = scal[SC_SMAX];
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const double res = __dmul_rn(fn, Ub[(m - 1) * m + i]);
        R[i] = res;
        if (i < nu && fabs(res) < tol * smax) atomicAdd(&conv_sm, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int conv = conv_sm;
        ctrl[CT_CONV] = conv;
        if (conv >= nu) {
            ctrl[CT_DONE] = 1;
            ctrl[CT_KEEP] = keep_prev;
        } else {
            int k = max(conv + nu, keep_prev);
            k = min(k, m - 3);
            ctrl[CT_DONE] = 0;
            ctrl[CT_KEEP] = k;
        }
    }
}

// ---------------------------------------------------------------------------------------
// The same SVD on one thread-block cluster of SV_CL CTAs.  One-sided Jacobi is a chain of ~520 dependent rounds
// (dot -> reduce -> divide / sqrt / rsqrt -> rotate -> barrier; ~2700 cycles each on one SM: 0.75 ms per restart,
// replicated on every rank, 19 % of the 8-GPU step).  Here the ROWS of G are split across the cluster: CTA c keeps
// rows [m*c/SV_CL, m*(c+1)/SV_CL) of every column in shared memory.  A round = partial inner products of the pairs
// over the local rows -> every CTA sends its partials into all CTAs' inboxes with st.async (distributed shared
// memory; each store completes 8 bytes of the RECEIVER's mbarrier, so a CTA only waits on its own barrier - no
// cluster-wide barrier in the round) -> every CTA adds the SV_CL partials in CTA order, so all CTAs hold
// bit-identical inner products, derive the same rotation and apply it to their own rows.  Running column norms, the
// "dirty" flag and the pair schedule are replicated and stay identical.  Right vectors as before:
// V = B^T U diag(1/s), formed from G written to global memory.
// ---------------------------------------------------------------------------------------
static constexpr int SV_CL = 8;
static constexpr int SV_LP = 4;                          // lanes per column pair
static constexpr int SV_THREADS = (IR_MAXWORK / 2) * SV_LP;      // 256: one pass covers every pair of a round
static constexpr int SV_RS = IR_MAXWORK / SV_CL + 1;     // row-slice stride (17: odd -> conflict-free column walks)

__global__ void __launch_bounds__(SV_THREADS, 1)
svd_cluster_kernel(const double *__restrict__ B, int work, double *__restrict__ Ub, double *__restrict__ sb,
                   double *__restrict__ Vb, double *__restrict__ Gws /* work*work scratch, G column-major */,
                   double *__restrict__ scal, int *__restrict__ ctrl, double *__restrict__ R, int nu, double tol,
                   int it, int keep_prev) {
    __shared__ double Gs[IR_MAXWORK * SV_RS];            // Gs[c * SV_RS + (r - r0)]
    __shared__ double inbox[2][SV_CL][IR_MAXWORK / 2];   // partial inner products, double-buffered by round parity
    __shared__ double inbox_n[SV_CL][IR_MAXWORK];        // partial squared column norms
    __shared__ double cn[IR_MAXWORK];
    __shared__ double nrm[IR_MAXWORK];
    __shared__ int order[IR_MAXWORK];
    __shared__ int rot_sm, conv_sm;
    __shared__ __align__(8) uint64_t rbar[2];            // "round r's partials have all landed", by round parity
    const int m = work;
    const uint32_t cr = ab_cluster_ctarank();
    if (threadIdx.x == 0) {
        ab_mbar_init(rbar + 0, 1);
        ab_mbar_init(rbar + 1, 1);
        ab_mbar_init_fence();
    }
    const int r0 = (m * (int)cr) / SV_CL, r1 = (m * ((int)cr + 1)) / SV_CL, nr = r1 - r0;
    for (int e = threadIdx.x; e < m * nr; e += SV_THREADS) {
        const int c = e / nr, r = e - c * nr;
        Gs[c * SV_RS + r] = B[(r0 + r) * m + c];
    }
    __syncthreads();
    ab_cluster_sync();                                   // all CTAs of the cluster run: inboxes may be written

    // squared column norms, summed over the cluster (partials -> every CTA's inbox_n -> barrier -> CTA-order sum)
    auto refresh_norms = [&]() {
        for (int c = threadIdx.x; c < m; c += SV_THREADS) {
            double acc = 0.0;
            for (int r = 0; r < nr; ++r) acc = fma(Gs[c * SV_RS + r], Gs[c * SV_RS + r], acc);
            const uint32_t a = ab_smem_u32(&inbox_n[cr][c]);
#pragma unroll
            for (int q = 0; q < SV_CL; ++q) ab_st_cluster_f64(ab_mapa(a, q), acc);
        }
        ab_cluster_sync();
        for (int c = threadIdx.x; c < m; c += SV_THREADS) {
            double t = 0.0;
#pragma unroll
            for (int q = 0; q < SV_CL; ++q) t += inbox_n[q][c];
            cn[c] = t;
        }
        __syncthreads();
        ab_cluster_sync();                               // inbox_n is free again before anyone refills it
    };

    const int mp = (m + 1) & ~1;
    const int pr = threadIdx.x / SV_LP, hl = threadIdx.x % SV_LP;
    // this thread's column pair of every round of a sweep (circle method: player mp-1 fixed, the others rotate), packed
    // p | q << 8, computed once: the two integer modulo sequences sat on the critical path of each of the ~520 rounds
    extern __shared__ unsigned short sched[];            // [(mp - 1) * (mp / 2)], dynamic
    for (int e = threadIdx.x; e < (mp - 1) * (mp / 2); e += SV_THREADS) {
        const int round = e / (mp / 2), pp = e - round * (mp / 2);
        const int a = (pp == 0) ? mp - 1 : (round + pp) % (mp - 1);
        const int b = (round + mp - 1 - pp) % (mp - 1);
        sched[e] = (unsigned short)(min(a, b) | (max(a, b) << 8));
    }
    __syncthreads();
    uint32_t round_id = 0;
    int sweeps = 0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        if (threadIdx.x == 0) rot_sm = 0;
        refresh_norms();
        for (int round = 0; round < mp - 1; ++round, ++round_id) {
            const unsigned pq = pr < mp / 2 ? sched[round * (mp / 2) + pr] : 0xFFFFu;
            const int p = pq & 255, q = pq >> 8;
            const bool valid = pr < mp / 2 && q < m;
            double *gp = Gs + (valid ? p : 0) * SV_RS, *gq = Gs + (valid ? q : 0) * SV_RS;
            double ga = 0.0;
            if (valid)
                for (int r = hl; r < nr; r += SV_LP) ga = fma(gp[r], gq[r], ga);
#pragma unroll
            for (int o = SV_LP / 2; o > 0; o >>= 1) ga += ab_shfl_xor_f64(ga, o);
            const int par = round_id & 1;
            // a barrier is re-armed for round r+2 only after this CTA has consumed round r (two rounds ago), and a
            // peer can only send round r+2 after it received this CTA's round r+1, sent after that consumption
            if (threadIdx.x == 0) ab_mbar_expect_tx(rbar + par, (uint32_t)(SV_CL * (mp / 2) * 8));
            if (pr < mp / 2) {
                const uint32_t ad = ab_smem_u32(&inbox[par][cr][pr]), bd = ab_smem_u32(rbar + par);
#pragma unroll
                for (int qq = hl; qq < SV_CL; qq += SV_LP) ab_st_async_f64(ab_mapa(ad, qq), ga, ab_mapa(bd, qq));
            }
            ab_mbar_wait(rbar + par, (round_id >> 1) & 1);
            // rotation parameters: one lane per pair runs the FP64 divide / sqrt / rsqrt chain (every lane of the
            // pair doing it was most of a round on a part with ~8 FP64 operations per clock and SM), the others
            // receive cos / sin by shuffle
            double cs = 1.0, sn = 0.0;
            if (valid && hl == 0) {
                double gs = 0.0;
#pragma unroll
                for (int qq = 0; qq < SV_CL; ++qq) gs += inbox[par][qq][pr];
                const double al = cn[p], be = cn[q];
                const double g2 = gs * gs, ab2 = al * be;
                // same thresholds as the one-CTA kernel: rotate down to rounding level, "dirty" only above 1e-14
                if (g2 > 1e-30 * ab2 && fabs(gs) > 1e-300) {
                    if (g2 > 1e-28 * ab2) atomicMax(&rot_sm, (g2 > 1e-14 * ab2) ? 2 : 1);   // 2: a cosine above 1e-7 was seen
                    // two dependent rsqrt instead of sqrt -> divide -> rsqrt (see svd_restart_kernel)
                    const double dd = be - al;
                    const double rs = rsqrt(fma(dd, dd, 4.0 * g2));
                    const double h = fma(0.5 * fabs(dd), rs, 0.5);
                    const double rc = rsqrt(h);
                    cs = h * rc;
                    sn = (dd >= 0.0 ? gs : -gs) * rs * rc;
                    const double tg = sn * rc * gs;                 // tan(theta) g
                    cn[p] = al - tg;
                    cn[q] = be + tg;
                }
            }
            cs = ab_shfl_f64(cs, (threadIdx.x & 31) & ~(SV_LP - 1));
            sn = ab_shfl_f64(sn, (threadIdx.x & 31) & ~(SV_LP - 1));
            if (valid && sn != 0.0) {
                for (int r = hl; r < nr; r += SV_LP) {
                    const double x = gp[r], y = gq[r];
                    gp[r] = cs * x - sn * y; gq[r] = sn * x + cs * y;
                }
            }
            __syncthreads();
        }
        const int any = rot_sm;
        __syncthreads();
        sweeps = sweep + 1;
        // Stop rule.  Convergence is quadratic: a sweep whose largest cosine was below 1e-7 leaves cosines below ~1e-14,
        // so it is the last one that changes anything - the sweep that merely confirms it (one in 7.5) is skipped.
        // Checked on the restart matrices of the C1 fixtures (m_b = 70 and 120, CPU restatement): 103 -> 91 and 276 -> 244
        // sweeps, orthogonality and singular values unchanged at 7e-15 / 9e-15.
        if (any < 2) break;
    }
    if (cr == 0 && threadIdx.x == 0) ctrl[CT_ROT] = sweeps;
    // singular values = column norms; sort descending by counting rank (ties by index)
    refresh_norms();
    for (int c = threadIdx.x; c < m; c += SV_THREADS) nrm[c] = sqrt(cn[c]);
    __syncthreads();
    for (int c = threadIdx.x; c < m; c += SV_THREADS) {
        int rank = 0;
        for (int o = 0; o < m; ++o) rank += (nrm[o] > nrm[c]) || (nrm[o] == nrm[c] && o < c);
        order[rank] = c;
    }
    __syncthreads();
    // left vectors (own rows) and G to global memory for the right-vector products
    for (int e = threadIdx.x; e < m * nr; e += SV_THREADS) {
        const int i = e / nr, r = e - i * nr;
        const int c = order[i];
        const double sv = nrm[c];
        const double g = Gs[c * SV_RS + r];
        Ub[(r0 + r) * m + i] = sv > 0.0 ? g / sv : 0.0;
        Gws[(size_t)c * m + r0 + r] = g;
    }
    if (cr == 0)
        for (int i = threadIdx.x; i < m; i += SV_THREADS) sb[i] = nrm[order[i]];
    __threadfence();
    __syncthreads();
    ab_cluster_sync();
    // right vector i = B^T u_i / s_i = B^T G[:, c] / s_i^2 ; this CTA forms rows [r0, r1) of Vb
    for (int e = threadIdx.x; e < m * nr; e += SV_THREADS) {
        const int i = e % m, r = r0 + e / m;
        const int c = order[i];
        const double sv = nrm[c];
        double acc = 0.0;
        const double *gc = Gws + (size_t)c * m;
        for (int l = 0; l < m; ++l) acc = fma(B[l * m + r], __ldcg(gc + l), acc);
        Vb[r * m + i] = sv > 0.0 ? acc / (sv * sv) : 0.0;
    }
    // residuals and restart bookkeeping (irlb.py:225-234): the CTA that owns the last row of U
    if (cr == SV_CL - 1) {
        const double fn = sqrt(scal[SC_FN2]);
        double smax = nrm[order[0]];
        if (it > 0) smax = fmax(smax, scal[SC_SMAX]);
        if (threadIdx.x == 0) conv_sm = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < m; i += SV_THREADS) {
            const int c = order[i];
            const double sv = nrm[c];
            const double ulast = sv > 0.0 ? Gs[c * SV_RS + (m - 1 - r0)] / sv : 0.0;
            const double res = __dmul_rn(fn, ulast);
            R[i] = res;
            if (i < nu && fabs(res) < tol * smax) atomicAdd(&conv_sm, 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            scal[SC_SMAX] = smax;
            const int conv = conv_sm;
            ctrl[CT_CONV] = conv;
            if (conv >= nu) {
                ctrl[CT_DONE] = 1;
                ctrl[CT_KEEP] = keep_prev;
            } else {
                int k = max(conv + nu, keep_prev);
                k = min(k, m - 3);
                ctrl[CT_DONE] = 0;
                ctrl[CT_KEEP] = k;
            }
        }
    }
}

// B = 0; B[l,l] = sb[l] (l<keep); B[l,keep] = R[l]   (irlb.py:240-244)
__global__ void reset_b_kernel(double *__restrict__ B, int work, const double *__restrict__ sb,
                               const double *__restrict__ R, int keep) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < work * work; e += gridDim.x * blockDim.x) {
        const int r = e / work, c = e % work;
        double v = 0.0;
        if (r < keep && c == r) v = sb[r];
        if (r < keep && c == keep) v = R[r];
        B[e] = v;
    }
}

// ---------------------------------------------------------------------------------------
// Ritz update: out[:, c] = X[:, :work] . M[:, c]  for c < ncols     (irlb.py:238,246,249-250)
// X column-major (ld even, >= n), M row-major work x work (columns 0..ncols-1 used).
// out column-major (ldo even, >= n) or, if rowmajor_out, row-major [n x ncols] scaled by colscale[c].
//
// The one GEMM-shaped step of IRLBA (n x work times work x keep, FP64, 1.3 GB / 5e9 FMA per restart at C3).
// Measured on this part (benchmarks/micro_atomics.cu, profiles/r2_ritz_svd_norm_ncu.txt): the FP64 CUDA cores sustain
// 61 FMA/clk/SM (35.6 TFLOP/s), the FP64 tensor instruction mma.sync.m8n8k4 only ~28 (its sub-pipe was 69 % busy at
// 10.9 TFLOP/s in the DMMA version of this kernel), so the products are register-tiled DFMA (thread tile below).
// What bounds a naive version is operand delivery, not
// arithmetic: a producer warp streams the column-major panel through shared memory with 1-D bulk copies (one 1 KB
// column segment per copy, one copy per lane, completion on an mbarrier, a ring of stages of KC panel columns); the
// result leaves straight from the registers (32 contiguous bytes per thread and column).
// Thread tile: rows 2l, 2l+1, 64+2l, 65+2l (l = lane) of the 128-row tile, columns CT*w .. CT*w+CT-1 (w = warp): per
// k-row 2 LDS.128 of A (each 512 contiguous bytes per warp, conflict-free) and CT doubles of B at a warp-UNIFORM
// address (one wavefront per load), 4*CT FMAs.  The first DFMA version split a warp 8 row groups x 4 column groups:
// its B loads then cost a wavefront per quarter warp, 24 shared-memory wavefronts per 28 FMA instructions - bound by
// the shared-memory pipe (ncu: 65 % busy), not by the FP64 cores.  Eight consumer warps = 128 rows x 8*CT columns.
// ---------------------------------------------------------------------------------------
static constexpr int RZ_TR = 128;                // rows per tile
static constexpr int RZ_XS = RZ_TR + 4;          // staged-tile column stride (doubles)
static constexpr int RZ_CONSUMERS = 8;           // consumer warps; warp RZ_CONSUMERS is the producer
static constexpr int RZ_THREADS = (RZ_CONSUMERS + 1) * 32;
static constexpr int RZ_MAXST = 6;

struct RitzCfg { int kc, st, ct; size_t smem; };
static RitzCfg ritz_cfg(int work, int ncols, size_t smem_optin) {
    RitzCfg c;
    c.ct = std::max(1, (ncols + 7) / 8);                          // columns per thread
    const int ctp = (c.ct + 1) & ~1;                              // padded to even: B rows are read with 16-byte loads
    c.kc = 24;
    const size_t fixed = (size_t)work * 8 * ctp * 8 + 2 * RZ_MAXST * 8 + 64;
    const size_t stage = (size_t)c.kc * RZ_XS * 8;
    c.st = (int)std::min<size_t>(RZ_MAXST, (std::min<size_t>(smem_optin, 227 * 1024) - 1024 - fixed) / stage);
    c.smem = fixed + (size_t)std::max(c.st, 1) * stage;
    return c;
}

template <int CT>
__global__ void __launch_bounds__(RZ_THREADS, 1)
ritz_kernel(const double *__restrict__ X, int64_t ld, int64_t n, int work, const double *__restrict__ M, int ncols,
            double *__restrict__ out, int64_t ldo, int rowmajor_out, const double *__restrict__ colscale,
            double *__restrict__ out2 /* optional second row-major output without scaling */, int KC, int ST) {
    extern __shared__ __align__(128) unsigned char rz_raw[];
    constexpr int CTP = (CT + 1) & ~1;                                // thread's B slice, padded to an even length
    constexpr int MS = 8 * CTP;                                       // M row stride in shared memory
    double *Ms = reinterpret_cast<double *>(rz_raw);                  // [work][8 thread-columns][CTP]
    double *Xs = Ms + (size_t)work * MS;                              // [ST][KC][RZ_XS]
    uint64_t *full = reinterpret_cast<uint64_t *>(Xs + (size_t)ST * KC * RZ_XS);
    uint64_t *empty = full + RZ_MAXST;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nchunks = (work + KC - 1) / KC;
    const int64_t ntiles = (n + RZ_TR - 1) / RZ_TR;

    // M in the order the threads consume it: thread-column group t (= 4*wc + tc) holds output columns t*CT .. t*CT+CT-1
    for (int e = threadIdx.x; e < work * MS; e += blockDim.x) {
        const int l = e / MS, r = e - l * MS;
        const int t = r / CTP, j = r - t * CTP;
        const int c = t * CT + j;
        Ms[e] = (j < CT && c < ncols) ? M[l * work + c] : 0.0;
    }
    for (int e = threadIdx.x; e < ST * KC * RZ_XS; e += blockDim.x) Xs[e] = 0.0;      // rows past a clipped tile stay finite
    if (threadIdx.x == 0) {
        for (int s2 = 0; s2 < ST; ++s2) { ab_mbar_init(full + s2, 1); ab_mbar_init(empty + s2, RZ_CONSUMERS); }
        ab_mbar_init_fence();
    }
    ab_fence_proxy_async();
    __syncthreads();

    if (warp == RZ_CONSUMERS) {
        // ---------------- producer warp: lane l issues the bulk copy of k-row l of every (tile, k-chunk) ------------
        uint32_t it = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const double *xt = X + tile * RZ_TR;
            const uint32_t cb = (uint32_t)min((int64_t)RZ_TR, ld - tile * RZ_TR) * 8;     // last tile: clipped to ld
            for (int ch = 0; ch < nchunks; ++ch, ++it) {
                const int st = it % ST;
                if (it >= (uint32_t)ST) ab_mbar_wait(empty + st, ((it / ST) - 1) & 1);
                const int k0 = ch * KC, nk = min(KC, work - k0);
                if (lane == 0) ab_mbar_expect_tx(full + st, (uint32_t)nk * cb);
                __syncwarp();
                if (lane < nk)
                    ab_bulk_g2s(Xs + ((size_t)st * KC + lane) * RZ_XS, xt + (int64_t)(k0 + lane) * ld, cb, full + st);
            }
        }
        return;
    }

    // ---------------- consumers -----------------------------------------------------------------------------
    const int rloc = 2 * lane;                            // this thread's rows inside the tile: rloc, rloc+1, 64+rloc, 65+rloc
    const int tg = warp;                                  // thread-column group = warp
    const int cglob = tg * CT;                            // first of this thread's CT output columns
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        double acc[4][CT];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < CT; ++j) acc[i][j] = 0.0;
        for (int ch = 0; ch < nchunks; ++ch, ++it) {
            const int st = it % ST;
            ab_mbar_wait(full + st, (it / ST) & 1);
            const int k0 = ch * KC, nk = min(KC, work - k0);
            const double *xs = Xs + (size_t)st * KC * RZ_XS + rloc;
            const double *ms = Ms + (size_t)k0 * MS + tg * CTP;
#pragma unroll 4
            for (int k = 0; k < nk; ++k) {
                const double2 a01 = *reinterpret_cast<const double2 *>(xs + (size_t)k * RZ_XS);
                const double2 a23 = *reinterpret_cast<const double2 *>(xs + (size_t)k * RZ_XS + 64);
                double b[CTP];
#pragma unroll
                for (int j = 0; j < CTP; j += 2) {
                    const double2 bb = *reinterpret_cast<const double2 *>(ms + (size_t)k * MS + j);
                    b[j] = bb.x; b[j + 1] = bb.y;
                }
#pragma unroll
                for (int j = 0; j < CT; ++j) {
                    acc[0][j] = fma(a01.x, b[j], acc[0][j]);
                    acc[1][j] = fma(a01.y, b[j], acc[1][j]);
                    acc[2][j] = fma(a23.x, b[j], acc[2][j]);
                    acc[3][j] = fma(a23.y, b[j], acc[3][j]);
                }
            }
            __syncwarp();
            if (lane == 0) ab_mbar_arrive(empty + st);
        }
        const int64_t row0 = tile * RZ_TR + rloc;
        if (rowmajor_out) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int64_t r = row0 + (i & 1) + 64 * (i >> 1);
                if (r >= n) continue;
#pragma unroll
                for (int j = 0; j < CT; ++j) {
                    const int col = cglob + j;
                    if (col < ncols) {
                        out[r * (int64_t)ncols + col] = __dmul_rn(acc[i][j], colscale ? colscale[col] : 1.0);
                        if (out2) out2[r * (int64_t)ncols + col] = acc[i][j];
                    }
                }
            }
        } else {
            // straight from the registers: a thread owns two pairs of consecutive rows (16 bytes each) of each of its
            // columns, a warp 2 x 512 contiguous bytes.  (A shared-memory staged tile + bulk stores cost two block
            // barriers per tile: 18 % of the kernel's stall samples.)  Rows in [n, ldo) are padding.
#pragma unroll
            for (int j = 0; j < CT; ++j) {
                const int col = cglob + j;
                if (col < ncols) {
                    double *o = out + (int64_t)col * ldo + row0;
                    if (row0 + 2 <= ldo) *reinterpret_cast<double2 *>(o) = make_double2(acc[0][j], acc[1][j]);
                    if (row0 + 66 <= ldo) *reinterpret_cast<double2 *>(o + 64) = make_double2(acc[2][j], acc[3][j]);
                }
            }
        }
    }
}

typedef void (*ritz_fn)(const double *, int64_t, int64_t, int, const double *, int, double *, int64_t, int, const double *,
                        double *, int, int);
static ritz_fn ritz_pick(int ct) {
    switch (ct) {
        case 1: return ritz_kernel<1>;   case 2: return ritz_kernel<2>;   case 3: return ritz_kernel<3>;
        case 4: return ritz_kernel<4>;   case 5: return ritz_kernel<5>;   case 6: return ritz_kernel<6>;
        case 7: return ritz_kernel<7>;   case 8: return ritz_kernel<8>;   case 9: return ritz_kernel<9>;
        case 10: return ritz_kernel<10>; case 11: return ritz_kernel<11>; case 12: return ritz_kernel<12>;
        case 13: return ritz_kernel<13>; case 14: return ritz_kernel<14>; case 15: return ritz_kernel<15>;
        default: return ritz_kernel<16>;
    }
}

__global__ void copy_vec_kernel(const double *__restrict__ src, double *__restrict__ dst, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

// ---------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------
struct IrlbWs {
    double *V, *V2, *W, *W2, *F, *B, *Ub, *Vb, *sb, *R, *G, *scal, *coef, *wraw;
    double *dot_partial, *nrm_partial, *av_partial;
    int *ctrl;
    int av_grid, av_warps;
    int64_t dot_grid;
};

static int irlb_work_dim(int nval, int64_t n_total) {
    int64_t w = nval + 20;
    if (3LL * nval < w) w = 3LL * nval;
    if (n_total < w) w = n_total;
    return (int)w;
}

// leading dimension of the column-major panels: even (16-byte aligned columns for the bulk copies of ritz_kernel) and
// with a column stride that is not close to a multiple of 4 KB (panel_dot / panel_update stream 8-16 columns at once:
// a stride of 64 rows x 8 B cost them 11 % against the odd stride the first version happened to have)
static int64_t irlb_ld(int64_t rows) {
    int64_t ld = (std::max<int64_t>(rows, 2) + 1) & ~(int64_t)1;
    const int64_t r = (ld * 8) & 4095;
    if (r < 256 || r > 3840) ld += 64;
    return ld;
}

static int av_warps_for(int H) {
    int w = (int)((216 * 1024) / ((size_t)H * 8));
    if (w > 12) w = 12;
    return w;
}
// gene-split form: warps per gene half (two halves of (H+1)/2 accumulators each)
static int av_split_hs(int H) { return (H + 1) / 2; }
static int av_warps_split(int H) {
    int w = (int)((216 * 1024) / ((size_t)2 * av_split_hs(H) * 8));
    if (w > 12) w = 12;
    return w;
}

static size_t irlb_layout(int64_t n, int H, int nval, int64_t n_total_hint, int sm_count, void *ws, size_t bytes,
                          IrlbWs *out) {
    const int work = irlb_work_dim(nval, n_total_hint);
    ab_arena a(ws ? ws : (void *)256, ws ? bytes : (size_t)-1 / 2);
    IrlbWs w;
    // (ritz_kernel clips its last row tile to ld, so its bulk copies never leave a column)
    const int64_t ld = irlb_ld(n);
    const int64_t ldw = irlb_ld(H);
    w.V = a.take<double>((size_t)ld * work);
    w.V2 = a.take<double>((size_t)ld * work);
    w.W = a.take<double>((size_t)ldw * work);
    w.W2 = a.take<double>((size_t)ldw * work);
    w.F = a.take<double>(ld);
    w.B = a.take<double>((size_t)work * work);
    w.Ub = a.take<double>((size_t)work * work);
    w.Vb = a.take<double>((size_t)work * work);
    w.sb = a.take<double>(work);
    w.R = a.take<double>(work);
    w.G = a.take<double>((size_t)2 * work * work);
    w.scal = a.take<double>(SC_COUNT);
    w.coef = a.take<double>(IR_MAXWORK);
    w.wraw = a.take<double>(H);
    w.dot_grid = std::min<int64_t>((n + DOT_THREADS - 1) / DOT_THREADS, (int64_t)sm_count * 4);
    if (w.dot_grid < 1) w.dot_grid = 1;
    w.dot_partial = a.take<double>((size_t)w.dot_grid * IR_MAXWORK);
    w.nrm_partial = a.take<double>((size_t)(w.dot_grid > 1024 ? w.dot_grid : 1024));
    w.av_warps = av_warps_for(H);
    w.av_grid = sm_count;
    w.av_partial = a.take<double>((size_t)w.av_grid * H);
    w.ctrl = a.take<int>(CT_COUNT);
    if (out) *out = w;
    return a.ok ? a.off : 0;
}

extern "C" size_t ab_irlb_ws_bytes(int64_t n_local, int32_t n_hvg, int32_t nval) {
    // work dimension is at most nval+20; n_total >= that in every supported call
    return irlb_layout(n_local, n_hvg, nval, (int64_t)1 << 40, 256, nullptr, 0, nullptr) + 256;
}

// compact A_H (see ah_compact_kernel): entries, worst-case value tables (every value distinct), table offsets
struct IrlbCompact {
    uint32_t *ent, *taboff;
    double *tabv;
    unsigned int *cursor;
    // gene-split copy of A_H for A.x (ah_split_*_kernel)
    uint16_t *col2;
    double *val2;
    int64_t *rp0, *rp1;
    void *scan;
    size_t scan_bytes;
};
static size_t irlb_compact_layout(int64_t n, int64_t nnz, void *base, size_t bytes, IrlbCompact *out) {
    ab_arena a(base ? base : (void *)256, base ? bytes : (size_t)-1 / 2);
    IrlbCompact c;
    c.tabv = a.take<double>((size_t)std::max<int64_t>(nnz, 1) + 32);      // +32: a lane-wide table read never leaves the buffer
    c.ent = a.take<uint32_t>((size_t)std::max<int64_t>(nnz, 1));
    c.taboff = a.take<uint32_t>((size_t)std::max<int64_t>(n, 1));
    c.cursor = a.take<unsigned int>(4);
    c.col2 = a.take<uint16_t>((size_t)std::max<int64_t>(nnz, 1) + 8);
    c.val2 = a.take<double>((size_t)std::max<int64_t>(nnz, 1));
    c.rp0 = a.take<int64_t>((size_t)n + 1);
    c.rp1 = a.take<int64_t>((size_t)n + 1);
    c.scan_bytes = ab_scan_ws_bytes(n);
    c.scan = a.take<char>(c.scan_bytes);
    if (out) *out = c;
    return a.ok ? a.off : 0;
}
static bool irlb_compact_eligible(int64_t n, int H, int64_t nnz) {
    return n > 0 && H <= 65536 && nnz > 0 && nnz < ((int64_t)1 << 32);
}

extern "C" size_t ab_irlb_ws_bytes_compact(int64_t n_local, int32_t n_hvg, int32_t nval, int64_t nnz_local) {
    const size_t base = ab_irlb_ws_bytes(n_local, n_hvg, nval);
    if (!irlb_compact_eligible(n_local, n_hvg, nnz_local)) return base;
    return base + irlb_compact_layout(n_local, nnz_local, nullptr, 0, nullptr) + 256;
}

extern "C" int ab_irlb(ab_ctx *ctx, const int64_t *rowptr_h, const int32_t *col_h, const double *val_h,
                       int64_t n_local, int64_t n_total, int32_t n_hvg, int32_t nval, double tol, int32_t maxit,
                       const double *v0_local, const double *center, double *scores_local, double *v_local, double *s_out, double *u_out,
                       int64_t *h_info, void *ws, size_t ws_bytes, void *stream) {
    AB_REQUIRE(ctx, "ab_irlb: null ctx");
    AB_REQUIRE(n_hvg >= 2 && n_total >= 2, "ab_irlb: the input matrix must be at least 2x2 (irlb.py:119-120)");
    AB_REQUIRE(nval >= 1, "ab_irlb: nval must be >= 1");
    const int work = irlb_work_dim(nval, n_total);
    AB_REQUIRE(work <= IR_MAXWORK, "ab_irlb: work dimension %d exceeds %d (nval too large)", work, IR_MAXWORK);
    AB_REQUIRE(work >= nval + 1 && work - 3 >= 1, "ab_irlb: nval=%d too large for this matrix (work=%d)", nval, work);
    AB_REQUIRE(av_warps_for(n_hvg) >= 1, "ab_irlb: n_hvg=%d too large for the shared-memory accumulators", n_hvg);
    IrlbWs w;
    size_t need = irlb_layout(n_local, n_hvg, nval, n_total, ctx->sm_count, ws, ws_bytes, &w);
    AB_REQUIRE(need != 0, "ab_irlb: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int H = n_hvg;
    const int64_t n = n_local, ld = irlb_ld(n_local), ldw = irlb_ld(H);
    int *cnt = ctx->d_counter;
    // small all-reduces: fused one-shot pushes over NVLink peer memory when the exchange buffers are mapped
    // (ab_peer_open), ncclAllReduce between the kernels otherwise
    const bool fused = ctx->world > 1 && ctx->peer_ready && H <= AB_PEER_VEC && !getenv("AB_IRLB_NCCL");
    const AbPeer *peer = fused ? (const AbPeer *)ctx->d_peer : nullptr;      // device copy of the peer table; NULL: no exchange
    uint32_t *seq = ctx->peer_seq;

    ab_prof_begin(ctx, AB_PROF_IRLB_TOTAL, st);
    AB_CUDA(cudaMemsetAsync(w.V, 0, (size_t)ld * work * 8, st));
    AB_CUDA(cudaMemsetAsync(w.V2, 0, (size_t)ld * work * 8, st));
    AB_CUDA(cudaMemsetAsync(w.W, 0, (size_t)ldw * work * 8, st));
    AB_CUDA(cudaMemsetAsync(w.W2, 0, (size_t)ldw * work * 8, st));
    AB_CUDA(cudaMemsetAsync(w.B, 0, (size_t)work * work * 8, st));
    AB_CUDA(cudaMemsetAsync(w.scal, 0, SC_COUNT * 8, st));
    AB_CUDA(cudaMemsetAsync(w.ctrl, 0, CT_COUNT * 4, st));

    // compact row-dot operand when the caller's workspace has room for it (ab_irlb_ws_bytes_compact)
    bool compact = false;
    IrlbCompact cp{};
    if (n > 0 && (size_t)H * 8 <= (size_t)48 * 1024 && !getenv("AB_IRLB_PLAIN")) {
        int64_t nnz = 0;
        AB_CUDA(cudaMemcpyAsync(&nnz, rowptr_h + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        AB_CUDA(cudaStreamSynchronize(st));
        const size_t tail_off = (need + 255) / 256 * 256;
        if (irlb_compact_eligible(n, H, nnz) && ws_bytes > tail_off) {
            const size_t got = irlb_compact_layout(n, nnz, (char *)ws + tail_off, ws_bytes - tail_off, &cp);
            if (got != 0) {
                ab_prof_begin(ctx, AB_PROF_IRLB_ATW, st);
                AB_CUDA(cudaMemsetAsync(cp.cursor, 0, 4 * sizeof(unsigned int), st));
                const unsigned g_cb = ab_wave_grid(ctx, ah_compact_kernel, CB_WARPS * 32, 0, (n + CB_WARPS - 1) / CB_WARPS);
                ah_compact_kernel<<<g_cb, CB_WARPS * 32, 0, st>>>(rowptr_h, col_h, val_h, n, cp.ent, cp.tabv, cp.taboff,
                                                                  cp.cursor);
                AB_LAUNCH_CHECK(ctx);
                ab_prof_end(ctx, AB_PROF_IRLB_ATW, st);
                unsigned int h_cur[2] = {0, 0};
                AB_CUDA(cudaMemcpyAsync(h_cur, cp.cursor, sizeof(h_cur), cudaMemcpyDeviceToHost, st));
                AB_CUDA(cudaStreamSynchronize(st));
                compact = h_cur[1] == 0;                             // a row with > CB_CAP distinct values: plain kernels
            }
        }
    }
    // A.x operand: the gene-split copy when the workspace tail holds it, the caller's CSR otherwise
    AvOperand<int32_t> avop{};
    avop.rowptr[0] = rowptr_h; avop.col[0] = col_h; avop.val[0] = val_h; avop.halves = 1; avop.hs = H;
    AvOperand<uint16_t> avop2{};
    int av_nw = w.av_warps;
    if (cp.col2 && H >= 64 && av_split_hs(H) <= 65536 && av_warps_split(H) >= 1) {
        const int hs = av_split_hs(H);
        ab_prof_begin(ctx, AB_PROF_IRLB_AV, st);
        ah_split_count_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rowptr_h, col_h, n, hs, cp.rp0);
        AB_LAUNCH_CHECK(ctx);
        if (ab_exclusive_scan_i64(ctx, cp.rp0, cp.rp0, n, cp.scan, cp.scan_bytes, st)) return 1;
        const unsigned g_sp = ab_wave_grid(ctx, ah_split_fill_kernel, 256, 0, (n + 8) / 8);
        ah_split_fill_kernel<<<g_sp, 256, 0, st>>>(rowptr_h, col_h, val_h, n, hs, cp.rp0, cp.rp1, cp.col2, cp.val2);
        AB_LAUNCH_CHECK(ctx);
        ab_prof_end(ctx, AB_PROF_IRLB_AV, st);
        avop2.rowptr[0] = cp.rp0; avop2.rowptr[1] = cp.rp1;
        avop2.col[0] = avop2.col[1] = cp.col2;
        avop2.val[0] = avop2.val[1] = cp.val2;
        avop2.halves = 2; avop2.hs = hs;
        av_nw = av_warps_split(H);
    }
    const size_t av_smem = avop2.halves == 2 ? (size_t)2 * av_nw * avop2.hs * 8 : (size_t)av_nw * H * 8;
    AB_CUDA(cudaFuncSetAttribute(av_kernel<AV_U_WIDE, 384, int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)av_smem));
    AB_CUDA(cudaFuncSetAttribute(av_kernel<AV_U_SPLIT, 768, uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)av_smem));
    // wstep: one cluster of WS_CL CTAs; the W slice is cached in shared memory when it fits
    const int ws_nsp = (H + WS_CL - 1) / WS_CL + 1;
    const int cache_w = ((size_t)ws_nsp * (work + 1)) * 8 <= (size_t)200 * 1024 ? 1 : 0;
    const size_t ws_smem = (size_t)ws_nsp * (cache_w ? work + 1 : 1) * 8;
    AB_CUDA(cudaFuncSetAttribute(wstep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ws_smem));
    const int svd_mp = (work + 1) & ~1;
    const size_t svd_sched = (size_t)(svd_mp - 1) * (svd_mp / 2) * sizeof(unsigned short) + 16;      // pair schedule
    const int q_in_smem = ((size_t)work * (work | 1) + (size_t)work * work) * 8 + svd_sched <= (size_t)200 * 1024 ? 1 : 0;
    const size_t svd_smem = ((size_t)work * (work | 1) + (q_in_smem ? (size_t)work * work : 0)) * 8 + svd_sched;
    AB_CUDA(cudaFuncSetAttribute(svd_restart_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)svd_smem));
    const bool svd_one_cta = getenv("AB_IRLB_SVD1") != nullptr;       // A/B switch: the one-CTA Jacobi (42.6 ms per C3 pass against 32.9)
    const size_t svc_smem = (size_t)(((work + 1) & ~1) - 1) * (((work + 1) & ~1) / 2) * sizeof(unsigned short);   // pair schedule
    AB_CUDA(cudaFuncSetAttribute(svd_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)svc_smem));
    auto ritz = [&](const double *X, int64_t ldx, int64_t rows, const double *M, int ncols, double *out, int64_t ldo,
                    int rowmajor, const double *colscale, double *out2) -> int {
        if (rows <= 0 || ncols <= 0) return 0;
        const RitzCfg rc = ritz_cfg(work, ncols, ctx->smem_optin);
        AB_REQUIRE(rc.st >= 2, "ab_irlb: work dimension %d leaves no room for the Ritz pipeline", work);
        AB_REQUIRE(rc.smem <= ctx->smem_optin, "ab_irlb: work dimension %d needs %zu bytes of shared memory", work, rc.smem);
        ritz_fn fn = ritz_pick(rc.ct);
        AB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rc.smem));
        const int64_t tiles = (rows + RZ_TR - 1) / RZ_TR;
        fn<<<(unsigned)std::min<int64_t>(tiles, ctx->sm_count), RZ_THREADS, rc.smem, st>>>(X, ldx, rows, work, M, ncols, out,
                                                                                          ldo, rowmajor, colscale, out2,
                                                                                          rc.kc, rc.st);
        AB_LAUNCH_CHECK(ctx);
        return 0;
    };
    auto wstep = [&](int jd, int ncols, int sub_prev, int jb, uint32_t seq_vec) -> int {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(WS_CL);
        cfg.blockDim = dim3(WS_THREADS);
        cfg.dynamicSmemBytes = ws_smem;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = WS_CL;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        ab_prof_begin(ctx, AB_PROF_IRLB_SMALL, st);
        AB_CUDA(cudaLaunchKernelEx(&cfg, wstep_kernel, (const double *)w.wraw, w.W, H, ldw, jd, ncols, sub_prev, w.scal, w.B,
                                   work, jb, cache_w, peer, seq_vec));
        AB_LAUNCH_CHECK(ctx);
        if (center) {                                       // mu^T w of the new left vector, read by the next A^T.w
            muw_kernel<<<1, 256, 0, st>>>(center, w.W + (size_t)jd * ldw, H, w.scal + SC_MUW);
            AB_LAUNCH_CHECK(ctx);
        }
        ab_prof_end(ctx, AB_PROF_IRLB_SMALL, st);
        return 0;
    };
    const unsigned g_rows = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + 7) / 8, (int64_t)ctx->sm_count * 32));
    const unsigned g_vec = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 8));
    // at least two rows per thread (a 163 k-row shard on 592 CTAs left one row per thread: the per-CTA reductions and the
    // 592-partial tail were the kernel; 17.1 -> 15.3 ms per pass at that size), never more CTAs than the layout sized
    const unsigned g_dot = (unsigned)std::max<int64_t>(1, std::min<int64_t>(w.dot_grid, (n + 2 * DOT_THREADS - 1) / (2 * DOT_THREADS)));

    // compact A^T.w: 8 resident CTAs per SM, each with its own copy of w in shared memory
    AB_CUDA(cudaFuncSetAttribute(atw_compact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)H * 8)));
    const unsigned g_atw = ab_wave_grid(ctx, atw_compact_kernel, 256, (size_t)H * 8, (n + 7) / 8);
    int64_t nmult = 0, sum_panel = 0;

    // A.x, summed over the CTAs (and, sharded, pushed to every rank); returns the instance id the consumer waits for
    uint32_t seq_fn2 = 0;                                    // instance of the last ||F||^2 exchange
    auto av = [&](const double *src, bool scaled, double *vdst, uint32_t *seq_vec_out) -> int {
        if (center) {
            // implicit centring needs 1^T x over all cells (x = src, scaled inside av_reduce_kernel)
            vecsum_kernel<<<(unsigned)std::min<int64_t>(1024, std::max<int64_t>(1, (n + 255) / 256)), DOT_THREADS, 0, st>>>(
                src, n, w.nrm_partial, w.scal + SC_SUMX, cnt + 4);
            AB_LAUNCH_CHECK(ctx);
            if (ab_allreduce_f64(ctx, w.scal + SC_SUMX, 1, st)) return 1;
        }
        ab_prof_begin(ctx, AB_PROF_IRLB_AV, st);
        // (A.x keeps the plain operand: its 12 warps per SM leave ~30 KB of L1, a table look-up per nonzero at consume
        // time misses it, and the compact variant measured 344 ms against 195)
        if (avop2.halves == 2)
            av_kernel<AV_U_SPLIT, 768, uint16_t><<<w.av_grid, 64 * av_nw, av_smem, st>>>(
                avop2, n, H, src, scaled ? w.scal : nullptr, vdst, av_nw, w.av_partial, peer, seq_fn2);
        else
            av_kernel<AV_U_WIDE, 384, int32_t><<<w.av_grid, 384, av_smem, st>>>(
                avop, n, H, src, scaled ? w.scal : nullptr, vdst, av_nw, w.av_partial, peer, seq_fn2);
        AB_LAUNCH_CHECK(ctx);
        const uint32_t sv = fused ? ++seq[AB_CH_VEC] : 0;
        av_reduce_kernel<<<(H + 31) / 32, 128, 0, st>>>(w.av_partial, w.av_grid, H, w.wraw, cnt + 3, peer, sv,
                                                          (center && ctx->rank == 0) ? center : nullptr, w.scal, scaled ? 1 : 0);
        AB_LAUNCH_CHECK(ctx);
        ab_prof_end(ctx, AB_PROF_IRLB_AV, st);
        if (!fused && ab_allreduce_f64(ctx, w.wraw, H, st)) return 1;
        *seq_vec_out = sv;
        nmult++;
        return 0;
    };

    // start vector (irlb.py:151-153)
    sumsq_kernel<<<(unsigned)std::min<int64_t>(1024, std::max<int64_t>(1, (n + 255) / 256)), DOT_THREADS, 0, st>>>(
        v0_local, n, w.nrm_partial, w.scal + SC_V0N2, cnt);
    AB_LAUNCH_CHECK(ctx);
    if (ab_allreduce_f64(ctx, w.scal + SC_V0N2, 1, st)) return 1;
    scale_by_invnorm_kernel<<<g_vec, 256, 0, st>>>(v0_local, w.V, n, w.scal + SC_V0N2);
    AB_LAUNCH_CHECK(ctx);

    int it = 0, keep = nval;
    int h_ctrl[CT_COUNT] = {0};
    bool converged = false;
    while (it < maxit) {
        int j = it > 0 ? keep : 0;
        uint32_t sv = 0;
        // W[:,j] = A.V[:,j]  (irlb.py:165), orthogonalised against W[:, :j] on restarts (:173-175)
        if (av(w.V + (int64_t)j * ld, false, nullptr, &sv)) return 1;
        if (wstep(j, it > 0 ? j : 0, 0, -1, sv)) return 1;
        while (j < work) {
            ab_prof_begin(ctx, AB_PROF_IRLB_ATW, st);
            if (compact)
                atw_compact_kernel<<<g_atw, 256, (size_t)H * 8, st>>>(rowptr_h, cp.ent, cp.tabv, cp.taboff, n, H,
                                                                       w.W + (size_t)j * ldw, w.V + (int64_t)j * ld, w.scal, w.F);
            else
                atw_kernel<<<g_rows, 256, 0, st>>>(rowptr_h, col_h, val_h, n, w.W + (size_t)j * ldw, w.V + (int64_t)j * ld,
                                                    w.scal, w.F);
            AB_LAUNCH_CHECK(ctx);
            ab_prof_end(ctx, AB_PROF_IRLB_ATW, st);
            nmult++;
            ab_prof_begin(ctx, AB_PROF_IRLB_PANEL, st);
            const uint32_t sc = fused ? ++seq[AB_CH_COEF] : 0;
            panel_dot_kernel<<<g_dot, DOT_THREADS, 0, st>>>(w.V, ld, n, j + 1, w.F, w.dot_partial, w.coef, cnt + 1, peer, sc);
            AB_LAUNCH_CHECK(ctx);
            if (!fused && ab_allreduce_f64(ctx, w.coef, j + 1, st)) return 1;
            seq_fn2 = fused ? ++seq[AB_CH_SCAL] : 0;
            panel_update_kernel<<<g_dot, DOT_THREADS, 0, st>>>(w.V, ld, n, j + 1, w.coef, w.F, w.nrm_partial,
                                                                w.scal + SC_FN2, cnt + 2, peer, sc, seq_fn2);
            AB_LAUNCH_CHECK(ctx);
            ab_prof_end(ctx, AB_PROF_IRLB_PANEL, st);
            if (!fused && ab_allreduce_f64(ctx, w.scal + SC_FN2, 1, st)) return 1;
            sum_panel += j + 1;
            if (j < work - 1) {
                // V[:,j+1] = F/fn ; W[:,j+1] = A.V[:,j+1] - fn W[:,j], CGS, normalise (irlb.py:195-219)
                if (av(w.F, true, w.V + (int64_t)(j + 1) * ld, &sv)) return 1;
                if (wstep(j + 1, j + 1, 1, j, sv)) return 1;
            } else {
                fscale_kernel<<<g_vec, 256, 0, st>>>(w.F, n, w.scal, w.B, work, j, peer, seq_fn2);
                AB_LAUNCH_CHECK(ctx);
            }
            ++j;
        }
        ab_prof_begin(ctx, AB_PROF_IRLB_SVD, st);
        if (svd_one_cta) {
            svd_restart_kernel<<<1, (8 * (svd_mp / 2) + 31) / 32 * 32, svd_smem, st>>>(w.B, work, w.Ub, w.sb, w.Vb, w.G, w.scal, w.ctrl, w.R, nval, tol, it,
                                                         keep, q_in_smem);
        } else {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(SV_CL);
            cfg.blockDim = dim3(SV_THREADS);
            cfg.dynamicSmemBytes = svc_smem;
            cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = SV_CL;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            AB_CUDA(cudaLaunchKernelEx(&cfg, svd_cluster_kernel, (const double *)w.B, work, w.Ub, w.sb, w.Vb, w.G, w.scal,
                                       w.ctrl, w.R, (int)nval, tol, it, keep));
        }
        AB_LAUNCH_CHECK(ctx);
        ab_prof_end(ctx, AB_PROF_IRLB_SVD, st);
        AB_CUDA(cudaMemcpyAsync(h_ctrl, w.ctrl, sizeof(h_ctrl), cudaMemcpyDeviceToHost, st));
        AB_CUDA(cudaStreamSynchronize(st));
        if (getenv("AB_IRLB_DEBUG")) fprintf(stderr, "[irlb] restart %d: jacobi sweeps %d, converged %d, keep %d\n", it, h_ctrl[CT_ROT], h_ctrl[CT_CONV], h_ctrl[CT_KEEP]);
        if (h_ctrl[CT_DONE]) { converged = true; break; }
        keep = h_ctrl[CT_KEEP];
        // Ritz vectors (irlb.py:238-246)
        ab_prof_begin(ctx, AB_PROF_IRLB_RITZ, st);
        if (ritz(w.V, ld, n, w.Vb, keep, w.V2, ld, 0, nullptr, nullptr)) return 1;
        copy_vec_kernel<<<g_vec, 256, 0, st>>>(w.F, w.V2 + (int64_t)keep * ld, n);
        AB_LAUNCH_CHECK(ctx);
        if (ritz(w.W, ldw, H, w.Ub, keep, w.W2, ldw, 0, nullptr, nullptr)) return 1;
        reset_b_kernel<<<8, 256, 0, st>>>(w.B, work, w.sb, w.R, keep);
        AB_LAUNCH_CHECK(ctx);
        ab_prof_end(ctx, AB_PROF_IRLB_RITZ, st);
        std::swap(w.V, w.V2);
        std::swap(w.W, w.W2);
        ++it;
    }
    // outputs (irlb.py:249-257; clustering.py:111)
    if (ritz(w.V, ld, n, w.Vb, nval, scores_local, 0, 1, w.sb, v_local)) return 1;
    if (u_out) {
        if (ritz(w.W, ldw, H, w.Ub, nval, u_out, 0, 1, nullptr, nullptr)) return 1;
    }
    if (s_out) AB_CUDA(cudaMemcpyAsync(s_out, w.sb, (size_t)nval * 8, cudaMemcpyDeviceToDevice, st));
    ab_prof_end(ctx, AB_PROF_IRLB_TOTAL, st);
    AB_CUDA(cudaStreamSynchronize(st));
    if (h_info) {
        h_info[0] = it;
        h_info[1] = nmult;
        h_info[2] = sum_panel;
        h_info[3] = (converged ? 0 : 1) | (compact ? 2 : 0) | (fused ? 4 : 0);
    }
    return 0;
}
