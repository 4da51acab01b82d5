// K9 tensor-core candidate pass: tiled  s(q,x) = ||x||^2 - 2 q.x  on tcgen05 (FP16 inputs, FP32
// accumulation in TMEM) with a per-query top-m selection epilogue, followed by an exact FP64
// re-rank that reproduces the reference's number and certifies the result.
//
// Replaces the search of AlonaClustering.knn (alona/clustering.py:184-186, scikit-learn BallTree).
//
//  prep      centre the embedding (translation invariant, shrinks norms), scale by a power of two
//            so that the largest norm lies in [32,64) (keeps every operand inside FP16 range,
//            exact), round to FP32 and split every coordinate into two FP16 pieces
//            (hi = rn(x), lo = rn(x-hi): 22 significant bits).  The DATABASE side carries both pieces, the
//            QUERY side only q_hi: the GEMM then measures exact distances to the perturbed query q_hi
//            (an FP16 vector, so nothing is lost on that side), and |q - q_hi| - known per query - enters the
//            certificate through the triangle inequality.  Two FP16 products per coordinate pair instead of
//            the three a 3xTF32-style split needs (K = 2d+3 instead of 3d+3).  ONE GEMM emits finished
//            scores, norms folded into three extra K columns:
//               A[q] = [  q_hi |  q_hi | 1 1 1 | 0.. ]
//               B[x] = [-2x_hi |-2x_lo | n0 n1 n2 | 0.. ],  n0+n1+n2 = ||x||^2
//  main      persistent warp-specialised kernel, one CTA per SM:
//              warp 0   TMA producer: B arrives as contiguous 8 KB bulk copies (one 128-point x 32-half K chunk,
//                       stored pre-swizzled in HBM) through an mbarrier ring; A by 2-D tensor-map loads;
//              warp 1   tcgen05.mma issuer (kind::f16, M=128, N=128): every B chunk is multiplied
//                       against TWO resident 128-query A tiles (256 queries per CTA halve the L2->SMEM
//                       operand traffic per flop); two 256-column TMEM buffers so the epilogue of
//                       tile t overlaps the MMAs of tile t+1;
//              warps 2-9 epilogue: thread <-> query row (= TMEM lane), 32 columns per tcgen05.ld,
//                       two loads per wait; a min-tree + one
//                       compare per 32 scores; on a hit the 32 scores are staged to shared memory,
//                       a hit mask is built, and only the hits are appended to a per-thread buffer
//                       that a warp-cooperative bitonic merge folds into the sorted m-list.
//            For wide embeddings (A tiles of 256 queries do not fit) the same kernel runs with ONE
//            query tile against 256-point B tiles (two lists per query, merged by the re-rank).
//  re-rank   one warp per query: exact distances of the candidates with the reference's
//            arithmetic, sort by (distance, index), certificate
//                 sqrt(tau + ||q_hi||^2 - eps) - |q - q_hi|  >  sqrt(D_k)      (tau = m-th smallest approximate score)
//            which proves that no point outside the candidate list can enter the top k: such a point has
//            |x - q_hi|^2 >= tau + ||q_hi||^2 - eps and |x - q| >= |x - q_hi| - |q - q_hi|;
//            eps = KAPPA (||q|| + R)^2 bounds FP32 rounding + split + accumulation error.
//            Uncertified queries are recomputed by the exact FP64 scan (ab_knn_exact.cu).
#include "ab_common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <algorithm>

size_t ab_knn_exact_ws_bytes(int64_t nq, int k, int splits);
int ab_knn_exact_pick_splits(ab_ctx *ctx, int64_t nq);
int ab_knn_exact_run(ab_ctx *ctx, const double *emb, int64_t n_total, int d, const int64_t *qlist, int64_t q_begin,
                     int64_t nq, int k, int splits, const int64_t *out_rows, int32_t *idx, double *rdist,
                     unsigned long long *stats, void *ws, size_t ws_bytes, cudaStream_t st);

static constexpr int TC_TILE = 128;          // UMMA M and N
static constexpr int TC_KCH = 32;            // halves per K chunk = one 64-byte swizzle row
static constexpr int TC_SLAB = TC_TILE * 64; // bytes of one 128-row x 64-byte operand slab (8 KB)
static constexpr int TC_MAX_STAGES = 16;
static constexpr int TC_MAX_CHUNKS = 10;     // 2d+3 <= 320  ->  d <= 158
static constexpr int TC_NPROD = 2;           // FP16 products per coordinate pair: q_hi.x_hi + q_hi.x_lo
// Error bound of an approximate score as a distance to q_hi, relative to (|q| + R)^2, that the certificate uses:
//   * accumulation: the K products of a score are exact in FP32 (two 11-bit significands) and are added in FP32:
//     at most K roundings of 2^-24 relative to sum |a_i b_i| <= 2|q|R + R^2 <= (|q| + R)^2 (any summation order,
//     so it also covers the tensor pipe's internal one);
//   * database side: each coordinate carries 22 significant bits (hi + lo); what x_hi + x_lo misses of x, the FP32
//     rounding of the coordinates and the FP16 rounding of the three norm pieces add less than 2^-20 (|q| + R)^2.
// kappa = 1.25 K 2^-24 + 2^-19 (a quarter of margin on the first term; K = 112 at d = 50 gives 1.03e-5).  The query
// side has no term here: q_hi is exact in FP16 and the distance between q and q_hi is handled exactly (see above).
// The largest error observed on the candidates is reported (stats[4]) and ab_knn_tc_run refuses to certify anything
// if it ever exceeds kappa / 2.
// (Round 1 multiplied three products - q_lo.x_hi as well - at K = 3d+3: 507 ms against 421 ms at C3 for the same table.
// One product, q_hi.x_hi, would need max |x - x_hi| in the certificate as well and no longer certifies.)
__host__ __device__ inline double tc_kappa(int kt) {
    return 1.25 * (double)kt * 5.9604644775390625e-08 + 1.9073486328125e-06;
}

// ---------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug must fault, not hang the GPU
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 28)) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// exactly one lane of the (converged) warp gets true; ptxas keeps the code under it on uniform registers
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
// 1-D bulk copy global -> shared (UBLKCP), completion on an mbarrier
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// pins the uses of registers filled by an asynchronous tcgen05.ld behind the wait that precedes this call
__device__ __forceinline__ void tc_touch32(uint32_t (&r)[32]) {
    asm volatile(""
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}

// K-major, 64-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// rows of 64 bytes, 8-row groups 512 bytes apart
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);          // start address
    d |= (uint64_t)1 << 16;                           // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(512 >> 4) << 32;                  // stride byte offset: 8 rows x 64 B
    d |= (uint64_t)1 << 46;                           // descriptor version (sm_100)
    d |= (uint64_t)4 << 61;                           // SWIZZLE_64B
    return d;
}

// power-of-two scale that puts the largest centred norm R into [32, 64)
__device__ __forceinline__ double knn_scale_of(float R) {
    if (!(R > 0.f) || !isfinite(R)) return 1.0;
    return ldexp(1.0, 5 - ilogbf(R));
}

// ---------------------------------------------------------------------------------------
// prep
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
knn_colsum_kernel(const double *__restrict__ emb, int64_t n, int d, double *__restrict__ partial /*[grid][d]*/) {
    __shared__ double sm[8][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = 0.0;
    const int64_t per = (n + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * per, r1 = min(n, r0 + per);
    for (int64_t r = r0 + warp; r < r1; r += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j = lane + 32 * u;
            if (j < d) acc[u] += emb[r * d + j];
        }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) sm[warp][lane + 32 * u] = acc[u];
    __syncthreads();
    for (int j = threadIdx.x; j < d; j += 256) {
        double t = 0.0;
        for (int w2 = 0; w2 < 8; ++w2) t += sm[w2][j];
        partial[(size_t)blockIdx.x * d + j] = t;
    }
}

__global__ void knn_colmean_kernel(const double *__restrict__ partial, int nparts, int d, int64_t n, double *__restrict__ mean) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= d) return;
    double t = 0.0;
    for (int b = 0; b < nparts; ++b) t += partial[(size_t)b * d + j];
    mean[j] = t / (double)n;
}

// largest centred norm (one warp per point), as float bits rounded up
__global__ void __launch_bounds__(256)
knn_rmax_kernel(const double *__restrict__ emb, int64_t n, int d, const double *__restrict__ mean,
                unsigned int *__restrict__ rmax_bits) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * 8;
    float best = 0.f;
    for (int64_t i = w; i < n; i += nw) {
        double n2 = 0.0;
        for (int j = lane; j < d; j += 32) {
            const double x = emb[i * d + j] - mean[j];
            n2 += x * x;
        }
        n2 = ab_warp_sum_f64(n2);
        best = fmaxf(best, (float)sqrt(n2) * 1.000001f);
    }
    if (lane == 0) atomicMax(rmax_bits, __float_as_uint(best));
}

// Position (in halves) of element (row i, K column j) of the B operand.  B is stored as the exact
// shared-memory image the MMA reads: slabs of 128 rows, inside a slab one 8 KB block per K chunk
// (128 rows x 64 B) with the 64-byte swizzle (16-byte unit index ^= (row>>1)&3) already applied, so
// a pipeline stage is ONE contiguous 8 KB bulk copy instead of 128 strided 64-byte row requests
// (a 2-D tensor-map load of such narrow rows is bound by the TMA's per-row request rate).
__host__ __device__ __forceinline__ size_t knn_b_pos(int64_t i, int j, int chunks) {
    const int64_t slab = i >> 7;
    const int r = (int)(i & 127), c = j >> 5, jj = j & 31;
    return ((size_t)slab * chunks + c) * 4096 + (size_t)r * 32 + ((((jj >> 3) ^ ((r >> 1) & 3)) << 3) | (jj & 7));
}

// one warp per point: writes the FP16 A and B operand rows and ||x~||^2 (unscaled units).
// B has n_pad >= n rows (a whole number of tiles): padding rows are all zero except a norm of
// 60000, larger than any real score (<= (2*64)^2), so padded columns never pass a threshold.
__global__ void __launch_bounds__(256)
knn_prep_kernel(const double *__restrict__ emb, int64_t n, int64_t n_pad, int d, int kt, const double *__restrict__ mean,
                const unsigned int *__restrict__ rmax_bits, __half *__restrict__ aop, __half *__restrict__ bop,
                double *__restrict__ qn, double *__restrict__ qh2 /* |q_hi|^2 */, double *__restrict__ qdel /* |q - q_hi| */) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * 8;
    const double sc = knn_scale_of(__uint_as_float(*rmax_bits));
    const __half zero = __float2half_rn(0.f), one = __float2half_rn(1.f);
    const int chunks = kt / TC_KCH;
    for (int64_t i = w; i < n; i += nw) {
        __half *a = aop + i * kt;
        double n2 = 0.0, h2 = 0.0, dl2 = 0.0;
        const int nd = TC_NPROD * d;                             // first norm column
        for (int j = nd + lane; j < kt; j += 32) {               // clear the tail (norm slots are rewritten below)
            a[j] = zero; bop[knn_b_pos(i, j, chunks)] = zero;
        }
        for (int j = lane; j < d; j += 32) {
            const double xs = (emb[i * d + j] - mean[j]) * sc;
            const float x = (float)xs;
            const __half hi = __float2half_rn(x);
            const double hd = (double)__half2float(hi);
            h2 += hd * hd;
            dl2 += (xs - hd) * (xs - hd);
            const __half lo = __float2half_rn(x - __half2float(hi));
            const __half mhi = __float2half_rn(-2.f * __half2float(hi)), mlo = __float2half_rn(-2.f * __half2float(lo));
            // q_hi.(x_hi + x_lo)
            a[j] = hi;
            bop[knn_b_pos(i, j, chunks)] = mhi;
            a[d + j] = hi;
            bop[knn_b_pos(i, d + j, chunks)] = mlo;
            n2 += (double)x * (double)x;
        }
        n2 = ab_warp_sum_f64(n2);
        h2 = ab_warp_sum_f64(h2);
        dl2 = ab_warp_sum_f64(dl2);
        __syncwarp();
        if (lane == 0) {
            qh2[i] = h2 / (sc * sc);
            qdel[i] = sqrt(dl2) / sc * (1.0 + 1e-12);
            const __half n0 = __double2half(n2);
            const double r1 = n2 - (double)__half2float(n0);
            const __half n1 = __double2half(r1);
            const __half n2h = __double2half(r1 - (double)__half2float(n1));
            bop[knn_b_pos(i, nd, chunks)] = n0;
            bop[knn_b_pos(i, nd + 1, chunks)] = n1;
            bop[knn_b_pos(i, nd + 2, chunks)] = n2h;
            a[nd] = one; a[nd + 1] = one; a[nd + 2] = one;
            qn[i] = n2 / (sc * sc);
        }
    }
    for (int64_t i = n + w; i < n_pad; i += nw) {
        for (int j = lane; j < kt; j += 32) bop[knn_b_pos(i, j, chunks)] = (j == TC_NPROD * d) ? __float2half_rn(60000.f) : zero;
    }
}

// ---------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------
static constexpr int TC_EPI_WARPS = 8;                       // two halves of four warps
static constexpr int TC_EPI_THREADS = TC_EPI_WARPS * 32;
static constexpr int TC_NTHREADS = 64 + TC_EPI_THREADS;      // warp0 TMA, warp1 MMA, warps 2..9 epilogue
static constexpr int TC_GRP = 8;                             // points per candidate group (4 or 8; 4: main +5 %, re-rank -50 %, same total)
static constexpr int TC_BUF = 32;                            // per-thread append buffer (entries; one prune merges <= 32)
static constexpr unsigned long long TC_KEY_MAX = 0xFFFFFFFFFFFFFFFFull;

// candidate key = order-preserving bits of the FP32 score in the high word, point id in the low word
__device__ __forceinline__ unsigned long long cand_key(float s, int idx) {
    const uint32_t b = __float_as_uint(s);
    const uint32_t o = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return ((unsigned long long)o << 32) | (uint32_t)idx;
}
__host__ __device__ __forceinline__ float key_score(unsigned long long key) {
    const uint32_t o = (uint32_t)(key >> 32);
    const uint32_t b = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
#ifdef __CUDA_ARCH__
    return key == TC_KEY_MAX ? INFINITY : __uint_as_float(b);
#else
    float f; memcpy(&f, &b, 4); return key == TC_KEY_MAX ? INFINITY : f;
#endif
}
// Low word of a candidate key: first point of the group (a multiple of TC_GRP, below 2^31) with the pair mask in the bits
// that leaves free - the low bits and bit 31.  (The mask only perturbs the order of keys with EQUAL scores.)
__host__ __device__ __forceinline__ uint32_t tc_key_low(int first, unsigned pair_mask) {
    constexpr unsigned LOWB = TC_GRP == 8 ? 3 : 2;
    return (uint32_t)first | (pair_mask & ((1u << LOWB) - 1u)) | ((pair_mask >> LOWB) << 31);
}
__host__ __device__ __forceinline__ int tc_key_first(unsigned long long key) {
    return (int)((uint32_t)key & 0x7FFFFFFFu & ~(uint32_t)(TC_GRP - 1));
}
__host__ __device__ __forceinline__ unsigned tc_key_pairs(unsigned long long key) {
    constexpr unsigned LOWB = TC_GRP == 8 ? 3 : 2;
    const uint32_t lo = (uint32_t)key;
    return (lo & ((1u << LOWB) - 1u)) | ((lo >> 31) << LOWB);
}
__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

struct TcParams {
    int64_t n_total;       // database points
    int64_t q_begin;       // first query (row of the operand arrays)
    int64_t nq;            // number of queries
    int chunks;            // K chunks of 32 halves
    int ksteps;            // UMMA K steps (16 halves) that carry data: ceil((3d+3)/16)
    int stages;            // B ring depth, in slots of `group` chunks
    int group;             // K chunks per ring slot = per bulk copy and per barrier round
    int split;             // 1: the two 128-column halves of an accumulator buffer are committed and released
                           //    separately (needs one ring slot per tile): the epilogue of half 0 starts while the
                           //    MMAs of half 1 run, and a half is reusable as soon as ITS four warps are done
    int mt;                // 2: two query tiles x one 128-point B tile; 1: one query tile x 256-point B tile
    int m;                 // candidates kept per list
    unsigned long long *cand;   // [lists][nq][m] sorted keys, lists = 3 - mt
    const unsigned char *bop;   // tiled, pre-swizzled B operand (knn_b_pos)
    unsigned long long *bufg;   // [grid][TC_BUF][256] append buffers
    int *cursor;                // tile CTA 0 is currently loading: the other CTAs start their sweeps there
    // Work items of a CTA: `full` whole sweeps (query block blockIdx.x + i * gridDim.x against every tile), then the items
    // of the TAIL phase.  The query blocks left over when the blocks do not divide by the grid would keep a few SMs busy
    // for a whole sweep while the others idle (163,265 queries per rank at 8 GPUs: 4.31 rounds cost 5); instead each of
    // the R tail blocks is cut into S parts, a contiguous S-th of the tiles each, the R * S parts are dealt round-robin
    // over the CTAs, and every part fills its own candidate list (cand_tail) that knn_tail_merge_kernel folds into the
    // block's ordinary list afterwards.  The host picks (full, R, S) for the shortest schedule (tc_schedule).
    int full, R, S;
    unsigned long long *cand_tail;   // [S * lists][R * q_per_block][m]
};

struct TcItem { int64_t qb; int tbeg, tcnt, part, tail; };
__device__ __forceinline__ bool tc_item(const TcParams &P, int i, int n_tiles, TcItem &o) {
    if (i < P.full) {
        o.qb = blockIdx.x + (int64_t)i * gridDim.x; o.tbeg = 0; o.tcnt = n_tiles; o.part = 0; o.tail = 0;
        return true;
    }
    const int64_t pi = blockIdx.x + (int64_t)(i - P.full) * gridDim.x;           // part item of the tail phase
    if (pi < (int64_t)P.R * P.S) {
        const int j = (int)(pi / P.S), part = (int)(pi - (int64_t)j * P.S);
        o.qb = (int64_t)P.full * gridDim.x + j;
        o.tbeg = (int)((int64_t)part * n_tiles / P.S);
        o.tcnt = (int)((int64_t)(part + 1) * n_tiles / P.S) - o.tbeg;
        o.part = part; o.tail = P.S > 1;
        return true;
    }
    return false;
}

// 32-wide bitonic sort across the warp, descending (lane 0 ends with the largest key)
__device__ __forceinline__ unsigned long long warp_sort_desc(unsigned long long v, int lane) {
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const unsigned long long o = __shfl_xor_sync(AB_FULL, v, stride);
            const bool up = ((lane & size) == 0);            // this block sorts descending when `up`
            const bool lower = ((lane & stride) == 0);
            const unsigned long long mx = v > o ? v : o, mn = v > o ? o : v;
            v = (lower == up) ? mx : mn;
        }
    }
    return v;
}

__device__ __forceinline__ void sts_b64(uint32_t addr, unsigned long long v) {
    asm volatile("st.shared.b64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
// 64-bit shared store of (lo, hi) without assembling a 64-bit register pair first
__device__ __forceinline__ void sts_v2(uint32_t addr, uint32_t lo, uint32_t hi) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(lo), "r"(hi) : "memory");
}
__device__ __forceinline__ unsigned long long lds_b64(uint32_t addr) {
    unsigned long long v;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
    return v;
}

// Warp-cooperative prune: for every lane in `lanes`, merge its append buffer (cnt <= 32 unsorted
// keys in shared memory, entry e of epilogue thread t at buf + (e*256 + t)*8) into its sorted m-list
// (global memory, m = 32*E keys, element i lives in register x[i/32] of lane i%32), keep the m
// smallest.  Returns, in each pruned lane, its refreshed threshold (the caller resets its count).
// Bitonic split + merge entirely in registers/shuffles.  Warp-uniform call.
template <int E>
__device__ __noinline__ float warp_prune_t(unsigned lanes, unsigned long long *cand_list0, long long q_lane, int cnt,
                                           const unsigned long long *bufp, int lane) {
    float my_tau = 0.f;
    while (lanes) {
        const int L = __ffs(lanes) - 1;
        lanes &= lanes - 1;
        const int cntL = __shfl_sync(AB_FULL, cnt, L);
        const long long qL = __shfl_sync(AB_FULL, q_lane, L);
        const unsigned long long *colL = (const unsigned long long *)__shfl_sync(AB_FULL, (unsigned long long)bufp, L);
        unsigned long long *list = cand_list0 + (size_t)qL * (32 * E);
        unsigned long long x[E];
#pragma unroll
        for (int e = 0; e < E; ++e) x[e] = list[e * 32 + lane];
        unsigned long long b = TC_KEY_MAX;
        if (lane < cntL) {                                   // buffer entries are raw (score bits, point id) pairs
            const unsigned long long raw = __ldcg(colL + (size_t)lane * TC_EPI_THREADS);
            b = cand_key(__uint_as_float((uint32_t)(raw >> 32)), (int)(uint32_t)raw);
        }
        b = warp_sort_desc(b, lane);
        // bitonic split against the (virtual) upper half [MAX.. , b descending]: only the top 32 slots matter
        x[E - 1] = x[E - 1] < b ? x[E - 1] : b;
        // bitonic merge of the lower half (length 32*E)
#pragma unroll
        for (int stride = 16 * E; stride >= 32; stride >>= 1) {
            const int se = stride / 32;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                if ((e & se) == 0) {
                    const unsigned long long lo = x[e], hi = x[e + se];
                    x[e] = lo < hi ? lo : hi;
                    x[e + se] = lo < hi ? hi : lo;
                }
            }
        }
#pragma unroll
        for (int stride = 16; stride > 0; stride >>= 1) {
            const bool lower = ((lane & stride) == 0);
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const unsigned long long o = __shfl_xor_sync(AB_FULL, x[e], stride);
                const unsigned long long mn = x[e] < o ? x[e] : o, mx = x[e] < o ? o : x[e];
                x[e] = lower ? mn : mx;
            }
        }
#pragma unroll
        for (int e = 0; e < E; ++e) list[e * 32 + lane] = x[e];
        const unsigned long long kth = __shfl_sync(AB_FULL, x[E - 1], 31);
        if (lane == L) my_tau = key_score(kth);
        __syncwarp();
    }
    return my_tau;
}

// prunes the lanes in `lanes` (warp-uniform mask), updating their (cnt, tau)
__device__ __forceinline__ void warp_prune(unsigned lanes, const TcParams &P, unsigned long long *cg, long long q_lane, int &cnt,
                                           float &tau, unsigned long long *bufp, int lane) {
    float t;
    switch (P.m) {
        case 32: t = warp_prune_t<1>(lanes, cg, q_lane, cnt, bufp, lane); break;
        case 64: t = warp_prune_t<2>(lanes, cg, q_lane, cnt, bufp, lane); break;
        case 128: t = warp_prune_t<4>(lanes, cg, q_lane, cnt, bufp, lane); break;
        default: t = warp_prune_t<8>(lanes, cg, q_lane, cnt, bufp, lane); break;
    }
    if ((lanes >> lane) & 1u) { tau = fminf(tau, t); cnt = 0; }       // (a list that is not full yet returns +inf)
}

// one unit = 32 consecutive scores of this thread's query row.  Common path: a min-tree and one
// compare.  On a hit the unit is revisited in groups of 8: only groups that hold a score below the
// lane's threshold are scanned, hits go straight to the lane's append buffer (room for a whole
// group is guaranteed beforehand; lanes short of room are pruned first, which also tightens
// their threshold).  Padding columns carry a huge score (see knn_prep_kernel) and never pass.
__device__ __forceinline__ void epi_unit(const uint32_t (&r)[32], int cbase, bool live, const TcParams &P, unsigned long long *cg,
                                         long long q, int &cnt, float &tau, unsigned long long *bufp, int lane) {
    constexpr int NG = 32 / TC_GRP;
    float gmin[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        if (TC_GRP == 8) {
            float v = fmin3(__uint_as_float(r[8 * g]), __uint_as_float(r[8 * g + 1]), __uint_as_float(r[8 * g + 2]));
            v = fmin3(v, __uint_as_float(r[8 * g + 3]), __uint_as_float(r[8 * g + 4]));
            v = fmin3(v, __uint_as_float(r[8 * g + 5]), __uint_as_float(r[8 * g + 6]));
            gmin[g] = fminf(v, __uint_as_float(r[8 * g + 7]));
        } else {
            gmin[g] = fminf(fmin3(__uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1]), __uint_as_float(r[4 * g + 2])),
                            __uint_as_float(r[4 * g + 3]));
        }
    }
    float mn = gmin[0];
#pragma unroll
    for (int g = 1; g + 1 < NG; g += 2) mn = fmin3(mn, gmin[g], gmin[g + 1]);
    mn = fminf(mn, gmin[NG - 1]);
    if (!__any_sync(AB_FULL, live && mn < tau)) return;
    // Rare path.  A candidate is a GROUP of TC_GRP consecutive points keyed by its smallest score: appending it is one
    // compare and one store per group, with no search inside the 32 registers (measured: that search, executed by
    // the whole warp for the ~1.7 lanes that hit, was everything the kernel cost above its MMA/copy floor).  The
    // certificate is unchanged - every point outside the listed groups scores >= its group's minimum >= tau - and
    // the re-rank expands the listed groups to their points.
    const unsigned tight = __ballot_sync(AB_FULL, cnt > TC_BUF - NG);
    if (tight) warp_prune(tight, P, cg, q, cnt, tau, bufp, lane);
    // (a warp-uniform branch per group - skip the groups no lane lists - measured no better than this predicated form)
    if (live && mn < tau) {
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            if (gmin[g] < tau) {
                // which PAIRS of the group hold a score below the threshold (tc_key_pairs): the re-rank only computes
                // exact distances for those - a point at or above the threshold of its time can be left out under the
                // same certificate as a point of an unlisted group
                unsigned pm = 0;
#pragma unroll
                for (int pp = 0; pp < TC_GRP / 2; ++pp)
                    pm |= fminf(__uint_as_float(r[TC_GRP * g + 2 * pp]), __uint_as_float(r[TC_GRP * g + 2 * pp + 1])) < tau ? (1u << pp) : 0u;
                // append buffers live in global memory (L2): appends are rare fire-and-forget stores, and the
                // shared memory they would take is worth more as operand ring
                bufp[(size_t)cnt * TC_EPI_THREADS] = ((unsigned long long)__float_as_uint(gmin[g]) << 32) | tc_key_low(cbase + TC_GRP * g, pm);
                ++cnt;
            }
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(TC_NTHREADS, 1)
knn_tc_kernel(const __grid_constant__ CUtensorMap map_a, const TcParams P) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // layout: [A chunks][B ring: stages x group chunks][barriers]
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int a_chunk_bytes = P.mt * TC_SLAB;             // 128*mt query rows x 64 B
    const int b_stage_bytes = (3 - P.mt) * TC_SLAB;       // 128*(3-mt) database rows x 64 B
    unsigned char *sA = smem;
    unsigned char *sB = sA + (size_t)P.chunks * a_chunk_bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sB + (size_t)P.stages * P.group * b_stage_bytes);
    uint64_t *full = bars;                          // [stages]
    uint64_t *empty = bars + TC_MAX_STAGES;         // [stages]
    uint64_t *a_full = bars + 2 * TC_MAX_STAGES;
    uint64_t *a_empty = a_full + 1;
    uint64_t *t_full = a_full + 2;                  // [2]
    uint64_t *t_empty = a_full + 4;                 // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(a_full + 6);
    volatile int *sweep_start = reinterpret_cast<volatile int *>(a_full + 7);     // [8], by local query-block index
    uint64_t *h_full = a_full + 12;                 // [2 buffers][2 halves]   (split protocol)
    uint64_t *h_empty = a_full + 16;                // [2 buffers][2 halves]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q_per_block = TC_TILE * P.mt;
    const int x_per_tile = TC_TILE * (3 - P.mt);
    const int n_tiles = (int)((P.n_total + x_per_tile - 1) / x_per_tile);

    if (threadIdx.x == 0) {
        for (int s = 0; s < P.stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        for (int b = 0; b < 2; ++b) { mbar_init(t_full + b, 1); mbar_init(t_empty + b, TC_EPI_WARPS); }
        for (int b = 0; b < 4; ++b) { mbar_init(h_full + b, 1); mbar_init(h_empty + b, TC_EPI_WARPS / 2); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer (whole warp runs the loop, one elected lane issues) =================
        int stage = 0;
        uint32_t phase = 0, aphase = 0;
        TcItem item;
        for (int ii = 0; tc_item(P, ii, n_tiles, item); ++ii) {
            const int64_t qb = item.qb;
            mbar_wait(a_empty, aphase ^ 1);
            const int qrow = (int)(P.q_begin + qb * q_per_block);
            if (elect_one()) {
                mbar_expect_tx(a_full, (uint32_t)(P.chunks * a_chunk_bytes));
                for (int c = 0; c < P.chunks; ++c)
                    tma_load_2d(sA + (size_t)c * a_chunk_bytes, &map_a, a_full, c * TC_KCH, qrow);
            }
            __syncwarp();
            aphase ^= 1;
            // Every sweep covers all tiles, so where it starts is free: CTA 0 publishes the tile it is loading
            // and the other CTAs start their sweeps there.  CTAs drift apart (a few per cent per sweep); without
            // this they end up streaming different parts of an operand array that is larger than L2 (measured:
            // 370 GB of DRAM reads at C3 for 14 GB of ideal traffic), with it the pack stays inside an L2 window.
            // (an item of the tail round covers its own part of the tiles, front to back)
            int s0 = (blockIdx.x == 0 || item.tail) ? item.tbeg : *reinterpret_cast<volatile int *>(P.cursor);
            if (s0 < 0 || s0 >= n_tiles) s0 = 0;
            s0 = __shfl_sync(AB_FULL, s0, 0);
            if (lane == 0) { sweep_start[ii & 7] = s0; __threadfence_block(); }
            __syncwarp();
            for (int ti = 0; ti < item.tcnt; ++ti) {
                int64_t t = s0 + ti;
                if (t >= n_tiles) t -= n_tiles;
                if (blockIdx.x == 0 && lane == 0 && (ti & 7) == 0 && !item.tail) *reinterpret_cast<volatile int *>(P.cursor) = (int)t;
                // slabs of this tile: one (mt = 2) or two (mt = 1) blocks of `chunks` x 8 KB
                const unsigned char *slab0 = P.bop + (size_t)t * (3 - P.mt) * P.chunks * TC_SLAB;
                // one ring slot = `group` consecutive K chunks of the tile = ONE contiguous bulk copy and ONE
                // barrier round (measured: the per-copy / per-barrier issue overhead of 8 KB stages, not bytes
                // or MMAs, set the floor of this kernel)
                for (int c0 = 0; c0 < P.chunks; c0 += P.group) {
                    const int gsz = min(P.group, P.chunks - c0);
                    mbar_wait(empty + stage, phase ^ 1);
                    if (elect_one()) {
                        const uint32_t bytes = (uint32_t)gsz * TC_SLAB;
                        unsigned char *dst = sB + (size_t)stage * P.group * b_stage_bytes;
                        mbar_expect_tx(full + stage, bytes * (uint32_t)(3 - P.mt));
                        bulk_load(dst, slab0 + (size_t)c0 * TC_SLAB, bytes, full + stage);
                        if (P.mt == 1)                              // second 128-point slab of the 256-point tile
                            bulk_load(dst + (size_t)P.group * TC_SLAB, slab0 + (size_t)(P.chunks + c0) * TC_SLAB, bytes,
                                      full + stage);
                    }
                    __syncwarp();
                    if (++stage == P.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (warp-uniform control flow, one elected lane issues) =================
        // (Two issuer warps alternating tiles were tried: the barrier waits of one then overlap the MMAs of the
        // other, but two in-order consumers of one ring can alias mbarrier phases, and the full kernel is
        // epilogue-bound anyway; see profiles/r1_knn_notes.md.)
        // instruction descriptor: D=F32, A=B=F16, K-major both, N=128, M=128
        const uint32_t idesc = (1u << 4) | ((uint32_t)(TC_TILE >> 3) << 17) | ((uint32_t)(TC_TILE >> 4) << 24);
        int stage = 0;
        int64_t it = 0, nblocks_done = 0;
        uint32_t phase = 0, aphase = 0;
        // descriptors: only the 14-bit start-address field changes; every offset used here is a multiple of 16
        const uint64_t a_desc0 = umma_desc_sw64(smem_u32(sA)), b_desc0 = umma_desc_sw64(smem_u32(sB));
        const uint32_t a_step = (uint32_t)a_chunk_bytes >> 4;
        const uint32_t slot_step = (uint32_t)(P.group * b_stage_bytes) >> 4;      // ring slot
        const uint32_t a_half = P.mt == 2 ? (TC_SLAB >> 4) : 0;                   // second query tile
        const uint32_t b_half = P.mt == 1 ? ((uint32_t)(P.group * TC_SLAB) >> 4) : 0;   // second slab of a 256-point tile
        if (P.split) {
            // One ring slot per tile; all K steps of half 0 (query tile 0, or the first 128 points of a wide tile),
            // commit, then all K steps of half 1, commit.  Each half has its own ready/released barrier pair.
            TcItem item;
            for (int ii = 0; tc_item(P, ii, n_tiles, item); ++ii) {
                mbar_wait(a_full, aphase);
                aphase ^= 1;
                for (int t = 0; t < item.tcnt; ++t, ++it) {
                    const int buf = (int)(it & 1);
                    const uint32_t tmem_d = tmem_base + (uint32_t)buf * 256;
                    const uint32_t eph = (uint32_t)(((it >> 1) & 1) ^ 1);
                    if (lane == 0) mbar_wait(full + stage, phase);
                    else if (lane == 31) mbar_wait(h_empty + buf * 2, eph);
                    __syncwarp();
                    tc_fence_after();
                    const uint64_t b_slot = b_desc0 + (uint64_t)((uint32_t)stage * slot_step);
                    for (int h = 0; h < 2; ++h) {
                        if (h == 1) { mbar_wait(h_empty + buf * 2 + 1, eph); tc_fence_after(); }
                        if (elect_one()) {
                            const uint32_t d_h = tmem_d + (uint32_t)h * TC_TILE;
                            const uint64_t a_h = a_desc0 + (uint64_t)((uint32_t)h * a_half);
                            const uint64_t b_h = b_slot + (uint64_t)((uint32_t)h * b_half);
                            for (int l = 0; l < P.chunks; ++l) {
                                const uint64_t a_c = a_h + (uint64_t)((uint32_t)l * a_step);
                                const uint64_t b_s = b_h + (uint64_t)((uint32_t)l * (TC_SLAB >> 4));
                                const bool two = P.ksteps - 2 * l >= 2;
                                tc_mma_f16(d_h, a_c, b_s, idesc, (uint32_t)(l != 0));
                                if (two) tc_mma_f16(d_h, a_c + 2, b_s + 2, idesc, 1u);
                            }
                            if (h == 1) tc_commit(empty + stage);   // ring slot free once every MMA of the tile retires
                            tc_commit(h_full + buf * 2 + h);        // this half's accumulators ready for its four warps
                        }
                        __syncwarp();
                    }
                    if (++stage == P.stages) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) tc_commit(a_empty);            // A tiles may be overwritten
                __syncwarp();
                ++nblocks_done;
            }
        } else {
        TcItem item;
        for (int ii = 0; tc_item(P, ii, n_tiles, item); ++ii) {
            mbar_wait(a_full, aphase);
            aphase ^= 1;
            for (int t = 0; t < item.tcnt; ++t, ++it) {
                const int buf = (int)(it & 1);
                const uint32_t tmem_d = tmem_base + (uint32_t)buf * 256;
                for (int c0 = 0; c0 < P.chunks; c0 += P.group) {
                    const int gsz = min(P.group, P.chunks - c0);
                    // lane 0 waits for the operands, lane 31 for the accumulator buffer
                    if (lane == 0) mbar_wait(full + stage, phase);
                    else if (lane == 31 && c0 == 0) mbar_wait(t_empty + buf, (uint32_t)(((it >> 1) & 1) ^ 1));
                    __syncwarp();
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t b_slot = b_desc0 + (uint64_t)((uint32_t)stage * slot_step);
                        for (int l = 0; l < gsz; ++l) {
                            const int c = c0 + l;
                            const uint64_t a_c = a_desc0 + (uint64_t)((uint32_t)c * a_step);
                            const uint64_t b_s = b_slot + (uint64_t)((uint32_t)l * (TC_SLAB >> 4));
                            const bool two = P.ksteps - 2 * c >= 2;
                            tc_mma_f16(tmem_d, a_c, b_s, idesc, (uint32_t)(c != 0));
                            if (two) tc_mma_f16(tmem_d, a_c + 2, b_s + 2, idesc, 1u);
                            tc_mma_f16(tmem_d + TC_TILE, a_c + a_half, b_s + b_half, idesc, (uint32_t)(c != 0));
                            if (two) tc_mma_f16(tmem_d + TC_TILE, a_c + a_half + 2, b_s + b_half + 2, idesc, 1u);
                        }
                        tc_commit(empty + stage);                   // ring slot free once these MMAs retire
                        if (c0 + gsz == P.chunks) tc_commit(t_full + buf);   // accumulators ready for the epilogue
                    }
                    __syncwarp();
                    if (++stage == P.stages) { stage = 0; phase ^= 1; }
                }
            }
            if (elect_one()) tc_commit(a_empty);                // A tiles may be overwritten
            __syncwarp();
            ++nblocks_done;
        }
        }
        // do not let the CTA retire with a tcgen05.commit arrive still in flight
        if (nblocks_done > 0) mbar_wait(a_empty, (uint32_t)((nblocks_done - 1) & 1));
    } else {
        // ================= epilogue: thread <-> query row; warps 2-5 half 0, warps 6-9 half 1 =================
        const int ew = warp - 2;                                // 0..7
        const int half = ew >> 2;                               // accumulator columns [half*128, half*128+128)
        const int lane_base = (warp & 3) * 32;                  // TMEM lanes this warp may touch
        const int row = lane_base + lane;
        const int etid = ew * 32 + lane;
        const int list_id = P.mt == 1 ? half : 0;
        unsigned long long *bufp = P.bufg + (size_t)blockIdx.x * TC_BUF * TC_EPI_THREADS + etid;   // entry e at bufp[e * 256]
        const uint32_t taddr0 = tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)half * TC_TILE;
        int64_t it = 0;                                         // tiles of this CTA so far: accumulator buffer = it & 1
        TcItem item;
        for (int ii = 0; tc_item(P, ii, n_tiles, item); ++ii) {
            const long long qg = item.qb * q_per_block + (P.mt == 2 ? half * TC_TILE : 0) + row;    // query (row of the range)
            const bool live = qg < P.nq;
            // candidate list of this thread: the query's ordinary list, or - tail round - the list of this item's part
            unsigned long long *cg;
            long long q;
            if (item.tail) {
                const long long tail_rows = (long long)P.R * q_per_block;
                cg = P.cand_tail + (size_t)(item.part * (3 - P.mt) + list_id) * tail_rows * P.m;
                q = qg - (long long)P.full * gridDim.x * q_per_block;
            } else {
                cg = P.cand + (size_t)list_id * P.nq * P.m;
                q = qg;
            }
            float tau = INFINITY;
            int cnt = 0;
            int sweep0 = 0;
            for (int ti = 0; ti < item.tcnt; ++ti, ++it) {
                const int buf = (int)(it & 1);
                mbar_wait(P.split ? h_full + buf * 2 + half : t_full + buf, (uint32_t)((it >> 1) & 1));   // (one poller per warp + __syncwarp measured slower)
                tc_fence_after();
                if (ti == 0) sweep0 = sweep_start[ii & 7];          // published by the producer before this item's first load
                int64_t t = sweep0 + ti;
                if (t >= n_tiles) t -= n_tiles;
                const int64_t col0 = t * x_per_tile + (P.mt == 1 ? half * TC_TILE : 0);
                const uint32_t taddr = taddr0 + (uint32_t)buf * 256;
                uint32_t ra[32];
#pragma unroll 1
                for (int cc = 0; cc < TC_TILE / 32; ++cc) {
                    // one unit per load/wait: a single instantiation of the scan keeps the hot loop small enough for
                    // the instruction cache (two inlined copies cost 19 % of the cycles in no-instruction stalls);
                    // the other epilogue warp of the scheduler covers the tcgen05.ld latency
                    tc_ld32(taddr + cc * 32, ra);
                    tc_wait_ld();
                    tc_touch32(ra);
                    epi_unit(ra, (int)col0 + cc * 32, live, P, cg, q, cnt, tau, bufp, lane);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(P.split ? h_empty + buf * 2 + half : t_empty + buf);
                // prune AFTER handing the accumulator back: the merge (two L2 round trips + a bitonic network)
                // then overlaps the MMAs of the next tiles instead of sitting on the MMA->epilogue handoff
                const unsigned due = __ballot_sync(AB_FULL, cnt > TC_BUF - 8);
                if (due) warp_prune(due, P, cg, q, cnt, tau, bufp, lane);
            }
            warp_prune(__ballot_sync(AB_FULL, cnt > 0), P, cg, q, cnt, tau, bufp, lane);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---------------------------------------------------------------------------------------
// Tail round: fold the S part lists of a query into its ordinary list (m smallest keys of their union, sorted).
// Valid for the certificate: a point outside every part list scores >= that part's m-th key >= the m-th key of the
// union.  One warp per (tail query, list); every key finds its rank by binary searches in the other (sorted) parts.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
knn_tail_merge_kernel(const unsigned long long *__restrict__ cand_tail, int S, int lists, int64_t tail_rows, int m,
                      int64_t q0, int64_t nq, unsigned long long *__restrict__ cand) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (w >= tail_rows * lists) return;
    const int64_t r = w / lists;
    const int l = (int)(w - r * lists);
    const int64_t q = q0 + r;
    if (q >= nq) return;
    for (int e = lane; e < S * m; e += 32) {
        const int sp = e / m, j = e - sp * m;
        const unsigned long long key = cand_tail[((size_t)(sp * lists + l) * tail_rows + r) * m + j];
        int rank = j;
        for (int o = 0; o < S; ++o) {
            if (o == sp) continue;
            const unsigned long long *L = cand_tail + ((size_t)(o * lists + l) * tail_rows + r) * m;
            int lo = 0, hi = m;                        // keys of part o that sort before `key` (ties: the lower part first)
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                const unsigned long long v = L[mid];
                if (v < key || (v == key && o < sp)) lo = mid + 1; else hi = mid;
            }
            rank += lo;
        }
        if (rank < m) cand[((size_t)l * nq + q) * m + rank] = key;
    }
}

// ---------------------------------------------------------------------------------------
// exact re-rank + certificate, one warp per query
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ bool pair_lt(double da, int ia, double db, int ib) { return da < db || (da == db && ia < ib); }

__global__ void __launch_bounds__(256)
knn_rerank_kernel(const double *__restrict__ emb, int64_t n_total, int d, int64_t q_begin, int64_t nq, int k, int m,
                  int lists, const unsigned long long *__restrict__ cand, const double *__restrict__ qn,
                  const unsigned int *__restrict__ rmax_bits, int32_t *__restrict__ idx, double *__restrict__ rdist,
                  int64_t *__restrict__ fb_list, unsigned long long *__restrict__ stats, unsigned int *__restrict__ relerr_bits,
                  double kappa, const double *__restrict__ qh2, const double *__restrict__ qdel) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    const int mg = lists * m;                              // candidate groups of the query (all lists)
    const int mp = mg * TC_GRP;                            // candidate points; power of two
    double *sd = reinterpret_cast<double *>(sm_raw) + (size_t)warp * mp;
    int *si = reinterpret_cast<int *>(sm_raw + (size_t)nwarp * mp * sizeof(double)) + (size_t)warp * mp;
    const float R = __uint_as_float(*rmax_bits);
    const double sc = knn_scale_of(R);
    const double inv_sc2 = 1.0 / (sc * sc);                // scores are in scaled units
    int *sel = reinterpret_cast<int *>(sm_raw + (size_t)nwarp * mp * (sizeof(double) + sizeof(int))) + (size_t)warp * (2 * mg + 1);
    int *goff = sel + mg;                                  // [mg + 1]
    for (int64_t q = (int64_t)blockIdx.x * nwarp + warp; q < nq; q += (int64_t)gridDim.x * nwarp) {
        const int64_t qid = q_begin + q;
        const double *qv = emb + qid * d;
        const double qnorm2 = qn[qid];
        // The query the GEMM saw is q_hi, exactly.  A score plus |q_hi|^2 is the squared distance to q_hi within eps,
        // and | |x - q| - |x - q_hi| | <= del = |q - q_hi| for every x.
        const double qbase = qh2[qid];
        const double del = qdel[qid];
        const double scale = sqrt(qnorm2) * 1.0009765625 + (double)R;      // |q_hi| <= (1 + 2^-11) |q|
        const double eps = kappa * scale * scale;
        // Only groups that can hold one of the k nearest are expanded.  With the lists sorted by key, the k-th
        // smallest approximate group minimum a_k of any one list bounds D_k <= a_k + eps (k distinct groups, each
        // contributing its minimum); a group whose approximate minimum exceeds a_k + 2 eps has every point
        // strictly farther than D_k.
        double thr = INFINITY;
        if (lane < lists) {
            const unsigned long long kk = cand[((size_t)lane * nq + q) * m + (k - 1 < m ? k - 1 : m - 1)];
            if (kk != TC_KEY_MAX && k <= m) {
                const double ak = (double)key_score(kk) * inv_sc2 + qbase;
                if (del > 0.0) {
                    // D_k <= (sqrt(a_k + eps) + del)^2; a group with approximate minimum a has every point at
                    // distance >= sqrt(a - eps) - del from q
                    const double t = sqrt(fmax(ak + eps, 0.0)) + 2.0 * del;
                    thr = t * t * (1.0 + 1e-14) + eps;
                } else thr = ak + 2.0 * eps;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) thr = fmin(thr, ab_shfl_xor_f64(thr, o));
        int n_exp = 0;
        for (int g0 = 0; g0 < mg; g0 += 32) {
            const int gi = g0 + lane;
            bool take = false;
            if (gi < mg) {
                const int g = gi / m, e = gi - g * m;
                const unsigned long long key = cand[((size_t)g * nq + q) * m + e];
                take = key != TC_KEY_MAX && ((double)key_score(key) * inv_sc2 + qbase <= thr);
            }
            const unsigned bal = __ballot_sync(AB_FULL, take);
            if (take) sel[n_exp + __popc(bal & ((1u << lane) - 1))] = gi;
            n_exp += __popc(bal);
        }
        __syncwarp();
        // points to rank exactly: of every selected group the pairs its key flags (scores below the threshold of the time
        // the group was listed).  goff[i] = first position of group sel[i] in the point list.
        int np = 0;
        for (int i0 = 0; i0 < n_exp; i0 += 32) {
            const int i = i0 + lane;
            unsigned long long key = 0ull;
            int cntp = 0;
            if (i < n_exp) {
                const int gi = sel[i], g = gi / m, e = gi - g * m;
                key = cand[((size_t)g * nq + q) * m + e];
                cntp = 2 * __popc(tc_key_pairs(key));
            }
            int inc = cntp;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int up = __shfl_up_sync(AB_FULL, inc, o);
                if (lane >= o) inc += up;
            }
            const int off = np + inc - cntp;
            if (i < n_exp) {
                goff[i] = off;
                const int first = tc_key_first(key);
                unsigned pm = tc_key_pairs(key);
                int w = off;
                while (pm) {
                    const int pp = __ffs(pm) - 1;
                    pm &= pm - 1;
                    si[w++] = first + 2 * pp;
                    si[w++] = first + 2 * pp + 1;
                }
            }
            np += __shfl_sync(AB_FULL, inc, 31);
        }
        if (lane == 0) goff[n_exp] = np;
        __syncwarp();
        int p2 = 32;
        while (p2 < np || p2 < k + 1) p2 <<= 1;            // sort size: power of two covering the expanded points
        if (p2 > mp) p2 = mp;
        for (int c = lane; c < p2; c += 32) {
            double dist = INFINITY;
            int id = INT_MAX;
            if (c < np) {
                const long long cnd = si[c];
                if (cnd < n_total) {                          // (padding columns: ids past the end)
                    const double *xv = emb + cnd * d;
                    double acc = 0.0;
                    if ((d & 1) == 0) {
                        // same sequential, non-fused sum; 16-byte loads: every lane walks its own row, so the L1 sees
                        // one request per lane and load - half as many as with 8-byte loads (it was the re-rank's bound)
                        const double2 *xv2 = reinterpret_cast<const double2 *>(xv);
                        const double2 *qv2 = reinterpret_cast<const double2 *>(qv);
                        for (int j = 0; j < d / 2; ++j) {
                            const double2 x = __ldg(xv2 + j), qq = qv2[j];
                            const double d0 = __dsub_rn(qq.x, x.x);
                            acc = __dadd_rn(acc, __dmul_rn(d0, d0));
                            const double d1 = __dsub_rn(qq.y, x.y);
                            acc = __dadd_rn(acc, __dmul_rn(d1, d1));
                        }
                    } else {
                        for (int j = 0; j < d; ++j) {         // the reference's arithmetic, no FMA
                            const double df = __dsub_rn(qv[j], xv[j]);
                            acc = __dadd_rn(acc, __dmul_rn(df, df));
                        }
                    }
                    dist = acc;
                    id = (int)cnd;
                }
            }
            sd[c] = dist;
            si[c] = id;
        }
        __syncwarp();
        // guard of the error model: the key's score approximates the SMALLEST distance of the group, and the point that
        // attains that score is in a flagged pair
        double worst_rel = 0.0;
        for (int i = lane; i < n_exp; i += 32) {
            const int gi = sel[i], g = gi / m, e = gi - g * m;
            const unsigned long long key = cand[((size_t)g * nq + q) * m + e];
            double gmin = INFINITY;
            for (int c = goff[i]; c < goff[i + 1]; ++c) gmin = fmin(gmin, sd[c]);
            if (gmin < INFINITY) {
                const double approx = (double)key_score(key) * inv_sc2 + qbase;
                double err = fabs(approx - gmin);
                if (del > 0.0) {                              // what lies outside the band the perturbed query allows
                    const double sg = sqrt(gmin), up = (sg + del) * (sg + del), lo = sg > del ? (sg - del) * (sg - del) : 0.0;
                    err = fmax(0.0, fmax(approx - up, lo - approx));
                }
                worst_rel = fmax(worst_rel, err / (scale * scale));
            }
        }
        __syncwarp();
        for (int size = 2; size <= p2; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int i = lane; i < (p2 >> 1); i += 32) {
                    const int lo = ((i & ~(stride - 1)) << 1) | (i & (stride - 1));
                    const int hi = lo | stride;
                    const bool up = ((lo & size) == 0);
                    const double dl = sd[lo], dh = sd[hi];
                    const int il = si[lo], ih = si[hi];
                    const bool sw = up ? pair_lt(dh, ih, dl, il) : pair_lt(dl, il, dh, ih);
                    if (sw) { sd[lo] = dh; sd[hi] = dl; si[lo] = ih; si[hi] = il; }
                }
                __syncwarp();
            }
        }
        for (int c = lane; c < k; c += 32) {
            idx[q * k + c] = si[c];
            rdist[q * k + c] = sd[c];
        }
        if (lane == 0) {
            const double dk = sd[k - 1];
            // a list that is not full holds every group it saw -> no outside point there
            double tau = INFINITY;
            for (int g = 0; g < lists; ++g) {
                const unsigned long long last = cand[((size_t)g * nq + q) * m + (m - 1)];
                if (last != TC_KEY_MAX) tau = fmin(tau, (double)key_score(last) * inv_sc2);
            }
            // every point outside the listed groups has  D >= tau + ||q||^2 - eps ; strict '>' also settles ties
            const double t2 = tau + qbase - eps;
            bool certified = t2 > dk;
            if (del > 0.0 && certified) {                      // |x - q| >= |x - q_hi| - del > sqrt(t2) - del
                const double t = sqrt(t2) - del;
                certified = t > 0.0 && t * t * (1.0 - 1e-14) > dk;
            }
            if (certified) {
                atomicAdd(stats + 3, 1ull);
                if (k < p2 && sd[k] == dk) atomicAdd(stats + 0, 1ull);
                if (si[0] != (int)qid || (k > 1 && sd[1] == 0.0)) atomicAdd(stats + 2, 1ull);
            } else {
                const unsigned long long slot = atomicAdd(stats + 1, 1ull);
                fb_list[slot] = qid;
                fb_list[nq + slot] = q;                              // output row
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) worst_rel = fmax(worst_rel, ab_shfl_xor_f64(worst_rel, o));
        if (lane == 0) atomicMax(relerr_bits, __float_as_uint((float)worst_rel));
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
static int tc_candidates(int k) {
    const int want = k + std::max(8, k / 2);
    int m = 32;
    while (m < want) m <<= 1;                      // power of two (bitonic re-rank), 32..256
    return std::min(m, 256);
}
static int64_t tc_pad(int64_t n) { return (n + 255) / 256 * 256; }
// Schedule of `blocks` query blocks on `sms` persistent CTAs (TcParams): `full` whole rounds, then R tail blocks in S
// parts each, dealt round-robin.  Cost in sweeps = full + ceil(R S / grid) / S; the tail may take over the last whole
// round (R up to 2 sms - 1) when that balances better (1,276 blocks on 148 SMs: 8.62 rounds of work cost 9 as whole
// rounds, 8.67 with the last two rounds cut in thirds).
static void tc_schedule(int64_t blocks, int sms, int smax, int *full, int *R, int *S) {
    double best = 1e300;
    *full = 0; *R = 0; *S = 1;
    for (int pull = 0; pull < 2; ++pull) {
        const int64_t f = blocks / sms - pull;
        if (f < 0) break;
        const int64_t r = blocks - f * sms;
        for (int s = 1; s <= std::max(1, smax); ++s) {
            const int64_t grid = f > 0 ? sms : std::min<int64_t>(sms, std::max<int64_t>(1, r * s));
            const double cost = (double)f + (r > 0 ? (double)((r * s + grid - 1) / grid) / s : 0.0) + 1e-3 * s + 1e-3 * pull;
            if (cost < best - 1e-9) { best = cost; *full = (int)f; *R = (int)r; *S = s; }
        }
    }
}
static int tc_kt(int d) { return (TC_NPROD * d + 3 + TC_KCH - 1) / TC_KCH * TC_KCH; }
// two resident query tiles when their A operand (256 rows x kt halves) leaves room for the B ring
// (at d = 100 the two tiles take 112 KB and leave three ring slots of 32 KB: measured on 200,000 cells x 100 dims, k = 30:
// GEMM 25.4 -> 21.5 ms and, with one candidate list per query instead of two, re-rank 58 -> 32 ms)
static int tc_mt(int d) { return tc_kt(d) * 2 * 256 <= 112 * 1024 ? 2 : 1; }

struct TcWs {
    double *mean, *partial, *qn, *qh2, *qdel;
    __half *aop, *bop;
    unsigned long long *cand, *cand_tail;
    int64_t *fb_list;
    unsigned int *scal;      // [0] rmax bits, [1] relerr bits
    unsigned long long *bufg; // append buffers of the main kernel: [<=256 CTAs][TC_BUF][256]
    void *exact_ws;
    size_t exact_bytes;
};

static size_t tc_layout(int64_t n, int d, int64_t nq, int k, void *ws, size_t bytes, TcWs *out) {
    ab_arena a(ws ? ws : (void *)1024, ws ? bytes : (size_t)-1 / 2);
    TcWs w;
    const int kt = tc_kt(d), m = tc_candidates(k);
    w.aop = a.take<__half>((size_t)n * kt);
    w.bop = a.take<__half>((size_t)tc_pad(n) * kt);
    w.mean = a.take<double>(d);
    w.partial = a.take<double>((size_t)1024 * d);
    w.qn = a.take<double>(n);
    w.qh2 = a.take<double>(n);
    w.qdel = a.take<double>(n);
    w.cand = a.take<unsigned long long>((size_t)2 * nq * m);
    // part lists of the tail phase (TcParams): at most 4 lists for at most two grids of 256-query blocks
    w.cand_tail = a.take<unsigned long long>((size_t)4 * std::min<int64_t>((nq + 255) / 256 * 256, 2 * 256 * 256) * m);
    w.fb_list = a.take<int64_t>((size_t)2 * nq);
    w.scal = a.take<unsigned int>(4);
    w.bufg = a.take<unsigned long long>((size_t)256 * TC_BUF * TC_EPI_THREADS);
    // the exact scan only ever sees the uncertified minority; reserve room for 1/8 of the queries
    w.exact_bytes = ab_knn_exact_ws_bytes(std::max<int64_t>(1024, nq / 8), k, 64);
    w.exact_ws = a.take<char>(w.exact_bytes);
    if (out) *out = w;
    return a.ok ? a.off : 0;
}

// host-only diagnostic (include/alona_b200.h): the schedule ab_knn's tensor pass uses for `blocks` query blocks
extern "C" int ab_knn_schedule(int64_t blocks, int32_t n_sm, int32_t max_parts, int32_t *out3) {
    if (blocks < 0 || n_sm < 1 || max_parts < 1 || !out3) return ab_fail("ab_knn_schedule: bad argument");
    int full = 0, R = 0, S = 1;
    tc_schedule(blocks, std::min(n_sm, 256), max_parts, &full, &R, &S);
    out3[0] = full; out3[1] = R; out3[2] = S;
    return 0;
}

size_t ab_knn_tc_ws_bytes(int64_t n_total, int d, int64_t nq, int k) {
    return tc_layout(n_total, d, nq, k, nullptr, 0, nullptr) + 2048;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(CUtensorMap *map, __half *base, int64_t rows, int kt, int box_rows) {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e != cudaSuccess || !p) return ab_fail("cuTensorMapEncodeTiled entry point unavailable: %s", cudaGetErrorString(e));
        fn = (PFN_encodeTiled)p;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)kt, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)kt * sizeof(__half)};
    cuuint32_t box[2] = {(cuuint32_t)TC_KCH, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return ab_fail("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}

int ab_knn_tc_run(ab_ctx *ctx, const double *emb, int64_t n_total, int d, int64_t q_begin, int64_t nq, int k,
                  int32_t *idx, double *rdist, unsigned long long *stats, void *ws, size_t ws_bytes, cudaStream_t st) {
    const int m = tc_candidates(k);
    const int kt = tc_kt(d), chunks = kt / TC_KCH;
    if (chunks > TC_MAX_CHUNKS || n_total < m || n_total >= ((int64_t)1 << 31) - 512) {
        // outside the tensor path's envelope (very wide embeddings, tiny inputs): exact scan
        return ab_knn_exact_run(ctx, emb, n_total, d, nullptr, q_begin, nq, k, ab_knn_exact_pick_splits(ctx, nq), nullptr,
                                idx, rdist, stats, ws, ws_bytes, st);
    }
    TcWs w;
    AB_REQUIRE(tc_layout(n_total, d, nq, k, ws, ws_bytes, &w) != 0, "ab_knn: workspace too small");
    AB_CUDA(cudaMemsetAsync(w.scal, 0, 4 * sizeof(unsigned int), st));
    const int mt = tc_mt(d), lists = 3 - mt;
    // ---- prep ----
    ab_prof_begin(ctx, AB_PROF_KNN_PREP, st);
    const int parts = (int)std::min<int64_t>(1024, std::max<int64_t>(1, n_total / 64));
    knn_colsum_kernel<<<parts, 256, 0, st>>>(emb, n_total, d, w.partial);
    AB_LAUNCH_CHECK(ctx);
    knn_colmean_kernel<<<(d + 127) / 128, 128, 0, st>>>(w.partial, parts, d, n_total, w.mean);
    AB_LAUNCH_CHECK(ctx);
    knn_rmax_kernel<<<ab_wave_grid(ctx, knn_rmax_kernel, 256, 0, (n_total + 7) / 8), 256, 0, st>>>(emb, n_total, d, w.mean, w.scal);
    AB_LAUNCH_CHECK(ctx);
    knn_prep_kernel<<<ab_wave_grid(ctx, knn_prep_kernel, 256, 0, (n_total + 7) / 8), 256, 0, st>>>(emb, n_total, tc_pad(n_total), d, kt, w.mean, w.scal, w.aop, w.bop, w.qn,
                                            w.qh2, w.qdel);
    AB_LAUNCH_CHECK(ctx);
    AB_CUDA(cudaMemsetAsync(w.cand, 0xFF, (size_t)lists * nq * m * sizeof(unsigned long long), st));
    ab_prof_end(ctx, AB_PROF_KNN_PREP, st);
    // ---- main ----
    CUtensorMap map_a;
    if (make_map(&map_a, w.aop, n_total, kt, TC_TILE * mt)) return 1;
    const size_t a_bytes = (size_t)chunks * mt * TC_SLAB;
    const size_t b_stage = (size_t)(3 - mt) * TC_SLAB;
    const size_t bar_bytes = 512;
    const size_t avail = (size_t)ctx->smem_optin - 1024 - a_bytes - bar_bytes;
    // ring slot = `group` K chunks (one bulk copy, one barrier round): the largest group that still leaves
    // three slots in flight
    int group = 0, stages = 0;
    for (int parts = 1; parts <= chunks; ++parts) {
        const int g = (chunks + parts - 1) / parts;
        const int r = (int)(avail / ((size_t)g * b_stage));
        if (r >= 3 || g == 1) { group = g; stages = std::min(r, TC_MAX_STAGES); break; }
    }
    AB_REQUIRE(stages >= 2, "ab_knn: not enough shared memory for d=%d k=%d", d, k);
    const size_t smem = 1024 + a_bytes + (size_t)stages * group * b_stage + bar_bytes;
    AB_CUDA(cudaFuncSetAttribute(knn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TcParams P{};
    P.n_total = n_total; P.q_begin = q_begin; P.nq = nq; P.chunks = chunks; P.ksteps = (TC_NPROD * d + 3 + 15) / 16;
    P.stages = stages; P.group = group; P.bufg = w.bufg;
    // a tile that is one ring slot: per-half accumulator hand-off; otherwise (wide embeddings) one hand-off per tile
    P.split = group == chunks ? 1 : 0;
    P.cursor = (int *)(w.scal + 2); P.mt = mt; P.m = m; P.cand = w.cand; P.bop = (const unsigned char *)w.bop;
    const int64_t n_qblocks = (nq + TC_TILE * mt - 1) / (TC_TILE * mt);
    const int64_t n_tiles = (n_total + TC_TILE * (3 - mt) - 1) / (TC_TILE * (3 - mt));
    // work items: whole sweeps, then a tail round whose R blocks are swept by S CTAs each (see TcParams)
    const int sms = std::min(ctx->sm_count, 256);
    tc_schedule(n_qblocks, sms, (int)std::min<int64_t>(4 / lists, n_tiles), &P.full, &P.R, &P.S);
    const unsigned grid = (unsigned)(P.full > 0 ? sms : std::min<int64_t>(sms, std::max<int64_t>(1, (int64_t)P.R * P.S)));
    const int64_t tail_rows = (int64_t)P.R * TC_TILE * mt;
    P.cand_tail = w.cand_tail;
    if (P.S > 1)
        AB_CUDA(cudaMemsetAsync(w.cand_tail, 0xFF, (size_t)P.S * lists * tail_rows * m * sizeof(unsigned long long), st));
    ab_prof_begin(ctx, AB_PROF_KNN_MAIN, st);
    knn_tc_kernel<<<grid, TC_NTHREADS, smem, st>>>(map_a, P);
    AB_LAUNCH_CHECK(ctx);
    if (P.S > 1) {
        knn_tail_merge_kernel<<<(unsigned)((tail_rows * lists + 7) / 8), 256, 0, st>>>(
            w.cand_tail, P.S, lists, tail_rows, m, (int64_t)P.full * grid * TC_TILE * mt, nq, w.cand);
        AB_LAUNCH_CHECK(ctx);
    }
    ab_prof_end(ctx, AB_PROF_KNN_MAIN, st);
    // ---- exact re-rank + certificate ----
    ab_prof_begin(ctx, AB_PROF_KNN_RERANK, st);
    const int mp = lists * m * TC_GRP;                              // candidate points per query (m groups per list)
    int rr_warps = 8;
    while (rr_warps > 1 && (size_t)rr_warps * mp * (sizeof(double) + sizeof(int)) > (size_t)160 * 1024) rr_warps >>= 1;
    const size_t rr_smem = (size_t)rr_warps * mp * (sizeof(double) + sizeof(int)) + (size_t)rr_warps * (2 * (mp / TC_GRP) + 1) * sizeof(int);
    AB_CUDA(cudaFuncSetAttribute(knn_rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rr_smem));
    const unsigned g_rr = ab_wave_grid(ctx, knn_rerank_kernel, rr_warps * 32, rr_smem, (nq + rr_warps - 1) / rr_warps);
    knn_rerank_kernel<<<g_rr, rr_warps * 32, rr_smem, st>>>(emb, n_total, d, q_begin, nq, k, m, lists, w.cand, w.qn, w.scal,
                                                           idx, rdist, w.fb_list, stats, w.scal + 1, tc_kappa(kt), w.qh2,
                                                           w.qdel);
    AB_LAUNCH_CHECK(ctx);
    ab_prof_end(ctx, AB_PROF_KNN_RERANK, st);
    // ---- uncertified queries: exact scan ----
    unsigned long long h_stats[4];
    unsigned int h_scal[2];
    AB_CUDA(cudaMemcpyAsync(h_stats, stats, sizeof(h_stats), cudaMemcpyDeviceToHost, st));
    AB_CUDA(cudaMemcpyAsync(h_scal, w.scal, sizeof(h_scal), cudaMemcpyDeviceToHost, st));
    AB_CUDA(cudaStreamSynchronize(st));
    const int64_t n_fb = (int64_t)h_stats[1];
    {
        // guard of the certificate's error model: the largest error seen on the candidates (hundreds per query) must
        // stay far inside kappa; if it ever does not, nothing of this pass is trusted and every query is recomputed
        // by the exact FP64 scan
        float seen;
        memcpy(&seen, &h_scal[1], sizeof(float));
        if ((double)seen > 0.5 * tc_kappa(kt)) {
            fprintf(stderr, "alona_b200: kNN candidate scores off by %.3g of (|q|+R)^2 (model bound %.3g): exact scan for all queries\n",
                    (double)seen, tc_kappa(kt));
            AB_CUDA(cudaMemsetAsync(stats, 0, 8 * sizeof(unsigned long long), st));
            return ab_knn_exact_run(ctx, emb, n_total, d, nullptr, q_begin, nq, k, ab_knn_exact_pick_splits(ctx, nq), nullptr,
                                    idx, rdist, stats, ws, ws_bytes, st);
        }
    }
    // stats[4] = max observed |approx - exact| / (||q||+R)^2 as float bits, stats[5] = candidates per list,
    // stats[6] = query tiles per CTA (2: 256-query blocks, 1: wide embeddings)
    unsigned long long extra[3] = {(unsigned long long)h_scal[1], (unsigned long long)m, (unsigned long long)mt};
    AB_CUDA(cudaMemcpyAsync(stats + 4, extra, sizeof(extra), cudaMemcpyHostToDevice, st));
    if (n_fb > 0) {
        int64_t done = 0;
        const int64_t batch = std::max<int64_t>(1024, nq / 8);
        while (done < n_fb) {
            const int64_t cur = std::min(batch, n_fb - done);
            const int splits = ab_knn_exact_pick_splits(ctx, cur);
            if (ab_knn_exact_run(ctx, emb, n_total, d, w.fb_list + done, 0, cur, k, splits, w.fb_list + nq + done, idx,
                                 rdist, stats, w.exact_ws, w.exact_bytes, st))
                return 1;
            done += cur;
        }
    }
    return 0;
}
