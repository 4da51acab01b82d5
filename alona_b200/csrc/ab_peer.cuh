// One-shot all-reduce of small FP64 vectors over NVLink peer memory, fused into the tails / heads of the IRLBA
// kernels ("all-reduce ... fused into the kernels' tails").
//
// Every rank owns one exchange buffer (cudaMalloc + cudaIpc handle, mapped by all peers: ab_peer_alloc /
// ab_peer_open).  A reduction instance is identified by (channel, seq): the producer kernel's tail PUSHES this
// rank's partial vector into slot [seq & 1][rank] of every rank's buffer (plain stores over NVLink), fences at
// system scope and then stores flag[channel][rank] = seq on every rank (st.release.sys).  The consumer kernel's head
// spins on its LOCAL flags (ld.acquire.sys) until all `world` senders show seq, then sums the slots in rank order
// 0..world-1: every rank forms the same sum in the same order, so the result is bitwise identical on all ranks and
// independent of arrival order.  Slots are double-buffered by seq parity: a sender can only reach instance seq+2
// after every rank has consumed instance seq (it must first consume seq+1, which needs every rank's push of seq+1,
// issued after that rank consumed seq).  Waits are bounded (about 10 s) and trap instead of hanging the GPU.
// At world == 1 nothing here runs: kernels read and write the local scalars / vectors directly.
#pragma once
#include "ab_async.cuh"

static constexpr int AB_PEER_MAX = 8;                  // ranks on one NVSwitch box
static constexpr int AB_PEER_VEC = 32768;              // channel 0: up to 32768 doubles (the A.v result, H)
static constexpr int AB_PEER_COEF = 128;               // channel 1: CGS coefficients (<= work dimension)
static constexpr int AB_PEER_SCAL = 8;                 // channel 2: ||F||^2
enum { AB_CH_VEC = 0, AB_CH_COEF = 1, AB_CH_SCAL = 2, AB_CH_COUNT = 3 };

struct AbPeer {                                        // lives in device memory (ab_ctx::d_peer)
    int world, rank;
    char *buf[AB_PEER_MAX];                            // every rank's exchange buffer; buf[rank] is the local one
};

__host__ __device__ constexpr size_t ab_peer_ch_len(int ch) {
    return ch == AB_CH_VEC ? AB_PEER_VEC : ch == AB_CH_COEF ? AB_PEER_COEF : AB_PEER_SCAL;
}
__host__ __device__ constexpr size_t ab_peer_ch_off(int ch) {      // byte offset of a channel's data area
    return 256 * AB_CH_COUNT +
           (ch > 0 ? 2 * AB_PEER_MAX * (size_t)AB_PEER_VEC * 8 : 0) + (ch > 1 ? 2 * AB_PEER_MAX * (size_t)AB_PEER_COEF * 8 : 0);
}
static constexpr size_t AB_PEER_BYTES = ab_peer_ch_off(AB_CH_SCAL) + 2 * AB_PEER_MAX * (size_t)AB_PEER_SCAL * 8;

#ifdef __CUDACC__
// kernels take the table by pointer (a device copy kept in the context); NULL = single rank / exchange off
__device__ __forceinline__ int ab_peer_world(const AbPeer *p) { return p ? p->world : 1; }
__device__ __forceinline__ uint32_t *ab_peer_flag(char *buf, int ch, int sender) {
    return reinterpret_cast<uint32_t *>(buf + 256 * ch) + sender * 8;          // 32 bytes apart
}
__device__ __forceinline__ double *ab_peer_slot(char *buf, int ch, uint32_t seq, int sender) {
    return reinterpret_cast<double *>(buf + ab_peer_ch_off(ch)) + ((size_t)(seq & 1) * AB_PEER_MAX + sender) * ab_peer_ch_len(ch);
}
// this rank's value i of instance (ch, seq) -> every rank's slot (the caller fences and publishes afterwards)
__device__ __forceinline__ void ab_peer_push(const AbPeer &p, int ch, uint32_t seq, int i, double v) {
    for (int q = 0; q < p.world; ++q) ab_peer_slot(p.buf[q], ch, seq, p.rank)[i] = v;
}
// after all pushes of this rank are ordered before the calling thread (fence / ticket): raise the flags
__device__ __forceinline__ void ab_peer_publish(const AbPeer &p, int ch, uint32_t seq) {
    // ONE system-scope fence, then relaxed flag stores: a st.release.sys per peer is a fence per peer, i.e. `world` NVLink
    // round trips in series in front of every exchange (777 exchanges per C3 pass)
    __threadfence_system();
    for (int q = 0; q < p.world; ++q) ab_st_relaxed_sys_u32(ab_peer_flag(p.buf[q], ch, p.rank), seq);
}
// block-wide wait for every sender's push of (ch, seq); ends with a __syncthreads
__device__ __forceinline__ void ab_peer_wait(const AbPeer &p, int ch, uint32_t seq) {
    if ((int)threadIdx.x < p.world) {
        const uint32_t *f = ab_peer_flag(p.buf[p.rank], ch, threadIdx.x);
        const long long t0 = clock64();
        while ((int32_t)(ab_ld_acquire_sys_u32(f) - seq) < 0) {
            if (clock64() - t0 > 20000000000ll) __trap();               // ~10 s: a rank died or diverged
            __nanosleep(20);
        }
    }
    __syncthreads();
}
// sum over ranks, fixed order, of entry i (reads bypass L1: the slots are written by peers)
__device__ __forceinline__ double ab_peer_sum(const AbPeer &p, int ch, uint32_t seq, int i) {
    double t = 0.0;
    for (int q = 0; q < p.world; ++q) t += __ldcg(ab_peer_slot(p.buf[p.rank], ch, seq, q) + i);
    return t;
}
#endif
