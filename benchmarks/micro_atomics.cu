// Micro-benchmark (measurement only, not product code): rates that decide the design of the per-gene moment
// accumulation of K1 (alona/cell.py:241-243 + alona/hvg.py:148,157) on B200:
//   shared-memory ATOMS.ADD.32 and 64-bit CAS-loop adds at random table addresses, global RED.ADD.F64 / .U64
//   at random addresses of an L2-resident table, and the FP64 FMA rate of the CUDA cores.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o benchmarks/build/micro_atomics benchmarks/micro_atomics.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t rng(uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int MODE>
__global__ void __launch_bounds__(1024) k_smem(int iters, int table, unsigned long long *sink) {
    extern __shared__ unsigned long long sm[];
    for (int i = threadIdx.x; i < table; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    uint32_t s = blockIdx.x * 9781u + threadIdx.x * 31u + 1u;
    unsigned int *s32 = reinterpret_cast<unsigned int *>(sm);
    for (int i = 0; i < iters; ++i) {
        const uint32_t r = rng(s);
        if (MODE == 0) atomicAdd(s32 + (r % (uint32_t)(2 * table)), r & 1023u);                      // ATOMS.ADD.32
        if (MODE == 1) atomicAdd(sm + (r % (uint32_t)table), (unsigned long long)r);                  // CAS loop 64
        if (MODE == 2) atomicAdd(reinterpret_cast<double *>(sm) + (r % (uint32_t)table), (double)r);  // CAS loop f64
        if (MODE == 3) { atomicAdd(s32 + (r % (uint32_t)(2 * table)), r & 1023u); atomicAdd(s32 + ((r >> 3) % (uint32_t)(2 * table)), 1u); }
    }
    __syncthreads();
    unsigned long long t = 0;
    for (int i = threadIdx.x; i < table; i += blockDim.x) t += sm[i];
    if (t == 12345ull) sink[0] = t;
}

template <int MODE>
__global__ void __launch_bounds__(256) k_gmem(int iters, int table, unsigned long long *tab) {
    uint32_t s = blockIdx.x * 9781u + threadIdx.x * 31u + 1u;
    for (int i = 0; i < iters; ++i) {
        const uint32_t r = rng(s);
        if (MODE == 0) atomicAdd(reinterpret_cast<double *>(tab) + (r % (uint32_t)table), 1.0);
        if (MODE == 1) atomicAdd(tab + (r % (uint32_t)table), (unsigned long long)r);
        if (MODE == 2) atomicAdd(reinterpret_cast<unsigned int *>(tab) + (r % (uint32_t)(2 * table)), 1u);
    }
}

__global__ void __launch_bounds__(1024) k_dfma(int iters, double *out) {
    double a0 = threadIdx.x, a1 = 1.0, a2 = 2.0, a3 = 3.0, a4 = 4.0, a5 = 5.0, a6 = 6.0, a7 = 7.0;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

template <typename F>
static float timeit(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    unsigned long long *buf;
    cudaMalloc(&buf, 64 << 20);
    cudaMemset(buf, 0, 64 << 20);
    const int iters = 4096;
    const double clk = p.clockRate * 1e3;
    printf("device %s, %d SMs, clock %.0f MHz\n", p.name, sms, clk / 1e6);
    for (int table : {7000, 14000, 28000}) {
        const size_t smem = (size_t)table * 8;
        if (smem > 227 * 1024) continue;
        cudaFuncSetAttribute(k_smem<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_smem<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_smem<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_smem<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        const double ops = (double)sms * 1024 * iters;
        float t0 = timeit([&] { k_smem<0><<<sms, 1024, smem>>>(iters, table, buf); });
        float t1 = timeit([&] { k_smem<1><<<sms, 1024, smem>>>(iters, table, buf); });
        float t2 = timeit([&] { k_smem<2><<<sms, 1024, smem>>>(iters, table, buf); });
        float t3 = timeit([&] { k_smem<3><<<sms, 1024, smem>>>(iters, table, buf); });
        printf("smem table %6d x 8 B: ATOMS.ADD.32 %.2f lanes/clk/SM | u64 CAS-loop %.2f | f64 CAS-loop %.2f | 2x ATOMS.32 per item %.2f items/clk/SM\n",
               table, ops / (t0 * 1e-3) / sms / clk, ops / (t1 * 1e-3) / sms / clk, ops / (t2 * 1e-3) / sms / clk,
               ops / (t3 * 1e-3) / sms / clk);
    }
    for (int table : {28000, 56000, 1 << 20}) {
        const int grid = sms * 8;
        const double ops = (double)grid * 256 * iters;
        float t0 = timeit([&] { k_gmem<0><<<grid, 256>>>(iters, table, buf); });
        float t1 = timeit([&] { k_gmem<1><<<grid, 256>>>(iters, table, buf); });
        float t2 = timeit([&] { k_gmem<2><<<grid, 256>>>(iters, table, buf); });
        printf("gmem table %7d: RED.F64 %.3f lanes/clk/SM (%.1f G/s) | RED.U64 %.3f (%.1f G/s) | RED.U32 %.3f (%.1f G/s)\n", table,
               ops / (t0 * 1e-3) / sms / clk, ops / (t0 * 1e-3) / 1e9, ops / (t1 * 1e-3) / sms / clk, ops / (t1 * 1e-3) / 1e9,
               ops / (t2 * 1e-3) / sms / clk, ops / (t2 * 1e-3) / 1e9);
    }
    {
        double *out = reinterpret_cast<double *>(buf);
        const int it2 = 1 << 16;
        float t = timeit([&] { k_dfma<<<sms * 2, 1024>>>(it2, out); });
        const double fmas = (double)sms * 2 * 1024 * it2 * 8;
        printf("DFMA: %.2f FMA/clk/SM = %.2f TFLOP/s\n", fmas / (t * 1e-3) / sms / clk, 2 * fmas / (t * 1e-3) / 1e12);
    }
    return 0;
}
