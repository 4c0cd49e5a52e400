// alu_peak.cu -- measured issue peaks of the CUDA-core pipes that bound the pairwise kernels: the denominators of
// bench.py's per-kernel roofline fractions ("INT32 / popc issue peak, FP64 vector peak: microbenchmark
// required"; MEASURED_PEAKS.json carries HBM and bf16 numbers only).
//
//   kind 0  LOP3   three-input logic, the INT32 ALU pipe          (mismatch_kernel / wd_kernel: ALU-pipe bound)
//   kind 1  SHF    funnel shift, the same ALU pipe
//   kind 2  POPC   population count, the XU pipe                   (N_delta = popc(...), c_k = popc(...))
//   kind 3  DFMA   fp64 fused multiply-add, the FP64 pipe          (la_kernel: FP64-pipe bound)
//   kind 4  DADD   fp64 add
//   kind 5  DMUL   fp64 multiply
//
// Every thread runs CHAINS independent dependency chains of the one instruction, unrolled, in a counted loop: with 8
// resident warps per sub-partition and 8 chains per thread the pipe's issue rate, not its latency, sets the time.  The
// instructions are `asm volatile`, so the compiler can neither drop nor combine them; one value per thread is written
// at the end to keep the chains live.  ops = instructions executed per THREAD summed over the grid (a warp instruction
// counts 32).  The caller times the stream (kmg/device.py alu_peak).
#include <cuda_runtime.h>
#include <stdint.h>

#include "gram_i8.h"
#include "kmg_common.cuh"

namespace {

constexpr int AP_THREADS = 256;
constexpr int AP_CTAS_PER_SM = 4;  // 32 warps per SM: 8 per sub-partition
constexpr int AP_CHAINS = 8;
constexpr int AP_UNROLL = 8;

template <int KIND>
__global__ void __launch_bounds__(AP_THREADS) alu_peak_kernel(int iters, uint32_t seed, uint32_t* __restrict__ sink) {
    const uint32_t t = blockIdx.x * AP_THREADS + threadIdx.x;
    if (KIND <= 2) {
        uint32_t x[AP_CHAINS];
#pragma unroll
        for (int c = 0; c < AP_CHAINS; ++c) x[c] = (t + seed) * 2654435761u + c * 40503u;
        const uint32_t y = seed | 0x5a5a5a5au, z = ~seed;
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < AP_UNROLL; ++u) {
#pragma unroll
                for (int c = 0; c < AP_CHAINS; ++c) {
                    if (KIND == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(y), "r"(z));
                    else if (KIND == 1) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(x[c]) : "r"(y));
                    else asm volatile("popc.b32 %0, %0;" : "+r"(x[c]));
                }
            }
        }
        uint32_t acc = 0;
#pragma unroll
        for (int c = 0; c < AP_CHAINS; ++c) acc ^= x[c];
        if (acc == 0xdeadbeefu) sink[t & 1023] = acc;  // practically never true: keeps the chains live without a store stream
    } else {
        double x[AP_CHAINS];
#pragma unroll
        for (int c = 0; c < AP_CHAINS; ++c) x[c] = 1.0 + 1e-9 * (double)((t + c) & 1023);
        const double a = 1.0 + 1e-12 * (double)(seed & 15), b = 1e-300;
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < AP_UNROLL; ++u) {
#pragma unroll
                for (int c = 0; c < AP_CHAINS; ++c) {
                    if (KIND == 3) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[c]) : "d"(a), "d"(b));
                    else if (KIND == 4) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x[c]) : "d"(b));
                    else asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(x[c]) : "d"(a));
                }
            }
        }
        double acc = 0.0;
#pragma unroll
        for (int c = 0; c < AP_CHAINS; ++c) acc += x[c];
        if (acc == 12345.678) sink[t & 1023] = 1u;
    }
}

}  // namespace

int kmg_alu_peak_launch(int kind, int iters, int64_t* ops, cudaStream_t stream) {
    KMG_REQUIRE(kind >= 0 && kind <= 5 && iters >= 1 && ops != nullptr, KMG_ERR_ARG, "alu_peak: kind 0..5, iters >= 1");
    int dev = 0, sms = 0;
    KMG_CUDA_CHECK(cudaGetDevice(&dev));
    KMG_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    static uint32_t* sink[64] = {};
    if (sink[dev & 63] == nullptr) KMG_CUDA_CHECK(cudaMalloc(&sink[dev & 63], 1024 * sizeof(uint32_t)));
    const unsigned grid = (unsigned)(sms * AP_CTAS_PER_SM);
    static uint32_t seed = 1;
    ++seed;
    switch (kind) {
        case 0: alu_peak_kernel<0><<<grid, AP_THREADS, 0, stream>>>(iters, seed, sink[dev & 63]); break;
        case 1: alu_peak_kernel<1><<<grid, AP_THREADS, 0, stream>>>(iters, seed, sink[dev & 63]); break;
        case 2: alu_peak_kernel<2><<<grid, AP_THREADS, 0, stream>>>(iters, seed, sink[dev & 63]); break;
        case 3: alu_peak_kernel<3><<<grid, AP_THREADS, 0, stream>>>(iters, seed, sink[dev & 63]); break;
        case 4: alu_peak_kernel<4><<<grid, AP_THREADS, 0, stream>>>(iters, seed, sink[dev & 63]); break;
        default: alu_peak_kernel<5><<<grid, AP_THREADS, 0, stream>>>(iters, seed, sink[dev & 63]); break;
    }
    KMG_CUDA_CHECK(cudaGetLastError());
    *ops = (int64_t)grid * AP_THREADS * (int64_t)iters * AP_UNROLL * AP_CHAINS;
    return KMG_OK;
}
