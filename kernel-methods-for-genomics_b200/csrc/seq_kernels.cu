// seq_kernels.cu -- sequence encoding and the dense spectrum feature map.
//
//  * pack_planes_kernel : ASCII 'ACGT' (n x L bytes) -> two 128-bit bit-planes per sequence
//    (kmg_common.cuh).  Replaces letter_to_num/format (kernels.py:178-193).  One warp per sequence,
//    __ballot_sync builds one plane word per 32 bases.  Any byte outside {A,C,G,T} raises the error
//    flag (the reference raises ValueError from int() in `format` for MM/LA).
//  * spectrum_phi_kernel : planes -> int8 Phi (n x Dpad), the concatenation over the list `ks` of
//    the 4^k-wide k-mer count vectors of get_phi_u (kernels.py:12-25), columns in
//    product('ACGT', repeat=k) order, zero padded to a multiple of 128 (one TMA/UMMA K-slab).
//    One CTA per sequence: the row is assembled in shared memory (zero, scatter-increment with
//    byte-lane atomics, counts <= 101 so no carry between byte lanes) and written out once, fully
//    coalesced -- HBM traffic is exactly one write of Phi.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "kmg_common.cuh"
#include "seq_kernels.h"

namespace {

__global__ void pack_planes_kernel(const uint8_t* __restrict__ ascii, int64_t n, int L, uint32_t* __restrict__ planes,
                                   int* __restrict__ err) {
    const int lane = threadIdx.x & 31;
    const int64_t seq = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (seq >= n) return;
    const uint8_t* s = ascii + seq * L;
    bool bad = false;
#pragma unroll
    for (int w = 0; w < KMG_PLANE_WORDS; ++w) {
        const int pos = w * 32 + lane;
        int code = 0;
        if (pos < L) {
            const uint8_t ch = __ldg(s + pos);
            // A=0x41 C=0x43 G=0x47 T=0x54
            if (ch == 'A') code = 0;
            else if (ch == 'C') code = 1;
            else if (ch == 'G') code = 2;
            else if (ch == 'T') code = 3;
            else bad = true;
        }
        const uint32_t lo = __ballot_sync(0xffffffffu, code & 1);
        const uint32_t hi = __ballot_sync(0xffffffffu, code & 2);
        if (lane == 0) {
            planes[seq * KMG_SEQ_WORDS + w] = lo;
            planes[seq * KMG_SEQ_WORDS + KMG_PLANE_WORDS + w] = hi;
        }
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicExch(err, 1);
}

// codes (uint8 0..3) -> planes, same layout; used when the caller already holds integer codes.
__global__ void pack_codes_kernel(const uint8_t* __restrict__ codes, int64_t n, int L, uint32_t* __restrict__ planes,
                                  int* __restrict__ err) {
    const int lane = threadIdx.x & 31;
    const int64_t seq = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (seq >= n) return;
    const uint8_t* s = codes + seq * L;
    bool bad = false;
#pragma unroll
    for (int w = 0; w < KMG_PLANE_WORDS; ++w) {
        const int pos = w * 32 + lane;
        int code = 0;
        if (pos < L) {
            code = __ldg(s + pos);
            if (code > 3) { bad = true; code = 0; }
        }
        const uint32_t lo = __ballot_sync(0xffffffffu, code & 1);
        const uint32_t hi = __ballot_sync(0xffffffffu, code & 2);
        if (lane == 0) {
            planes[seq * KMG_SEQ_WORDS + w] = lo;
            planes[seq * KMG_SEQ_WORDS + KMG_PLANE_WORDS + w] = hi;
        }
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicExch(err, 1);
}

struct PhiParams {
    int nk;
    int ks[KMG_MAX_KS];
    int64_t off[KMG_MAX_KS];  // column offset of each k's segment
};

__global__ void __launch_bounds__(128)
spectrum_phi_kernel(const uint32_t* __restrict__ planes, int64_t n, int L, PhiParams pp, int64_t Dpad,
                    int8_t* __restrict__ phi, int64_t ld) {
    extern __shared__ uint32_t row32[];  // Dpad bytes
    __shared__ uint8_t codes[KMG_MAX_L];
    const int64_t words = Dpad / 4;
    // The row image is zeroed ONCE; after a row has been written out, the (at most L per k) counters it touched are
    // cleared again by the threads that incremented them: ~700 shared-memory words per sequence instead of a 21 888-byte
    // sweep (k = 1..7), which was half of the kernel's shared-memory traffic.
    for (int64_t w = threadIdx.x; w < words; w += blockDim.x) row32[w] = 0u;
    __syncthreads();
    for (int64_t seq = blockIdx.x; seq < n; seq += gridDim.x) {
        if (threadIdx.x < KMG_MAX_L) {
            const int pos = threadIdx.x;
            const uint32_t lo = __ldg(planes + seq * KMG_SEQ_WORDS + (pos >> 5));
            const uint32_t hi = __ldg(planes + seq * KMG_SEQ_WORDS + KMG_PLANE_WORDS + (pos >> 5));
            codes[pos] = (uint8_t)(((lo >> (pos & 31)) & 1u) | (((hi >> (pos & 31)) & 1u) << 1));
        }
        __syncthreads();
        // thread p owns window start p for every k in the list
        for (int q = 0; q < pp.nk; ++q) {
            const int k = pp.ks[q];
            const int p = threadIdx.x;
            if (p + k <= L) {
                uint32_t idx = 0;
                for (int t = 0; t < k; ++t) idx = (idx << 2) | codes[p + t];
                const int64_t col = pp.off[q] + idx;
                atomicAdd(&row32[col >> 2], 1u << (8 * (col & 3)));
            }
        }
        __syncthreads();
        uint4* dst = reinterpret_cast<uint4*>(phi + seq * ld);
        const uint4* src = reinterpret_cast<const uint4*>(row32);
        for (int64_t w = threadIdx.x; w < Dpad / 16; w += blockDim.x) dst[w] = src[w];
        __syncthreads();
        for (int q = 0; q < pp.nk; ++q) {  // un-count: the same windows, the same words
            const int k = pp.ks[q];
            const int p = threadIdx.x;
            if (p + k <= L) {
                uint32_t idx = 0;
                for (int t = 0; t < k; ++t) idx = (idx << 2) | codes[p + t];
                row32[(pp.off[q] + idx) >> 2] = 0u;
            }
        }
        __syncthreads();
    }
}

// Dense (k,m)-mismatch feature map (get_phi_km, kernels.py:161-175): phi[b] = #windows of the sequence
// within Hamming distance m of k-mer b.  Thread p owns window p and scatter-increments every k-mer
// of its m-neighbourhood (distinct position sets x non-zero XOR substitutions => each neighbour
// exactly once per window, so phi[b] <= #windows <= 127 fits one byte lane).
__device__ __forceinline__ void phi_add(uint32_t* row32, uint32_t col) { atomicAdd(&row32[col >> 2], 1u << (8 * (col & 3))); }

__global__ void __launch_bounds__(128)
mismatch_phi_kernel(const uint32_t* __restrict__ planes, int64_t n, int L, int k, int m, int64_t Dpad,
                    int8_t* __restrict__ phi, int64_t ld) {
    extern __shared__ uint32_t row32[];
    __shared__ uint8_t codes[KMG_MAX_L];
    const int64_t words = Dpad / 4;
    for (int64_t seq = blockIdx.x; seq < n; seq += gridDim.x) {
        for (int64_t w = threadIdx.x; w < words; w += blockDim.x) row32[w] = 0u;
        if (threadIdx.x < KMG_MAX_L) {
            const int pos = threadIdx.x;
            const uint32_t lo = __ldg(planes + seq * KMG_SEQ_WORDS + (pos >> 5));
            const uint32_t hi = __ldg(planes + seq * KMG_SEQ_WORDS + KMG_PLANE_WORDS + (pos >> 5));
            codes[pos] = (uint8_t)(((lo >> (pos & 31)) & 1u) | (((hi >> (pos & 31)) & 1u) << 1));
        }
        __syncthreads();
        const int p = threadIdx.x;
        if (p + k <= L) {
            uint32_t idx = 0;
            for (int t = 0; t < k; ++t) idx = (idx << 2) | codes[p + t];
            phi_add(row32, idx);
            if (m >= 1)
                for (int t1 = 0; t1 < k; ++t1)
                    for (uint32_t s1 = 1; s1 < 4; ++s1) {
                        const uint32_t i1 = idx ^ (s1 << (2 * t1));
                        phi_add(row32, i1);
                        if (m >= 2)
                            for (int t2 = t1 + 1; t2 < k; ++t2)
                                for (uint32_t s2 = 1; s2 < 4; ++s2) {
                                    const uint32_t i2 = i1 ^ (s2 << (2 * t2));
                                    phi_add(row32, i2);
                                    if (m >= 3)
                                        for (int t3 = t2 + 1; t3 < k; ++t3)
                                            for (uint32_t s3 = 1; s3 < 4; ++s3) phi_add(row32, i2 ^ (s3 << (2 * t3)));
                                }
                    }
        }
        __syncthreads();
        uint4* dst = reinterpret_cast<uint4*>(phi + seq * ld);
        const uint4* src = reinterpret_cast<const uint4*>(row32);
        for (int64_t w = threadIdx.x; w < Dpad / 16; w += blockDim.x) dst[w] = src[w];
        __syncthreads();
    }
}

// sd[i] = sqrt(sum_t Phi[i][t]^2): one warp per row, dp4a on 4 features at a time (exact integer sum)
__global__ void __launch_bounds__(256) phi_diag_sqrt_kernel(const int8_t* __restrict__ phi, int64_t n, int64_t width, int64_t ld,
                                                            double* __restrict__ sd) {
    const int64_t row = blockIdx.x * 8ll + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const int* r = reinterpret_cast<const int*>(phi + row * ld);
    int acc = 0;
    for (int64_t t = lane; t < width / 4; t += 32) {
        const int v = __ldg(r + t);
        acc = __dp4a(v, v, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) sd[row] = sqrt((double)acc);
}

}  // namespace

int kmg_phi_diag_sqrt_launch(const int8_t* d_phi, int64_t n, int64_t width, int64_t ld, double* d_sd, cudaStream_t stream) {
    KMG_REQUIRE(width % 4 == 0 && ld % 4 == 0, KMG_ERR_ARG, "phi_diag_sqrt: width and ld must be multiples of 4");
    if (n <= 0) return KMG_OK;
    phi_diag_sqrt_kernel<<<(unsigned)((n + 7) / 8), 256, 0, stream>>>(d_phi, n, width, ld, d_sd);
    KMG_CUDA_CHECK(cudaGetLastError());
    return KMG_OK;
}

int kmg_mismatch_phi_launch(const uint32_t* d_planes, int64_t n, int L, int k, int m, int8_t* d_phi, int64_t ld, cudaStream_t stream) {
    KMG_REQUIRE(L >= 1 && L <= KMG_MAX_L, KMG_ERR_UNSUPPORTED, "sequence length %d not supported (1..%d)", L, KMG_MAX_L);
    KMG_REQUIRE(k >= 1 && k <= KMG_MAX_DENSE_K && k <= L, KMG_ERR_UNSUPPORTED, "dense mismatch feature map supports 1 <= k <= %d", KMG_MAX_DENSE_K);
    KMG_REQUIRE(m >= 0 && m <= 3, KMG_ERR_UNSUPPORTED, "dense mismatch feature map supports 0 <= m <= 3");
    KMG_REQUIRE(L - k + 1 <= 127, KMG_ERR_UNSUPPORTED, "window count must fit int8");
    const int64_t Dpad = ((1ll << (2 * k)) + 127) / 128 * 128;
    KMG_REQUIRE(ld >= Dpad && ld % 16 == 0, KMG_ERR_ARG, "mismatch_phi: ld must be >= %lld and a multiple of 16", (long long)Dpad);
    if (n <= 0) return KMG_OK;
    static bool attr_set[64] = {};
    int dev = 0;
    KMG_CUDA_CHECK(cudaGetDevice(&dev));
    if (!attr_set[dev & 63]) {
        KMG_CUDA_CHECK(cudaFuncSetAttribute(mismatch_phi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        attr_set[dev & 63] = true;
    }
    int sms = 0;
    KMG_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int per_sm = (int)((200 * 1024) / (Dpad + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 16) per_sm = 16;
    int64_t grid = (int64_t)sms * per_sm;
    if (grid > n) grid = n;
    mismatch_phi_kernel<<<(unsigned)grid, 128, (size_t)Dpad, stream>>>(d_planes, n, L, k, m, Dpad, d_phi, ld);
    KMG_CUDA_CHECK(cudaGetLastError());
    return KMG_OK;
}

int kmg_pack_launch(const uint8_t* d_in, int is_ascii, int64_t n, int L, uint32_t* d_planes, int* d_err, cudaStream_t stream) {
    KMG_REQUIRE(L >= 1 && L <= KMG_MAX_L, KMG_ERR_UNSUPPORTED, "sequence length %d not supported (1..%d)", L, KMG_MAX_L);
    if (n <= 0) return KMG_OK;
    const int threads = 256;
    const int64_t blocks = (n * 32 + threads - 1) / threads;
    if (is_ascii)
        pack_planes_kernel<<<(unsigned)blocks, threads, 0, stream>>>(d_in, n, L, d_planes, d_err);
    else
        pack_codes_kernel<<<(unsigned)blocks, threads, 0, stream>>>(d_in, n, L, d_planes, d_err);
    KMG_CUDA_CHECK(cudaGetLastError());
    return KMG_OK;
}

int64_t kmg_spectrum_width(const int* ks, int nk, int L) {
    (void)L;
    int64_t D = 0;
    for (int q = 0; q < nk; ++q) D += 1ll << (2 * ks[q]);
    return D;
}

int64_t kmg_spectrum_padded_width(const int* ks, int nk, int L) {
    const int64_t D = kmg_spectrum_width(ks, nk, L);
    return (D + 127) / 128 * 128;
}

int kmg_spectrum_phi_launch(const uint32_t* d_planes, int64_t n, int L, const int* ks, int nk, int8_t* d_phi, int64_t ld,
                            cudaStream_t stream) {
    KMG_REQUIRE(nk >= 1 && nk <= KMG_MAX_KS, KMG_ERR_ARG, "spectrum: between 1 and %d values of k", KMG_MAX_KS);
    KMG_REQUIRE(L >= 1 && L <= KMG_MAX_L, KMG_ERR_UNSUPPORTED, "sequence length %d not supported (1..%d)", L, KMG_MAX_L);
    PhiParams pp;
    pp.nk = nk;
    int64_t off = 0;
    for (int q = 0; q < nk; ++q) {
        KMG_REQUIRE(ks[q] >= 1 && ks[q] <= KMG_MAX_DENSE_K, KMG_ERR_UNSUPPORTED,
                    "dense spectrum feature map supports 1 <= k <= %d (got %d)", KMG_MAX_DENSE_K, ks[q]);
        KMG_REQUIRE(L - ks[q] + 1 <= 127, KMG_ERR_UNSUPPORTED, "window count must fit int8");
        pp.ks[q] = ks[q];
        pp.off[q] = off;
        off += 1ll << (2 * ks[q]);
    }
    const int64_t Dpad = (off + 127) / 128 * 128;
    KMG_REQUIRE(ld >= Dpad && ld % 16 == 0, KMG_ERR_ARG, "spectrum_phi: ld must be >= %lld and a multiple of 16", (long long)Dpad);
    KMG_REQUIRE(Dpad <= 160 * 1024, KMG_ERR_UNSUPPORTED, "spectrum_phi: feature row (%lld B) exceeds shared memory", (long long)Dpad);
    if (n <= 0) return KMG_OK;
    static bool attr_set[64] = {};
    int dev = 0;
    KMG_CUDA_CHECK(cudaGetDevice(&dev));
    if (!attr_set[dev & 63]) {
        KMG_CUDA_CHECK(cudaFuncSetAttribute(spectrum_phi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        attr_set[dev & 63] = true;
    }
    int sms = 0;
    KMG_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int per_sm = (int)((200 * 1024) / (Dpad + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 16) per_sm = 16;
    int64_t grid = (int64_t)sms * per_sm;
    if (grid > n) grid = n;
    // (256 threads per CTA were measured: 1.54 ms against 1.49 ms for the k = 1..7 map of 200 000 sequences)
    spectrum_phi_kernel<<<(unsigned)grid, 128, (size_t)Dpad, stream>>>(d_planes, n, L, pp, Dpad, d_phi, ld);
    KMG_CUDA_CHECK(cudaGetLastError());
    return KMG_OK;
}
