// kmg_common.cuh -- shared definitions for the B200 (sm_100a) Gram-construction kernels.
//
// Sequence layout in HBM ("planes"): every DNA sequence (L <= 128 bases, alphabet A<C<G<T = 0..3,
// reference kernels.py:37,184) is stored as two 128-bit bit-planes in 8 little-endian u32 words:
//   words 0..3 : bit p = low  bit of the code of base p
//   words 4..7 : bit p = high bit of the code of base p
// bits >= L are zero.  32 B per sequence; 200 000 sequences = 6.4 MB, replicated on every GPU.
// Bit-planes (rather than 2 interleaved bits per base) make the per-base (mis)match vector of two
// sequences one LOP3 per word: ne = (xl ^ yl) | (xh ^ yh), no fold step.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define KMG_PLANE_WORDS 4
#define KMG_SEQ_WORDS 8
#define KMG_MAX_L 128

// error codes of the C-ABI (include/kmg.h)
#define KMG_OK 0
#define KMG_ERR_CUDA (-1)
#define KMG_ERR_ARG (-2)
#define KMG_ERR_ALPHABET (-3)
#define KMG_ERR_UNSUPPORTED (-4)
#define KMG_ERR_NOMEM (-5)

void kmg_set_error(const char* fmt, ...);

#define KMG_CUDA_CHECK(expr)                                                                   \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            kmg_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return KMG_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

#define KMG_REQUIRE(cond, code, ...)                                                           \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            kmg_set_error(__VA_ARGS__);                                                        \
            return (code);                                                                     \
        }                                                                                      \
    } while (0)

struct SeqPlanes {
    uint32_t lo[KMG_PLANE_WORDS];
    uint32_t hi[KMG_PLANE_WORDS];
};

__device__ __forceinline__ SeqPlanes kmg_load_planes(const uint32_t* __restrict__ planes, int64_t i) {
    const uint4* p = reinterpret_cast<const uint4*>(planes + i * KMG_SEQ_WORDS);
    uint4 a = __ldg(p), b = __ldg(p + 1);
    SeqPlanes s;
    s.lo[0] = a.x; s.lo[1] = a.y; s.lo[2] = a.z; s.lo[3] = a.w;
    s.hi[0] = b.x; s.hi[1] = b.y; s.hi[2] = b.z; s.hi[3] = b.w;
    return s;
}

// 128-bit logical shift right by 1 (bit p <- bit p+1), words little-endian.
__device__ __forceinline__ void kmg_shr1_128(uint32_t (&w)[4]) {
    w[0] = __funnelshift_r(w[0], w[1], 1);
    w[1] = __funnelshift_r(w[1], w[2], 1);
    w[2] = __funnelshift_r(w[2], w[3], 1);
    w[3] = w[3] >> 1;
}

// 128-bit rotate right by 1 (bit p <- bit (p+1) mod 128).
__device__ __forceinline__ void kmg_rotr1_128(uint32_t (&w)[4]) {
    uint32_t w0 = w[0];
    w[0] = __funnelshift_r(w[0], w[1], 1);
    w[1] = __funnelshift_r(w[1], w[2], 1);
    w[2] = __funnelshift_r(w[2], w[3], 1);
    w[3] = __funnelshift_r(w[3], w0, 1);
}

// rotate right by one inside the low L bits (bits >= L are and stay zero); wrap[] has the single bit L-1 set
__device__ __forceinline__ void kmg_rotr1_len(uint32_t (&w)[4], const uint32_t (&wrap)[4]) {
    const uint32_t carry = 0u - (w[0] & 1u);
    w[0] = __funnelshift_r(w[0], w[1], 1);
    w[1] = __funnelshift_r(w[1], w[2], 1);
    w[2] = __funnelshift_r(w[2], w[3], 1);
    w[3] = w[3] >> 1;
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] |= carry & wrap[j];
}

// out = in >> S (128-bit logical, compile-time S in [0,127]).
template <int S>
__device__ __forceinline__ void kmg_shr_128(const uint32_t (&in)[4], uint32_t (&out)[4]) {
    constexpr int ws = S / 32, bs = S % 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t lo = (i + ws < 4) ? in[i + ws] : 0u;
        uint32_t hi = (i + ws + 1 < 4) ? in[i + ws + 1] : 0u;
        out[i] = bs == 0 ? lo : __funnelshift_r(lo, hi, bs);
    }
}

// mask with bits [a, b] set (inclusive, 0 <= a, b <= 127); empty if a > b.  Host + device.
__host__ __device__ inline void kmg_range_mask_128(int a, int b, uint32_t* m) {
    for (int w = 0; w < 4; ++w) {
        int lo = a - 32 * w, hi = b - 32 * w;
        if (lo < 0) lo = 0;
        if (hi > 31) hi = 31;
        m[w] |= (lo <= hi) ? ((hi == 31 ? 0xFFFFFFFFu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u)) : 0u;
    }
}
