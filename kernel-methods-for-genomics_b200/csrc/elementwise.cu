// elementwise.cu -- HBM-bound passes over stored fp64 Grams: cosine normalisation, centring,
// sub-block gather, linear / polynomial kernel combination and the Frobenius / quadratic-form
// reductions ALIGNF and NLCK need.
//
//   normalize_K  kernels.py:398-415      center_K   kernels.py:387-395
//   ALIGNF       ALIGNF.py:28 (sub-block), :36-41 (centre), :43-48 (a), :50-58 (M), :91-94 (sum u_i K_i)
//   NLCK         NLCKernels.py:43-48 (normalise), :52 and :97 ((sum u_m K_m)**degree), :61-66 (grad)
//
// All kernels stream the matrices once with fully coalesced 8-byte accesses; reductions are
// deterministic two-stage trees (no atomics), so repeated runs are bit-identical.
#include <cuda_runtime.h>
#include <math.h>

#include <algorithm>
#include <stdint.h>

#include "elementwise.h"
#include "epi_ops.cuh"
#include "kmg_common.cuh"

namespace {

constexpr int EW_PER_THREAD = 4;  // elements per thread of the row-wise passes (memory-level parallelism)

__global__ void diag_sqrt_kernel(const double* __restrict__ K, int64_t n, int64_t ld, double* __restrict__ sd) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) sd[i] = sqrt(K[i * ld + i]);  // np.sqrt(np.diag(K)), kernels.py:408
}

// In place.  Only upper-triangle 32x32 tiles run; each reads K[i,j] (i<j), writes the quotient to
// (i,j) and -- through a shared-memory transpose -- to (j,i), exactly the mirror of kernels.py:412-413.
__global__ void __launch_bounds__(256) normalize_kernel(double* __restrict__ K, int64_t n, int64_t ld, const double* __restrict__ sd) {
    const int64_t bi = blockIdx.y, bj = blockIdx.x;
    if (bj < bi) return;  // (enumerating only the upper-triangle tile pairs was measured: no faster)
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int rr = ty; rr < 32; rr += 8) {
        const int64_t i = bi * 32 + rr, j = bj * 32 + tx;
        double v = 0.0;
        if (i < n && j < n) {
            if (i < j) {
                v = __ddiv_rn(K[i * ld + j], __dmul_rn(sd[i], sd[j]));
                K[i * ld + j] = v;
            } else if (i == j) {
                v = 1.0;
                K[i * ld + j] = 1.0;  // np.fill_diagonal(K, 1), kernels.py:414
            }
        }
        tile[rr][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int rr = ty; rr < 32; rr += 8) {
        // element (j', i') of the mirrored tile = tile[i'][j']
        const int64_t jrow = bj * 32 + rr, icol = bi * 32 + tx;
        if (jrow < n && icol < n && icol < jrow) K[jrow * ld + icol] = tile[tx][rr];
    }
}

// row sums: one warp per row
__global__ void __launch_bounds__(256) row_sum_kernel(const double* __restrict__ K, int64_t rows, int64_t cols, int64_t ld,
                                                      double* __restrict__ rs) {
    const int64_t row = blockIdx.x * 8ll + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    double acc = 0.0;
    for (int64_t j = lane; j < cols; j += 32) acc += K[row * ld + j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) rs[row] = acc;
}

// column sums, stage 1: thread owns a column, block owns a 256-column x CHUNK-row slab
constexpr int CS_CHUNK = 256;
__global__ void __launch_bounds__(256) col_sum_partial_kernel(const double* __restrict__ K, int64_t rows, int64_t cols, int64_t ld,
                                                              double* __restrict__ part) {
    const int64_t j = blockIdx.x * 256ll + threadIdx.x;
    const int64_t i0 = (int64_t)blockIdx.y * CS_CHUNK;
    if (j >= cols) return;
    const int64_t i1 = (i0 + CS_CHUNK < rows) ? i0 + CS_CHUNK : rows;
    double acc = 0.0;
    for (int64_t i = i0; i < i1; ++i) acc += K[i * ld + j];
    part[(int64_t)blockIdx.y * cols + j] = acc;
}
__global__ void col_sum_final_kernel(const double* __restrict__ part, int64_t nchunks, int64_t cols, double* __restrict__ cs) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= cols) return;
    double acc = 0.0;
    for (int64_t c = 0; c < nchunks; ++c) acc += part[c * cols + j];
    cs[j] = acc;
}

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = v;
    __syncthreads();
    v = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
    if (w == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    }
    __syncthreads();
    return v;  // valid in thread 0
}

__global__ void __launch_bounds__(256) vec_sum_kernel(const double* __restrict__ v, int64_t n, double* __restrict__ out) {
    __shared__ double sh[8];
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 256) acc += v[i];
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) *out = acc;
}

// out[i,j] = K[i,j] - cs[j]/n - rs[i]/n + g/n^2   (closed form of (I-11'/n) K (I-11'/n))
__global__ void __launch_bounds__(256) center_apply_kernel(const double* __restrict__ K, int64_t rows, int64_t cols, int64_t n, int64_t ld,
                                                           const double* __restrict__ rs, const double* __restrict__ cs,
                                                           const double* __restrict__ g, double* __restrict__ out, int64_t ldo) {
    // four columns per thread, 256 apart: four independent 8-byte loads in flight per thread (one per thread left the
    // pass at 4.4 TB/s, latency bound: 16 KB in flight per SM against the ~35 KB that 6.5 TB/s needs)
    const int64_t j0 = blockIdx.x * (256ll * EW_PER_THREAD) + threadIdx.x;
    const double inv = 1.0 / (double)n;
    const double gg = (*g) * inv * inv;
    double c[EW_PER_THREAD];
#pragma unroll
    for (int q = 0; q < EW_PER_THREAD; ++q) c[q] = (j0 + 256 * q < cols) ? cs[j0 + 256 * q] * inv : 0.0;
    for (int64_t i = blockIdx.y; i < rows; i += gridDim.y) {  // gridDim.y is capped at 65535: rows beyond it loop
        const double ri = rs[i] * inv;
        double v[EW_PER_THREAD];
#pragma unroll
        for (int q = 0; q < EW_PER_THREAD; ++q) {
            const int64_t j = j0 + 256 * q;
            v[q] = j < cols ? K[i * ld + j] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < EW_PER_THREAD; ++q) {
            const int64_t j = j0 + 256 * q;
            if (j < cols) out[i * ldo + j] = v[q] - c[q] - ri + gg;
        }
    }
}

__global__ void __launch_bounds__(256) gather_kernel(const double* __restrict__ K, int64_t ld, const int64_t* __restrict__ idx,
                                                     int64_t m, double* __restrict__ out, int64_t ldo) {
    const int64_t b = blockIdx.x * 256ll + threadIdx.x;
    if (b >= m) return;
    const int64_t ib = idx[b];
    for (int64_t a = blockIdx.y; a < m; a += gridDim.y) out[a * ldo + b] = K[idx[a] * ld + ib];
}

struct CombineParams {
    int p;
    int degree;
    const double* K[KMG_MAX_COMBINE];
    int64_t ld[KMG_MAX_COMBINE];
    double u[KMG_MAX_COMBINE];
};

// out = (sum_m u_m K_m) ** degree, products and sums rounded separately in the order numpy uses
// (np.sum(kernels * u[:,None,None], axis=0): K_0 u_0, then + K_1 u_1, ...).
__global__ void __launch_bounds__(256) combine_kernel(CombineParams cp, int64_t rows, int64_t cols, double* __restrict__ out, int64_t ldo) {
    const int64_t j0 = blockIdx.x * (256ll * EW_PER_THREAD) + threadIdx.x;
    for (int64_t i = blockIdx.y; i < rows; i += gridDim.y) {
    double acc[EW_PER_THREAD];
#pragma unroll
    for (int q = 0; q < EW_PER_THREAD; ++q) {
        const int64_t j = j0 + 256 * q;
        acc[q] = j < cols ? __dmul_rn(cp.K[0][i * cp.ld[0] + j], cp.u[0]) : 0.0;
    }
    for (int m = 1; m < cp.p; ++m) {
#pragma unroll
        for (int q = 0; q < EW_PER_THREAD; ++q) {
            const int64_t j = j0 + 256 * q;
            if (j < cols) acc[q] = __dadd_rn(acc[q], __dmul_rn(cp.K[m][i * cp.ld[m] + j], cp.u[m]));
        }
    }
#pragma unroll
    for (int q = 0; q < EW_PER_THREAD; ++q) {
        const int64_t j = j0 + 256 * q;
        if (j >= cols) continue;
        double v = acc[q];
        if (cp.degree == 0) v = 1.0;
        else if (cp.degree == 2) v = __dmul_rn(acc[q], acc[q]);  // numpy: x**2 -> np.square
        else if (cp.degree != 1) v = pow(acc[q], (double)cp.degree);
        out[i * ldo + j] = v;
    }
    }
}

// partial[b] = sum over the block's elements of A_ij * (B ? B_ij : 1) * (w ? w_i w_j : 1)
// With rA / rB (row sums of the symmetric A / B) and gA / gB (their grand sums) the factors are centred on the fly,
// Ac_ij = A_ij - rA_j/n - rA_i/n + gA/n^2 -- the entries of (I-11'/n) A (I-11'/n) (kernels.py:387-395) without storing them.
__global__ void __launch_bounds__(256) weighted_dot_partial_kernel(const double* __restrict__ A, int64_t lda, const double* __restrict__ B,
                                                                   int64_t ldb, const double* __restrict__ w, int64_t n,
                                                                   double* __restrict__ partial, const double* __restrict__ rA,
                                                                   const double* __restrict__ gA, const double* __restrict__ rB,
                                                                   const double* __restrict__ gB) {
    __shared__ double sh[8];
    const int64_t i = blockIdx.x;
    double acc = 0.0;
    const double wi = w ? w[i] : 1.0;
    const double inv = 1.0 / (double)n;
    const double ai = rA ? rA[i] * inv : 0.0, ag = rA ? (*gA) * inv * inv : 0.0;
    const double bi = rB ? rB[i] * inv : 0.0, bg = rB ? (*gB) * inv * inv : 0.0;
    for (int64_t j = threadIdx.x; j < n; j += 256) {
        double v = A[i * lda + j];
        if (rA) v = v - rA[j] * inv - ai + ag;   // same association as center_apply_kernel
        if (B) {
            double b = B[i * ldb + j];
            if (rB) b = b - rB[j] * inv - bi + bg;
            v *= b;
        }
        if (w) v *= w[j];
        acc += v;
    }
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) partial[i] = acc * wi;
}

// out[r] = sum over chunks (in order) of partial[r][chunk]: the second, fixed-order stage of the epilogue row statistics
__global__ void __launch_bounds__(256) partial_rows_reduce_kernel(const double* __restrict__ partial, int64_t rows, int64_t n_chunks,
                                                                  double* __restrict__ out) {
    const int64_t r = blockIdx.x * 256ll + threadIdx.x;
    if (r >= rows) return;
    double acc = 0.0;
    for (int64_t c = 0; c < n_chunks; ++c) acc += partial[r * n_chunks + c];
    out[r] = acc;
}

__global__ void __launch_bounds__(256) vec_dot_kernel(const double* __restrict__ a, const double* __restrict__ b, int64_t n, double* __restrict__ out) {
    __shared__ double sh[8];
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 256) acc += a[i] * b[i];
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) *out = acc;
}

// rows of K times a weight vector: out[r] = sum_j K[r][j] w[j] (one warp per row)
__global__ void __launch_bounds__(256) row_wsum_kernel(const double* __restrict__ K, int64_t rows, int64_t cols, int64_t ld,
                                                       const double* __restrict__ w, double* __restrict__ out) {
    const int64_t row = blockIdx.x * 8ll + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    double acc = 0.0;
    for (int64_t j = lane; j < cols; j += 32) acc += K[row * ld + j] * w[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[row] = acc;
}

__global__ void __launch_bounds__(256) gather_vec_kernel(const double* __restrict__ v, const int64_t* __restrict__ idx, int64_t m, double* __restrict__ out) {
    const int64_t t = blockIdx.x * 256ll + threadIdx.x;
    if (t < m) out[t] = v[idx[t]];
}

// accumulate a stored Gram into a combination the way the fused epilogues do (for kernels whose producer has no
// fused normalisation, epi_ops.cuh): out = [out +] u * (K / (sd_i sd_j), diag 1), then power / normalisation of the last term
__global__ void __launch_bounds__(256) accumulate_kernel(const double* __restrict__ K, int64_t ldk, const double* __restrict__ sd, EpiOps e,
                                                         int64_t n, double* __restrict__ out, int64_t ldo) {
    const int64_t j = blockIdx.x * 256ll + threadIdx.x;
    if (j >= n) return;
    for (int64_t i = blockIdx.y; i < n; i += gridDim.y) {
        double v = K[i * ldk + j];
        if (sd != nullptr) {
            v = __ddiv_rn(v, __dmul_rn(sd[i], sd[j]));
            if (i == j) v = 1.0;
        }
        const double prev = e.accumulate == 2 ? out[i * ldo + j] : 0.0;
        const bool nrm = e.post_sd_rows != nullptr;
        out[i * ldo + j] = epi_finish(e, v, prev, i == j, nrm ? e.post_sd_rows[i] : 1.0, nrm ? e.post_sd_cols[j] : 1.0);
    }
}

// s32 counts -> u16 for the host link (host_link.cu d2h_rows widens them back to double): 8 entries per thread, 16-byte
// stores.  Any entry outside 0..65535 raises *flag and the caller ships the s32 block instead.
__global__ void __launch_bounds__(256) narrow_u16_kernel(const int32_t* __restrict__ src, int64_t count, uint16_t* __restrict__ dst,
                                                         int* __restrict__ flag) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 8;
    uint32_t bad = 0;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < count; i += stride) {
        if (i + 8 <= count) {
            const int4 a = *reinterpret_cast<const int4*>(src + i), b = *reinterpret_cast<const int4*>(src + i + 4);
            bad |= (uint32_t)(a.x | a.y | a.z | a.w | b.x | b.y | b.z | b.w);
            uint4 o;
            o.x = (uint32_t)a.x | ((uint32_t)a.y << 16); o.y = (uint32_t)a.z | ((uint32_t)a.w << 16);
            o.z = (uint32_t)b.x | ((uint32_t)b.y << 16); o.w = (uint32_t)b.z | ((uint32_t)b.w << 16);
            *reinterpret_cast<uint4*>(dst + i) = o;
        } else {
            for (int64_t j = i; j < count; ++j) { bad |= (uint32_t)src[j]; dst[j] = (uint16_t)src[j]; }
        }
    }
    if (bad >> 16) *reinterpret_cast<volatile int*>(flag) = 1;  // negative values have the top bit set; the flag may live in mapped host memory
}

}  // namespace

int kmg_ew_narrow_u16(const int32_t* src, int64_t count, uint16_t* dst, int* flag, cudaStream_t s) {
    if (count <= 0) return 0;
    const int64_t blocks = std::min<int64_t>((count + 2047) / 2048, 148 * 16);
    narrow_u16_kernel<<<(unsigned)blocks, 256, 0, s>>>(src, count, dst, flag);
    KMG_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int kmg_ew_diag_sqrt(const double* K, int64_t n, int64_t ld, double* sd, cudaStream_t s) {
    if (n <= 0) return KMG_OK;
    diag_sqrt_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(K, n, ld, sd);
    KMG_CUDA_CHECK(cudaGetLastError());
    return KMG_OK;
}

int kmg_ew_normalize(double* K, int64_t n, int64_t ld, const double* sd, cudaStream_t s) {
    if (n <= 0) return KMG_OK;
    const unsigned t = (unsigned)((n + 31) / 32);
    KMG_REQUIRE(t <= 65535, KMG_ERR_ARG, "normalize: n too large for one launch");
    normalize_kernel<<<dim3(t, t), 256, 0, s>>>(K, n, ld, sd);
    KMG_CUDA_CHECK(cudaGetLastError());
    return KMG_OK;
}

int64_t kmg_ew_center_workspace(int64_t n) {
    const int64_t chunks = (n + CS_CHUNK - 1) / CS_CHUNK;
    return (chunks * n + 2 * n + 8) * (int64_t)sizeof(double);
}

int kmg_ew_center(const double* K, int64_t n, int64_t ld, double* out, int64_t ldo, void* workspace, cudaStream_t s) {
    if (n <= 0) return KMG_OK;
    KMG_REQUIRE(n <= 65535ll * 256, KMG_ERR_ARG, "center: n too large for one launch");
    const int64_t chunks = (n + CS_CHUNK - 1) / CS_CHUNK;
    double* part = reinterpret_cast<double*>(workspace);
    double* rs = part + chunks * n;
    double* cs = rs + n;
    double* g = cs + n;
    row_sum_kernel<<<(unsigned)((n + 7) / 8), 256, 0, s>>>(K, n, n, ld, rs);
    col_sum_partial_kernel<<<dim3((unsigned)((n + 255) / 256), (unsigned)chunks), 256, 0, s>>>(K, n, n, ld, part);
    col_sum_final_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(part, chunks, n, cs);
    vec_sum_kernel<<<1, 256, 0, s>>>(rs, n, g);
    center_apply_kernel<<<dim3((unsigned)((n + 256 * EW_PER_THREAD - 1) / (256 * EW_PER_THREAD)), (unsigned)std::min<int64_t>(n, 65535)), 256, 0, s>>>(K, n, n, n, ld, rs, cs, g, out, ldo);
    KMG_CUDA_CHECK(cudaGetLastError());
    return KMG_OK;
}

int kmg_ew_gather(const double* K, int64_t ld, const int64_t* idx, int64_t m, double* out, int64_t ldo, cudaStream_t s) {
    if (m <= 0) return KMG_OK;
    gather_kernel<<<dim3((unsigned)((m + 255) / 256), (unsigned)std::min<int64_t>(m, 65535)), 256, 0, s>>>(K, ld, idx, m, out, ldo);
    KMG_CUDA_CHECK(cudaGetLastError());
    return KMG_OK;
}

int kmg_ew_combine(const double* const* Ks, const int64_t* lds, const double* u, int p, int degree, int64_t rows, int64_t cols,
                   double* out, int64_t ldo, cudaStream_t s) {
    KMG_REQUIRE(p >= 1 && p <= KMG_MAX_COMBINE, KMG_ERR_ARG, "combine: between 1 and %d kernels", KMG_MAX_COMBINE);
    KMG_REQUIRE(degree >= 0 && degree <= 64, KMG_ERR_ARG, "combine: degree out of range");
    if (rows <= 0 || cols <= 0) return KMG_OK;
    CombineParams cp;
    cp.p = p; cp.degree = degree;
    for (int m = 0; m < p; ++m) { cp.K[m] = Ks[m]; cp.ld[m] = lds[m]; cp.u[m] = u[m]; }
    combine_kernel<<<dim3((unsigned)((cols + 256 * EW_PER_THREAD - 1) / (256 * EW_PER_THREAD)), (unsigned)std::min<int64_t>(rows, 65535)), 256, 0, s>>>(cp, rows, cols, out, ldo);
    KMG_CUDA_CHECK(cudaGetLastError());
    return KMG_OK;
}

int kmg_ew_weighted_dot(const double* A, int64_t lda, const double* B, int64_t ldb, const double* w, int64_t n, double* partial,
                        double* result, cudaStream_t s) {
    if (n <= 0) return KMG_OK;
    weighted_dot_partial_kernel<<<(unsigned)n, 256, 0, s>>>(A, lda, B, ldb, w, n, partial, nullptr, nullptr, nullptr, nullptr);
    vec_sum_kernel<<<1, 256, 0, s>>>(partial, n, result);
    KMG_CUDA_CHECK(cudaGetLastError());
    return KMG_OK;
}

int kmg_ew_centered_dot(const double* A, int64_t lda, const double* rA, const double* gA, const double* B, int64_t ldb, const double* rB,
                        const double* gB, int64_t n, double* partial, double* result, cudaStream_t s) {
    if (n <= 0) return KMG_OK;
    weighted_dot_partial_kernel<<<(unsigned)n, 256, 0, s>>>(A, lda, B, ldb, nullptr, n, partial, rA, gA, rB, gB);
    vec_sum_kernel<<<1, 256, 0, s>>>(partial, n, result);
    KMG_CUDA_CHECK(cudaGetLastError());
    return KMG_OK;
}

int kmg_ew_partial_rows_reduce(const double* partial, int64_t rows, int64_t n_chunks, double* out, cudaStream_t s) {
    if (rows <= 0) return KMG_OK;
    partial_rows_reduce_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, s>>>(partial, rows, n_chunks, out);
    KMG_CUDA_CHECK(cudaGetLastError());
    return KMG_OK;
}

int kmg_ew_vec_sum(const double* v, int64_t n, double* out, cudaStream_t s) {
    vec_sum_kernel<<<1, 256, 0, s>>>(v, n, out);
    KMG_CUDA_CHECK(cudaGetLastError());
    return KMG_OK;
}

int kmg_ew_vec_dot(const double* a, const double* b, int64_t n, double* out, cudaStream_t s) {
    vec_dot_kernel<<<1, 256, 0, s>>>(a, b, n, out);
    KMG_CUDA_CHECK(cudaGetLastError());
    return KMG_OK;
}

int kmg_ew_row_wsums(const double* K, int64_t rows, int64_t cols, int64_t ld, const double* w, double* out, cudaStream_t s) {
    if (rows <= 0) return KMG_OK;
    row_wsum_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(K, rows, cols, ld, w, out);
    KMG_CUDA_CHECK(cudaGetLastError());
    return KMG_OK;
}

int kmg_ew_gather_vec(const double* v, const int64_t* idx, int64_t m, double* out, cudaStream_t s) {
    if (m <= 0) return KMG_OK;
    gather_vec_kernel<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(v, idx, m, out);
    KMG_CUDA_CHECK(cudaGetLastError());
    return KMG_OK;
}

int kmg_ew_accumulate(const double* K, int64_t ldk, const double* sd, const EpiOps* e, int64_t n, double* out, int64_t ldo, cudaStream_t s) {
    if (n <= 0) return KMG_OK;
    accumulate_kernel<<<dim3((unsigned)((n + 255) / 256), (unsigned)std::min<int64_t>(n, 65535)), 256, 0, s>>>(K, ldk, sd, *e, n, out, ldo);
    KMG_CUDA_CHECK(cudaGetLastError());
    return KMG_OK;
}

// ---- pieces of the centring for a sharded (block-row) Gram: kmg/dist.py all-reduces the column sums
int kmg_ew_row_sums(const double* K, int64_t rows, int64_t cols, int64_t ld, double* rs, cudaStream_t s) {
    if (rows <= 0) return KMG_OK;
    row_sum_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(K, rows, cols, ld, rs);
    KMG_CUDA_CHECK(cudaGetLastError());
    return KMG_OK;
}

int64_t kmg_ew_col_sums_workspace(int64_t rows, int64_t cols) {
    return ((rows + CS_CHUNK - 1) / CS_CHUNK) * cols * (int64_t)sizeof(double);
}

int kmg_ew_col_sums(const double* K, int64_t rows, int64_t cols, int64_t ld, double* cs, void* workspace, cudaStream_t s) {
    if (cols <= 0) return KMG_OK;
    const int64_t chunks = (rows + CS_CHUNK - 1) / CS_CHUNK;
    KMG_REQUIRE(chunks <= 65535, KMG_ERR_ARG, "col_sums: too many rows for one launch");
    if (rows <= 0) { KMG_CUDA_CHECK(cudaMemsetAsync(cs, 0, (size_t)cols * 8, s)); return KMG_OK; }
    double* part = reinterpret_cast<double*>(workspace);
    col_sum_partial_kernel<<<dim3((unsigned)((cols + 255) / 256), (unsigned)chunks), 256, 0, s>>>(K, rows, cols, ld, part);
    col_sum_final_kernel<<<(unsigned)((cols + 255) / 256), 256, 0, s>>>(part, chunks, cols, cs);
    KMG_CUDA_CHECK(cudaGetLastError());
    return KMG_OK;
}

int kmg_ew_center_apply(const double* K, int64_t rows, int64_t cols, int64_t n_total, int64_t ld, const double* rs, const double* cs,
                        const double* g, double* out, int64_t ldo, cudaStream_t s) {
    if (rows <= 0 || cols <= 0) return KMG_OK;
    center_apply_kernel<<<dim3((unsigned)((cols + 256 * EW_PER_THREAD - 1) / (256 * EW_PER_THREAD)), (unsigned)std::min<int64_t>(rows, 65535)), 256, 0, s>>>(K, rows, cols, n_total, ld, rs, cs, g, out, ldo);
    KMG_CUDA_CHECK(cudaGetLastError());
    return KMG_OK;
}
