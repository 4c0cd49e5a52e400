// solve.cu -- the K_fit algebra of the reference's closed-form solvers on device-resident Grams:
//   KRR.fit     a = inv(K_fit + lambda n I) y                                   KRR.py:30-33
//   KLR.WKRR    alpha = W^1/2 inv(W^1/2 K_fit W^1/2 + n lambda I) W^1/2 z       KLR.py:41-57
// Both are one symmetric positive definite system  (S K S + c I) x = b  with S = diag(s) (identity for KRR): the matrix is
// formed in a workspace, factored by a blocked right-looking Cholesky (32-column panels: diagonal block in one CTA,
// triangular solve of the panel below it, symmetric rank-32 update of the trailing matrix) and the two triangular
// systems are solved by one CTA.  fp64 throughout; n is the fit set (1 501 rows in the reference's pipeline), so the
// whole thing is latency-, not throughput-bound: ~3 launches per 32 columns.  A non-positive pivot raises the flag.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <algorithm>

#include "elementwise.h"
#include "kmg_common.cuh"

namespace {

constexpr int NB = 32;

// A = S K S + c I (lower triangle and diagonal are what the factorisation reads; the full matrix is written)
__global__ void __launch_bounds__(256) form_kernel(const double* __restrict__ K, int64_t n, int64_t ld, const double* __restrict__ s, double c,
                                                   double* __restrict__ A) {
    const int64_t j = blockIdx.x * 256ll + threadIdx.x;
    if (j >= n) return;
    const double sj = s ? s[j] : 1.0;
    for (int64_t i = blockIdx.y; i < n; i += gridDim.y) {
        double v = K[i * ld + j];
        if (s) v = __dmul_rn(__dmul_rn(s[i], v), sj);
        if (i == j) v = __dadd_rn(v, c);
        A[i * n + j] = v;
    }
}

// Cholesky of the nb x nb diagonal block at (k0, k0), in place (lower); one CTA of NB x NB threads
__global__ void __launch_bounds__(NB * NB) potf2_kernel(double* __restrict__ A, int64_t n, int64_t k0, int nb, int* __restrict__ flag) {
    __shared__ double a[NB][NB + 1];
    const int i = threadIdx.y, j = threadIdx.x;
    a[i][j] = (i < nb && j < nb) ? A[(k0 + i) * n + k0 + j] : (i == j ? 1.0 : 0.0);
    __syncthreads();
    for (int k = 0; k < nb; ++k) {
        if (i == k && j == k) {
            const double d = a[k][k];
            if (!(d > 0.0)) *flag = 1;
            a[k][k] = sqrt(d);
        }
        __syncthreads();
        if (j == k && i > k) a[i][k] /= a[k][k];
        __syncthreads();
        if (i > k && j > k && j <= i) a[i][j] -= a[i][k] * a[j][k];
        __syncthreads();
    }
    if (i < nb && j < nb) A[(k0 + i) * n + k0 + j] = (j <= i) ? a[i][j] : 0.0;
}

// rows below the diagonal block: A[i, k0:k0+nb] <- A[i, k0:k0+nb] L11^-T (one thread per row)
__global__ void __launch_bounds__(128) trsm_kernel(double* __restrict__ A, int64_t n, int64_t k0, int nb) {
    __shared__ double l[NB][NB + 1];
    for (int t = threadIdx.x; t < NB * NB; t += 128) {
        const int r = t / NB, c = t % NB;
        l[r][c] = (r < nb && c < nb) ? A[(k0 + r) * n + k0 + c] : 0.0;
    }
    __syncthreads();
    const int64_t i = k0 + nb + blockIdx.x * 128ll + threadIdx.x;
    if (i >= n) return;
    double x[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) x[j] = j < nb ? A[i * n + k0 + j] : 0.0;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        if (j < nb) {
            double v = x[j];
#pragma unroll
            for (int t = 0; t < NB; ++t)
                if (t < j) v -= x[t] * l[j][t];
            x[j] = v / l[j][j];
        }
    }
#pragma unroll
    for (int j = 0; j < NB; ++j)
        if (j < nb) A[i * n + k0 + j] = x[j];
}

// trailing update, lower tiles only: A[i, j] -= sum_t A[i, k0+t] A[j, k0+t]  for i, j >= k0 + nb
__global__ void __launch_bounds__(NB * NB) syrk_kernel(double* __restrict__ A, int64_t n, int64_t k0, int nb) {
    if (blockIdx.x > blockIdx.y) return;  // upper tile
    __shared__ double li[NB][NB + 1], lj[NB][NB + 1];
    const int64_t base = k0 + nb;
    const int64_t i = base + blockIdx.y * (int64_t)NB + threadIdx.y, j = base + blockIdx.x * (int64_t)NB + threadIdx.x;
    const int64_t ri = base + blockIdx.y * (int64_t)NB + threadIdx.y, rj = base + blockIdx.x * (int64_t)NB + threadIdx.y;
    li[threadIdx.y][threadIdx.x] = (ri < n && threadIdx.x < nb) ? A[ri * n + k0 + threadIdx.x] : 0.0;
    lj[threadIdx.y][threadIdx.x] = (rj < n && threadIdx.x < nb) ? A[rj * n + k0 + threadIdx.x] : 0.0;
    __syncthreads();
    if (i >= n || j >= n || j > i) return;
    double acc = 0.0;
#pragma unroll
    for (int t = 0; t < NB; ++t) acc += li[threadIdx.y][t] * lj[threadIdx.x][t];
    A[i * n + j] -= acc;
}

// x = A^-1 b from the Cholesky factor (lower, row major): forward then backward substitution, one CTA
__global__ void __launch_bounds__(1024) chol_solve_kernel(const double* __restrict__ Lm, int64_t n, const double* __restrict__ b, double* __restrict__ x,
                                                          double* __restrict__ w /* n doubles of scratch */) {
    __shared__ double xb[NB];
    const int tid = threadIdx.x;
    for (int64_t i = tid; i < n; i += 1024) w[i] = b[i];
    __syncthreads();
    // L y = b
    for (int64_t k0 = 0; k0 < n; k0 += NB) {
        const int nb = (int)((n - k0 < NB) ? n - k0 : NB);
        if (tid == 0) {
            for (int j = 0; j < nb; ++j) {
                double v = w[k0 + j];
                for (int t = 0; t < j; ++t) v -= Lm[(k0 + j) * n + k0 + t] * xb[t];
                xb[j] = v / Lm[(k0 + j) * n + k0 + j];
            }
        }
        __syncthreads();
        if (tid < nb) w[k0 + tid] = xb[tid];
        for (int64_t i = k0 + nb + tid; i < n; i += 1024) {
            double v = w[i];
            for (int t = 0; t < nb; ++t) v -= Lm[i * n + k0 + t] * xb[t];
            w[i] = v;
        }
        __syncthreads();
    }
    // L' x = y
    for (int64_t k1 = n; k1 > 0; k1 -= NB) {
        const int64_t k0 = (k1 >= NB) ? k1 - NB : 0;
        const int nb = (int)(k1 - k0);
        if (tid == 0) {
            for (int j = nb - 1; j >= 0; --j) {
                double v = w[k0 + j];
                for (int t = j + 1; t < nb; ++t) v -= Lm[(k0 + t) * n + k0 + j] * xb[t];
                xb[j] = v / Lm[(k0 + j) * n + k0 + j];
            }
        }
        __syncthreads();
        if (tid < nb) w[k0 + tid] = xb[tid];
        for (int64_t i = tid; i < k0; i += 1024) {
            double v = w[i];
            for (int t = 0; t < nb; ++t) v -= Lm[(k0 + t) * n + i] * xb[t];
            w[i] = v;
        }
        __syncthreads();
        if (k0 == 0) break;
    }
    for (int64_t i = tid; i < n; i += 1024) x[i] = w[i];
}

}  // namespace

int64_t kmg_solve_workspace(int64_t n) { return (n * n + n + 8) * (int64_t)sizeof(double) + 64; }

// (S K S + c I) x = b;  s nullable (S = I).  *flag (device int inside the workspace) != 0 afterwards: not positive definite.
int kmg_spd_solve_launch(const double* K, int64_t n, int64_t ld, const double* s, double c, const double* b, double* x, void* work,
                         int** d_flag, cudaStream_t st) {
    KMG_REQUIRE(n >= 1 && ld >= n && K && b && x && work, KMG_ERR_ARG, "spd_solve: bad arguments");
    double* A = reinterpret_cast<double*>(work);
    double* w = A + n * n;
    int* flag = reinterpret_cast<int*>(w + n);
    KMG_CUDA_CHECK(cudaMemsetAsync(flag, 0, sizeof(int), st));
    form_kernel<<<dim3((unsigned)((n + 255) / 256), (unsigned)std::min<int64_t>(n, 65535)), 256, 0, st>>>(K, n, ld, s, c, A);
    for (int64_t k0 = 0; k0 < n; k0 += NB) {
        const int nb = (int)std::min<int64_t>(NB, n - k0);
        potf2_kernel<<<1, dim3(NB, NB), 0, st>>>(A, n, k0, nb, flag);
        const int64_t rest = n - k0 - nb;
        if (rest > 0) {
            trsm_kernel<<<(unsigned)((rest + 127) / 128), 128, 0, st>>>(A, n, k0, nb);
            const unsigned t = (unsigned)((rest + NB - 1) / NB);
            syrk_kernel<<<dim3(t, t), dim3(NB, NB), 0, st>>>(A, n, k0, nb);
        }
    }
    chol_solve_kernel<<<1, 1024, 0, st>>>(A, n, b, x, w);
    KMG_CUDA_CHECK(cudaGetLastError());
    if (d_flag) *d_flag = flag;
    return KMG_OK;
}
