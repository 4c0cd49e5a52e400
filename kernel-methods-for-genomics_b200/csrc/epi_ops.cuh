// epi_ops.cuh -- what a Gram kernel's epilogue can do with an entry besides storing it: the ALIGNF / NLCK steps that the
// reference applies to finished Gram matrices, fused into the kernels that produce the entries.
//
//   accumulate   Km = sum_m u_m K_m                    ALIGNF.get_K (ALIGNF.py:93), NLCK K-line / get_K (NLCKernels.py:52,97)
//                one launch per kernel m, in order: the first stores u_0 K_0, every later one adds u_m K_m to what is
//                there -- product and sum rounded separately, the order numpy's (kernels * u[:,None,None]).sum(0) uses,
//                so the result is bit-identical to the reference's and no K_m is ever materialised.
//   post_degree  (.)**degree on the LAST term            NLCKernels.py:52,97 (degree 2 = x*x like np.square, else pow)
//   post_sd      normalize_K of the combination          NLCKernels.py:99 (kernels.py:398-415), diagonal := 1
//   row partials r = K 1 and K w per 32-column chunk     the centring statistics of ALIGNF.center (ALIGNF.py:36-41) and
//                a = y~' K y~ (ALIGNF.py:43-48): written per (row, chunk) without atomics, reduced in a fixed order.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

struct EpiOps {
    int accumulate;              // 0: store v;  1: store u*v;  2: store previous + u*v
    double u;
    int post_degree;             // <= 1: none
    const double* post_sd_rows;  // nullable: sqrt(diag) of the combination for the block's rows / columns
    const double* post_sd_cols;
    double* row_sum_partial;     // nullable: [rows][n_chunks] sum of the block's stored values over columns 32*chunk .. +31
    double* row_wsum_partial;    // nullable: same, every value weighted by w_cols[column]
    const double* w_cols;
    int64_t n_chunks;
};

__host__ __device__ inline bool epi_active(const EpiOps& e) {
    return e.accumulate != 0 || e.post_degree >= 2 || e.post_sd_rows != nullptr || e.row_sum_partial != nullptr || e.row_wsum_partial != nullptr;
}

// v: the kernel's own value (after its fused cosine normalisation, if any); prev: what the output holds (accumulate == 2)
__device__ __forceinline__ double epi_finish(const EpiOps& e, double v, double prev, bool on_diag, double psr, double psc) {
    if (e.accumulate) {
        const double t = __dmul_rn(v, e.u);
        v = e.accumulate == 2 ? __dadd_rn(prev, t) : t;
    }
    if (e.post_degree >= 2) v = e.post_degree == 2 ? __dmul_rn(v, v) : pow(v, (double)e.post_degree);
    if (e.post_sd_rows != nullptr) {
        v = __ddiv_rn(v, __dmul_rn(psr, psc));
        if (on_diag) v = 1.0;
    }
    return v;
}

// A warp holds 32 consecutive columns [c0, c0 + 32) of ONE row r (lane = column offset): fixed-order shuffle tree, lane 0 writes.
__device__ __forceinline__ void epi_row_partial_warp(const EpiOps& e, int64_t r, int64_t c0, double v, bool valid, int lane) {
    if (e.row_sum_partial == nullptr && e.row_wsum_partial == nullptr) return;
    double s = valid ? v : 0.0;
    double w = (valid && e.row_wsum_partial != nullptr) ? __dmul_rn(v, e.w_cols[c0 + lane]) : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        w += __shfl_xor_sync(0xffffffffu, w, o);
    }
    if (lane == 0) {
        if (e.row_sum_partial != nullptr) e.row_sum_partial[r * e.n_chunks + (c0 >> 5)] = s;
        if (e.row_wsum_partial != nullptr) e.row_wsum_partial[r * e.n_chunks + (c0 >> 5)] = w;
    }
}
