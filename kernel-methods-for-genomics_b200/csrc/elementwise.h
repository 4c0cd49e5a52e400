// elementwise.h -- internal launch interface of the stored-Gram passes (elementwise.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "epi_ops.cuh"

#define KMG_MAX_COMBINE 16

int kmg_ew_diag_sqrt(const double* K, int64_t n, int64_t ld, double* sd, cudaStream_t s);
int kmg_ew_normalize(double* K, int64_t n, int64_t ld, const double* sd, cudaStream_t s);
int64_t kmg_ew_center_workspace(int64_t n);
int kmg_ew_center(const double* K, int64_t n, int64_t ld, double* out, int64_t ldo, void* workspace, cudaStream_t s);
int kmg_ew_gather(const double* K, int64_t ld, const int64_t* idx, int64_t m, double* out, int64_t ldo, cudaStream_t s);
int kmg_ew_combine(const double* const* Ks, const int64_t* lds, const double* u, int p, int degree, int64_t rows, int64_t cols,
                   double* out, int64_t ldo, cudaStream_t s);
// result = sum_ij A_ij * (B ? B_ij : 1) * (w ? w_i w_j : 1);  partial: n doubles of scratch
int kmg_ew_weighted_dot(const double* A, int64_t lda, const double* B, int64_t ldb, const double* w, int64_t n, double* partial,
                        double* result, cudaStream_t s);
// s32 -> u16 (count % 1 == 0; src 16-byte, dst 16-byte aligned); *flag |= 1 when an entry does not fit
int kmg_ew_narrow_u16(const int32_t* src, int64_t count, uint16_t* dst, int* flag, cudaStream_t s);
int kmg_ew_row_sums(const double* K, int64_t rows, int64_t cols, int64_t ld, double* rs, cudaStream_t s);
int64_t kmg_ew_col_sums_workspace(int64_t rows, int64_t cols);
int kmg_ew_col_sums(const double* K, int64_t rows, int64_t cols, int64_t ld, double* cs, void* workspace, cudaStream_t s);
int kmg_ew_center_apply(const double* K, int64_t rows, int64_t cols, int64_t n_total, int64_t ld, const double* rs, const double* cs,
                        const double* g, double* out, int64_t ldo, cudaStream_t s);
// <Ac, Bc>_F with both symmetric factors centred on the fly from their row sums r and grand sums g (device scalars);
// rA == NULL / rB == NULL: that factor is taken as it is; B == NULL: sum of Ac
int kmg_ew_centered_dot(const double* A, int64_t lda, const double* rA, const double* gA, const double* B, int64_t ldb, const double* rB,
                        const double* gB, int64_t n, double* partial, double* result, cudaStream_t s);
// second stage of the epilogue row statistics (epi_ops.cuh): out[r] = sum_chunk partial[r][chunk], fixed order
int kmg_ew_partial_rows_reduce(const double* partial, int64_t rows, int64_t n_chunks, double* out, cudaStream_t s);
int kmg_ew_vec_sum(const double* v, int64_t n, double* out, cudaStream_t s);
int kmg_ew_vec_dot(const double* a, const double* b, int64_t n, double* out, cudaStream_t s);
int kmg_ew_row_wsums(const double* K, int64_t rows, int64_t cols, int64_t ld, const double* w, double* out, cudaStream_t s);
int kmg_ew_gather_vec(const double* v, const int64_t* idx, int64_t m, double* out, cudaStream_t s);
// out = [out +] u * normalised(K), power / normalisation of the last term: the stored-Gram form of the fused accumulate epilogue
int kmg_ew_accumulate(const double* K, int64_t ldk, const double* sd, const EpiOps* e, int64_t n, double* out, int64_t ldo, cudaStream_t s);
// solve.cu: (S K S + c I) x = b on a device-resident symmetric K (KRR.py:33, KLR.py:41-57); work: kmg_solve_workspace(n) bytes
int64_t kmg_solve_workspace(int64_t n);
int kmg_spd_solve_launch(const double* K, int64_t n, int64_t ld, const double* s, double c, const double* b, double* x, void* work,
                         int** d_flag, cudaStream_t st);
