// elementwise.h -- internal launch interface of the stored-Gram passes (elementwise.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

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
