// pair_kernels.h -- internal launch interface of the pairwise (bit-vector) Gram kernels:
// (k,m)-mismatch, weighted degree (+ shifts) and local alignment.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "epi_ops.cuh"

struct PairBlock {
    const uint32_t* planes_rows;  // bit-planes of the block's row sequences (device, 8 words each)
    const uint32_t* planes_cols;
    int64_t rows, cols;
    int64_t row_index0, col_index0;  // global indices of the block origin
    int L;
    void* out;                       // rows x cols
    int64_t ldo;
    int out_dtype;                   // KMG_OUT_S32 / KMG_OUT_F64 (fp kernels: F64 only)
    int symmetric;                   // square diagonal block: skip tiles below the diagonal, mirror-store instead
    void* out_t;
    int64_t ldo_t;
    const double* sd_rows;           // optional cosine normalisation (sqrt of the raw diagonal)
    const double* sd_cols;
    const EpiOps* epi;               // optional fused ALIGNF / NLCK steps (epi_ops.cuh); fp64 output only; row partials need a
                                     // plain (non-symmetric) block whose column origin is a multiple of 32
};

#define KMG_MM_MAX_M 3
int kmg_mismatch_table(int k, int m, int64_t* T /* k+1 entries */);
int kmg_mismatch_launch(const PairBlock* b, int k, int m, cudaStream_t stream);
// sd[i] = sqrt(K_raw(x_i, x_i))
int kmg_mismatch_diag_launch(const uint32_t* planes, int64_t n, int L, int k, int m, double* sd, cudaStream_t stream);

int kmg_wd_launch(const PairBlock* b, int d, cudaStream_t stream);
int kmg_wds_launch(const PairBlock* b, int d, int S, cudaStream_t stream);

int kmg_la_launch(const PairBlock* b, double e, double d, double beta, int smith, cudaStream_t stream);
