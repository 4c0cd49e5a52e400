// gram_i8.h -- internal launch interface of the tcgen05 int8 Gram GEMM (gram_i8_tcgen05.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "epi_ops.cuh"

#define KMG_OUT_S32 0
#define KMG_OUT_F64 1
#define KMG_MAX_PARTS 8  // power of two

struct GramI8Args {
    const int8_t* phi_rows;  // first row of the row block of Phi (device)
    const int8_t* phi_cols;  // first row of the column block of Phi (device)
    int64_t rows, cols;      // block shape
    int64_t Dpad;            // feature width, multiple of 128 (zero padded)
    int64_t ld_phi;          // bytes between rows of Phi
    int64_t row_index0, col_index0;  // global indices of the block origin (diagonal detection)
    void* out;               // rows x cols, element stride ldo
    int64_t ldo;
    int out_dtype;           // KMG_OUT_S32 / KMG_OUT_F64
    int symmetric;           // 1: square diagonal block, compute tiles touching col >= row and mirror-store the rest
    void* out_t;             // mirror destination: element (r,c) -> out_t[c*ldo_t + r] (== out for an in-place full Gram)
    int64_t ldo_t;
    const double* sd_rows;   // optional cosine normalisation: sqrt(diag) for the block's rows / cols
    const double* sd_cols;
    int m_sub;               // 0 = auto, 1 = 128x256 tiles, 2 = 256x256 tiles
    int max_ctas;            // 0 = all SMs
    int64_t* computed_entries;  // optional out: entries actually issued to the tensor cores
    int mirror_all;          // 1: plain block, but every tile is also stored transposed to out_t[c*ldo_t + r] (block-local r, c)
    // Sharded symmetric build (n_parts > 0): the n x n Gram is cut into n_parts block-rows, one per GPU; this launch
    // computes part `part`'s share of the upper-triangle work and stores every tile twice -- into its own block-row
    // and, transposed, into the block-row of the part that owns the tile's columns (peer device memory).
    // rows = the part's row count, cols = n, row_index0 = part_row0[part], col_index0 = 0; phi_cols = all of Phi.
    int n_parts, part;
    const int64_t* part_row0;  // host: n_parts + 1 boundaries, multiples of 256 except the last (= n)
    void* const* part_out;     // host: device base pointer of every part's block-row buffer (row stride ldo)
    const EpiOps* epi;         // optional fused ALIGNF / NLCK steps (epi_ops.cuh): plain fp64 block, CTA-pair kernel
};

// 1 when part `a` of `g` (boundaries part_row0, multiples of 256) computes tile (I, J) of the global 256 x 256 tile grid,
// I in a's row tiles, J in the column tiles owned by part b; see gram_i8_tcgen05.cu
int kmg_gram_sharded_takes(int g, const int64_t* part_row0, int a, int b, int64_t I, int64_t J);

int kmg_gram_i8_launch(const GramI8Args* a, cudaStream_t stream);
// frees the cached tile lists of every device (kmg_release)
void kmg_gram_i8_clear_cache();
// mma_peak.cu: back-to-back tcgen05.mma.cta_group::2.kind::i8 on every CTA pair; *ops = int8 operations issued
int kmg_mma_peak_i8_launch(int iters, int64_t* ops, cudaStream_t stream);
int kmg_gram_i8_simt_launch(const int8_t* A, const int8_t* B, int64_t ld, int64_t rows, int64_t cols, int64_t Dpad,
                            int32_t* out, int64_t ldo, cudaStream_t stream);
// alu_peak.cu: issue-rate microbenchmarks of the CUDA-core pipes (kind: 0 LOP3, 1 SHF, 2 POPC, 3 DFMA, 4 DADD, 5 DMUL)
int kmg_alu_peak_launch(int kind, int iters, int64_t* ops, cudaStream_t stream);
