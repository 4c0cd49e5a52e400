// seq_kernels.h -- internal launch interface of the encoding / feature-map kernels (seq_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define KMG_MAX_KS 8       // values of k summed in one spectrum Gram
#define KMG_MAX_DENSE_K 8  // 4^8 = 64 KiB feature row; larger k goes through the pairwise kernel

int kmg_pack_launch(const uint8_t* d_in, int is_ascii, int64_t n, int L, uint32_t* d_planes, int* d_err, cudaStream_t stream);
int64_t kmg_spectrum_width(const int* ks, int nk, int L);
int64_t kmg_spectrum_padded_width(const int* ks, int nk, int L);
int kmg_spectrum_phi_launch(const uint32_t* d_planes, int64_t n, int L, const int* ks, int nk, int8_t* d_phi, int64_t ld,
                            cudaStream_t stream);
int kmg_phi_diag_sqrt_launch(const int8_t* d_phi, int64_t n, int64_t width, int64_t ld, double* d_sd, cudaStream_t stream);
int kmg_mismatch_phi_launch(const uint32_t* d_planes, int64_t n, int L, int k, int m, int8_t* d_phi, int64_t ld, cudaStream_t stream);
