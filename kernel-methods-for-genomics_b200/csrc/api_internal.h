// api_internal.h -- helpers of api.cu shared with the other host-side translation units of libkmg.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "runtime.h"

// upload n x L sequence bytes (ASCII or codes) and pack them into bit-planes (n x 8 u32); KMG_ERR_ALPHABET on a bad byte
int kmg_api_upload_planes(const uint8_t* seqs, int64_t n, int L, int fmt, DevBuf* planes, cudaStream_t s);
