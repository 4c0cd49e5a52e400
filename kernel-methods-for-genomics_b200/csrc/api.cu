// api.cu -- the C-ABI of libkmg.so (include/kmg.h): argument checking and the orchestration of the entry points
// (runtime.cu: device buffers and streams; host_link.cu: delivery of Gram blocks to host memory).  No compute happens
// here and there is no CPU fallback: every path ends in one of the sm_100a kernels of this directory.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <sys/mman.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <algorithm>
#include <chrono>
#include <future>
#include <map>
#include <mutex>
#include <vector>

#include "../../include/kmg.h"
#include "api_internal.h"
#include "elementwise.h"
#include "gram_i8.h"
#include "kmg_common.cuh"
#include "pair_kernels.h"
#include "host_link.h"
#include "runtime.h"
#include "seq_kernels.h"

namespace {

// Upload + pack one set of sequences.  planes: n x 8 u32.
int upload_planes(const uint8_t* seqs, int64_t n, int L, int fmt, DevBuf* planes, cudaStream_t s) {
    KMG_REQUIRE(L >= 1 && L <= KMG_MAX_L, KMG_ERR_UNSUPPORTED, "sequence length %d not supported (1..%d)", L, KMG_MAX_L);
    int rc = planes->alloc((size_t)std::max<int64_t>(n, 1) * KMG_SEQ_WORDS * sizeof(uint32_t));
    if (rc) return rc;
    if (n == 0) return KMG_OK;
    DevBuf raw, err;
    if ((rc = raw.alloc((size_t)n * L))) return rc;
    if ((rc = err.alloc(sizeof(int)))) return rc;
    KMG_CUDA_CHECK(cudaMemsetAsync(err.p, 0, sizeof(int), s));
    if ((rc = kmg_hl_h2d(raw.p, seqs, (size_t)n * L, s))) return rc;
    if ((rc = kmg_pack_launch(raw.as<uint8_t>(), fmt == KMG_SEQ_ASCII, n, L, planes->as<uint32_t>(), err.as<int>(), s))) return rc;
    int herr = 0;
    KMG_CUDA_CHECK(cudaMemcpyAsync(&herr, err.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    KMG_CUDA_CHECK(cudaStreamSynchronize(s));
    KMG_REQUIRE(herr == 0, KMG_ERR_ALPHABET, "sequence contains a character outside {A,C,G,T}");
    return KMG_OK;
}

struct SeqPair {
    DevBuf prow, pcol;
    const uint32_t* rows = nullptr;
    const uint32_t* cols = nullptr;
    int64_t nr = 0, nc = 0;
    bool symmetric = false;
};

int upload_pair(const uint8_t* rows, int64_t nr, const uint8_t* cols, int64_t nc, int L, int fmt, SeqPair* sp) {
    KMG_REQUIRE(nr >= 0 && (cols == nullptr || nc >= 0), KMG_ERR_ARG, "negative sequence count");
    KMG_REQUIRE(rows != nullptr || nr == 0, KMG_ERR_ARG, "null sequence pointer");
    cudaStream_t s;
    int rc = kmg_rt_get_streams(&s, nullptr);
    if (rc) return rc;
    if ((rc = upload_planes(rows, nr, L, fmt, &sp->prow, s))) return rc;
    sp->rows = sp->prow.as<uint32_t>();
    sp->nr = nr;
    if (cols == nullptr) {
        sp->cols = sp->rows; sp->nc = nr; sp->symmetric = true;
    } else {
        if ((rc = upload_planes(cols, nc, L, fmt, &sp->pcol, s))) return rc;
        sp->cols = sp->pcol.as<uint32_t>(); sp->nc = nc; sp->symmetric = false;
    }
    return KMG_OK;
}

// ---- per-kernel block builders -----------------------------------------------------------
struct SpectrumCtx {
    const int8_t* phi_rows; const int8_t* phi_cols; int64_t nc; int64_t width;
    const double* sd_rows; const double* sd_cols;
    int out_dtype;
};
int spectrum_block(void* c, int64_t r0, int64_t rows, void* out, int64_t ldo, int symmetric, cudaStream_t s) {
    SpectrumCtx* x = (SpectrumCtx*)c;
    GramI8Args a;
    memset(&a, 0, sizeof(a));
    a.phi_rows = x->phi_rows + r0 * x->width; a.phi_cols = x->phi_cols;
    a.rows = rows; a.cols = x->nc; a.Dpad = x->width; a.ld_phi = x->width;
    a.row_index0 = symmetric ? 0 : r0; a.col_index0 = 0;
    a.out = out; a.ldo = ldo; a.out_dtype = x->out_dtype; a.symmetric = symmetric; a.out_t = out; a.ldo_t = ldo;
    a.sd_rows = x->sd_rows ? x->sd_rows + r0 : nullptr; a.sd_cols = x->sd_cols;
    return kmg_gram_i8_launch(&a, s);
}

struct PairCtx {
    const SeqPair* sp; int L; int kind; int k, m, d, smith; double e, dd, beta; const double* sd; bool index_diag;
};
int pair_block(void* c, int64_t r0, int64_t rows, void* out, int64_t ldo, int symmetric, cudaStream_t s) {
    PairCtx* x = (PairCtx*)c;
    PairBlock b;
    memset(&b, 0, sizeof(b));
    b.planes_rows = x->sp->rows + r0 * KMG_SEQ_WORDS; b.planes_cols = x->sp->cols;
    b.rows = rows; b.cols = x->sp->nc;
    // cross-Grams have no diagonal: offset the column indices so that row index never equals column index
    b.row_index0 = r0; b.col_index0 = x->index_diag ? 0 : (int64_t)1 << 40;
    b.L = x->L; b.out = out; b.ldo = ldo; b.out_dtype = KMG_OUT_F64;
    b.symmetric = symmetric; b.out_t = out; b.ldo_t = ldo;
    b.sd_rows = x->sd ? x->sd + r0 : nullptr; b.sd_cols = x->sd;
    if (x->kind == 0) return kmg_mismatch_launch(&b, x->k, x->m, s);
    if (x->kind == 1) return kmg_wd_launch(&b, x->d, s);
    if (x->kind == 3) return kmg_wds_launch(&b, x->d, x->k /* S */, s);
    return kmg_la_launch(&b, x->e, x->dd, x->beta, x->smith, s);
}

}  // namespace

int kmg_api_upload_planes(const uint8_t* seqs, int64_t n, int L, int fmt, DevBuf* planes, cudaStream_t s) {
    KMG_REQUIRE(n >= 0 && (seqs != nullptr || n == 0), KMG_ERR_ARG, "bad sequence buffer");
    return upload_planes(seqs, n, L, fmt, planes, s);
}

// ------------------------------------------------------------------------------------------
// library
// ------------------------------------------------------------------------------------------
extern "C" {

int kmg_version(void) { return 100; }
const char* kmg_last_error(void) { return kmg_rt_last_error(); }

int kmg_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int kmg_set_device(int device) {
    KMG_CUDA_CHECK(cudaSetDevice(device));
    return KMG_OK;
}

// ---- plain device buffers for callers that keep Grams resident between calls (kmg/resident.py) ----
int kmg_dev_malloc(int64_t bytes, void** ptr) {
    int rc = kmg_rt_require_device();
    if (rc) return rc;
    KMG_REQUIRE(bytes >= 0 && ptr != nullptr, KMG_ERR_ARG, "dev_malloc: bad arguments");
    *ptr = nullptr;
    if (bytes == 0) return KMG_OK;
    cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        kmg_rt_flush_cache();
        e = cudaMalloc(ptr, (size_t)bytes);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        kmg_set_error("cudaMalloc(%lld bytes) failed: %s", (long long)bytes, cudaGetErrorString(e));
        return KMG_ERR_NOMEM;
    }
    return KMG_OK;
}

int kmg_dev_free(void* ptr) {
    if (ptr) KMG_CUDA_CHECK(cudaFree(ptr));
    return KMG_OK;
}

int kmg_dev_upload(void* d_dst, const void* h_src, int64_t bytes) {
    if (bytes > 0) KMG_CUDA_CHECK(cudaMemcpy(d_dst, h_src, (size_t)bytes, cudaMemcpyHostToDevice));
    return KMG_OK;
}

int kmg_dev_download(void* h_dst, const void* d_src, int64_t bytes) {
    if (bytes > 0) KMG_CUDA_CHECK(cudaMemcpy(h_dst, d_src, (size_t)bytes, cudaMemcpyDeviceToHost));
    return KMG_OK;
}

// ---- recycled host memory for result arrays (kmg/host.py wraps a block as the numpy array it returns) ----
int kmg_host_alloc(int64_t bytes, void** ptr) { return kmg_hl_host_alloc(bytes, ptr); }
int kmg_host_free(void* ptr) { return kmg_hl_host_free(ptr); }

int kmg_set_d2h_mode(int mode) { return kmg_hl_set_mode(mode); }
int kmg_get_d2h_mode(void) { return kmg_hl_get_mode(); }

int kmg_release(void) {
    kmg_hl_trim_pool();
    kmg_rt_flush_cache();
    kmg_gram_i8_clear_cache();
    return KMG_OK;
}

int kmg_mismatch_table_host(int k, int m, int64_t* T) {
    KMG_REQUIRE(k >= 1 && k <= KMG_MAX_L && m >= 0 && T != nullptr, KMG_ERR_ARG, "mismatch_table: bad arguments");
    return kmg_mismatch_table(k, m, T);
}

// ------------------------------------------------------------------------------------------
// host-buffer entry points
// ------------------------------------------------------------------------------------------
int kmg_spectrum_host(const uint8_t* rows, int64_t nr, const uint8_t* cols, int64_t nc, int L, int seq_format,
                      const int* ks, int nk, double* K, int64_t ldk) {
    int rc = kmg_rt_require_device();
    if (rc) return rc;
    KMG_REQUIRE(ks != nullptr && nk >= 1 && nk <= KMG_MAX_KS, KMG_ERR_ARG, "spectrum: between 1 and %d values of k", KMG_MAX_KS);
    bool pairwise = false;
    for (int q = 0; q < nk; ++q) {
        KMG_REQUIRE(ks[q] >= 1, KMG_ERR_ARG, "spectrum: k must be >= 1");
        if (ks[q] > KMG_MAX_DENSE_K || ks[q] > L) pairwise = true;
    }
    if (pairwise) {
        // k too large for a dense 4^k feature row: SP(k) == raw MM(k, 0) (kernels.py:161-175 with m=0)
        KMG_REQUIRE(nk == 1, KMG_ERR_UNSUPPORTED, "spectrum: sums over several k need every k <= %d", KMG_MAX_DENSE_K);
        if (ks[0] > L) {  // no window at all: the reference returns zeros
            const int64_t c = cols ? nc : nr;
            for (int64_t i = 0; i < nr; ++i) memset(K + i * ldk, 0, (size_t)c * sizeof(double));
            return KMG_OK;
        }
        return kmg_mismatch_host(rows, nr, cols, nc, L, seq_format, ks[0], 0, 0, KMG_MM_PAIRWISE, K, ldk);
    }
    SeqPair sp;
    kmg_trace("spectrum_host: enter");
    if ((rc = upload_pair(rows, nr, cols, nc, L, seq_format, &sp))) return rc;
    kmg_trace("spectrum_host: sequences uploaded and packed");
    KMG_REQUIRE(K != nullptr && ldk >= sp.nc, KMG_ERR_ARG, "spectrum: bad output buffer");
    if (sp.nr == 0 || sp.nc == 0) return KMG_OK;
    cudaStream_t s;
    if ((rc = kmg_rt_get_streams(&s, nullptr))) return rc;
    const int64_t width = kmg_spectrum_padded_width(ks, nk, L);
    DevBuf phi_r, phi_c;
    if ((rc = phi_r.alloc((size_t)sp.nr * width))) return rc;
    if ((rc = kmg_spectrum_phi_launch(sp.rows, sp.nr, L, ks, nk, phi_r.as<int8_t>(), width, s))) return rc;
    const int8_t* pc = phi_r.as<int8_t>();
    if (!sp.symmetric) {
        if ((rc = phi_c.alloc((size_t)sp.nc * width))) return rc;
        if ((rc = kmg_spectrum_phi_launch(sp.cols, sp.nc, L, ks, nk, phi_c.as<int8_t>(), width, s))) return rc;
        pc = phi_c.as<int8_t>();
    }
    KMG_CUDA_CHECK(cudaStreamSynchronize(s));
    kmg_trace("spectrum_host: Phi built");
    // unnormalised counts: ship the s32 accumulators, widen to double on the host side of PCIe
    const bool s32 = (size_t)sp.nc * 4 <= kmg_hl_slot_bytes() && !getenv("KMG_D2H_F64") && !kmg_hl_direct_fp64(K, ldk, sp.nr);
    SpectrumCtx ctx{phi_r.as<int8_t>(), pc, sp.nc, width, nullptr, nullptr, s32 ? KMG_OUT_S32 : KMG_OUT_F64};
    return kmg_hl_build_to_host(sp.nr, sp.nc, sp.symmetric, spectrum_block, &ctx, K, ldk, s32);
}

int kmg_mismatch_host(const uint8_t* rows, int64_t nr, const uint8_t* cols, int64_t nc, int L, int seq_format,
                      int k, int m, int normalize, int algo, double* K, int64_t ldk) {
    int rc = kmg_rt_require_device();
    if (rc) return rc;
    KMG_REQUIRE(!(normalize && cols != nullptr), KMG_ERR_ARG, "mismatch: normalisation is defined for the symmetric Gram only");
    KMG_REQUIRE(algo >= KMG_MM_AUTO && algo <= KMG_MM_DENSE, KMG_ERR_ARG, "mismatch: algo must be 0 (auto), 1 (pairwise) or 2 (dense)");
    KMG_REQUIRE(k >= 1 && k <= L, KMG_ERR_ARG, "mismatch: need 1 <= k <= L (k=%d, L=%d)", k, L);
    SeqPair sp;
    if ((rc = upload_pair(rows, nr, cols, nc, L, seq_format, &sp))) return rc;
    KMG_REQUIRE(K != nullptr && ldk >= sp.nc, KMG_ERR_ARG, "mismatch: bad output buffer");
    if (sp.nr == 0 || sp.nc == 0) return KMG_OK;
    cudaStream_t s;
    if ((rc = kmg_rt_get_streams(&s, nullptr))) return rc;
    if (algo == KMG_MM_AUTO) algo = (k <= KMG_MAX_DENSE_K && m <= 3 && L - k + 1 <= 127) ? KMG_MM_DENSE : KMG_MM_PAIRWISE;
    if (algo == KMG_MM_DENSE) {
        // the reference's own structure (kernels.py:206-215): dense phi_km, then Phi Phi^T -- on the tensor cores
        const int64_t width = ((1ll << (2 * k)) + 127) / 128 * 128;
        DevBuf phi_r, phi_c, sd;
        if ((rc = phi_r.alloc((size_t)sp.nr * width))) return rc;
        if ((rc = kmg_mismatch_phi_launch(sp.rows, sp.nr, L, k, m, phi_r.as<int8_t>(), width, s))) return rc;
        const int8_t* pc = phi_r.as<int8_t>();
        if (!sp.symmetric) {
            if ((rc = phi_c.alloc((size_t)sp.nc * width))) return rc;
            if ((rc = kmg_mismatch_phi_launch(sp.cols, sp.nc, L, k, m, phi_c.as<int8_t>(), width, s))) return rc;
            pc = phi_c.as<int8_t>();
        }
        const double* sdp = nullptr;
        if (normalize) {
            if ((rc = sd.alloc((size_t)sp.nr * sizeof(double)))) return rc;
            if ((rc = kmg_phi_diag_sqrt_launch(phi_r.as<int8_t>(), sp.nr, width, width, sd.as<double>(), s))) return rc;
            double sd0 = 0.0;
            KMG_CUDA_CHECK(cudaMemcpyAsync(&sd0, sd.p, sizeof(double), cudaMemcpyDeviceToHost, s));
            KMG_CUDA_CHECK(cudaStreamSynchronize(s));
            if (sd0 != 1.0) sdp = sd.as<double>();  // normalize_K early-out, kernels.py:404
        }
        KMG_CUDA_CHECK(cudaStreamSynchronize(s));
        SpectrumCtx ctx{phi_r.as<int8_t>(), pc, sp.nc, width, sdp, sdp, KMG_OUT_F64};
        return kmg_hl_build_to_host(sp.nr, sp.nc, sp.symmetric, spectrum_block, &ctx, K, ldk);
    }
    DevBuf sd;
    const double* sdp = nullptr;
    if (normalize) {
        if ((rc = sd.alloc((size_t)sp.nr * sizeof(double)))) return rc;
        if ((rc = kmg_mismatch_diag_launch(sp.rows, sp.nr, L, k, m, sd.as<double>(), s))) return rc;
        // normalize_K early-out (kernels.py:404): a raw K[0,0] of exactly 1 leaves the matrix unnormalised
        double sd0 = 0.0;
        KMG_CUDA_CHECK(cudaMemcpyAsync(&sd0, sd.p, sizeof(double), cudaMemcpyDeviceToHost, s));
        KMG_CUDA_CHECK(cudaStreamSynchronize(s));
        if (sd0 != 1.0) sdp = sd.as<double>();
    }
    KMG_CUDA_CHECK(cudaStreamSynchronize(s));
    PairCtx ctx{&sp, L, 0, k, m, 0, 0, 0.0, 0.0, 0.0, sdp, sp.symmetric};
    return kmg_hl_build_to_host(sp.nr, sp.nc, sp.symmetric, pair_block, &ctx, K, ldk);
}

static int phi_host(const uint8_t* seqs, int64_t n, int L, int seq_format, int which, const int* ks, int nk, int k, int m,
                    int8_t* phi, int64_t ld) {
    int rc = kmg_rt_require_device();
    if (rc) return rc;
    KMG_REQUIRE(n >= 0 && (seqs != nullptr || n == 0) && (phi != nullptr || n == 0), KMG_ERR_ARG, "phi: bad arguments");
    cudaStream_t s;
    if ((rc = kmg_rt_get_streams(&s, nullptr))) return rc;
    DevBuf planes, dphi;
    if ((rc = upload_planes(seqs, n, L, seq_format, &planes, s))) return rc;
    const int64_t width = which == 0 ? kmg_spectrum_padded_width(ks, nk, L) : ((1ll << (2 * k)) + 127) / 128 * 128;
    KMG_REQUIRE(ld >= width, KMG_ERR_ARG, "phi: ld must be >= %lld", (long long)width);
    if (n == 0) return KMG_OK;
    if ((rc = dphi.alloc((size_t)n * width))) return rc;
    rc = which == 0 ? kmg_spectrum_phi_launch(planes.as<uint32_t>(), n, L, ks, nk, dphi.as<int8_t>(), width, s)
                    : kmg_mismatch_phi_launch(planes.as<uint32_t>(), n, L, k, m, dphi.as<int8_t>(), width, s);
    if (rc) return rc;
    KMG_CUDA_CHECK(cudaMemcpy2DAsync(phi, (size_t)ld, dphi.p, (size_t)width, (size_t)width, (size_t)n, cudaMemcpyDeviceToHost, s));
    KMG_CUDA_CHECK(cudaStreamSynchronize(s));
    return KMG_OK;
}

int kmg_spectrum_phi_host(const uint8_t* seqs, int64_t n, int L, int seq_format, const int* ks, int nk, int8_t* phi, int64_t ld) {
    KMG_REQUIRE(ks != nullptr && nk >= 1, KMG_ERR_ARG, "spectrum_phi: need at least one k");
    return phi_host(seqs, n, L, seq_format, 0, ks, nk, 0, 0, phi, ld);
}

int kmg_mismatch_phi_host(const uint8_t* seqs, int64_t n, int L, int seq_format, int k, int m, int8_t* phi, int64_t ld) {
    return phi_host(seqs, n, L, seq_format, 1, nullptr, 0, k, m, phi, ld);
}

int kmg_wd_host(const uint8_t* rows, int64_t nr, const uint8_t* cols, int64_t nc, int L, int seq_format,
                int d, double* K, int64_t ldk) {
    int rc = kmg_rt_require_device();
    if (rc) return rc;
    SeqPair sp;
    if ((rc = upload_pair(rows, nr, cols, nc, L, seq_format, &sp))) return rc;
    KMG_REQUIRE(K != nullptr && ldk >= sp.nc, KMG_ERR_ARG, "wd: bad output buffer");
    PairCtx ctx{&sp, L, 1, 0, 0, d, 0, 0.0, 0.0, 0.0, nullptr, sp.symmetric};
    return kmg_hl_build_to_host(sp.nr, sp.nc, sp.symmetric, pair_block, &ctx, K, ldk);
}

int kmg_wds_host(const uint8_t* rows, int64_t nr, const uint8_t* cols, int64_t nc, int L, int seq_format,
                 int d, int S, double* K, int64_t ldk) {
    int rc = kmg_rt_require_device();
    if (rc) return rc;
    SeqPair sp;
    if ((rc = upload_pair(rows, nr, cols, nc, L, seq_format, &sp))) return rc;
    KMG_REQUIRE(K != nullptr && ldk >= sp.nc, KMG_ERR_ARG, "wds: bad output buffer");
    PairCtx ctx{&sp, L, 3, /*k := S*/ S, 0, d, 0, 0.0, 0.0, 0.0, nullptr, sp.symmetric};
    return kmg_hl_build_to_host(sp.nr, sp.nc, sp.symmetric, pair_block, &ctx, K, ldk);
}

int kmg_la_host(const uint8_t* rows, int64_t nr, const uint8_t* cols, int64_t nc, int L, int seq_format,
                double e, double d, double beta, int smith, double* K, int64_t ldk) {
    int rc = kmg_rt_require_device();
    if (rc) return rc;
    SeqPair sp;
    if ((rc = upload_pair(rows, nr, cols, nc, L, seq_format, &sp))) return rc;
    KMG_REQUIRE(K != nullptr && ldk >= sp.nc, KMG_ERR_ARG, "la: bad output buffer");
    PairCtx ctx{&sp, L, 2, 0, 0, 0, smith, e, d, beta, nullptr, sp.symmetric};
    return kmg_hl_build_to_host(sp.nr, sp.nc, sp.symmetric, pair_block, &ctx, K, ldk);
}

int kmg_normalize_host(double* K, int64_t n, int64_t ldk) {
    int rc = kmg_rt_require_device();
    if (rc) return rc;
    KMG_REQUIRE(n >= 0 && (K != nullptr || n == 0) && ldk >= n, KMG_ERR_ARG, "normalize: bad arguments");
    if (n == 0) return 0;
    if (K[0] == 1.0) return 1;  // kernels.py:404-405
    cudaStream_t s;
    if ((rc = kmg_rt_get_streams(&s, nullptr))) return rc;
    DevBuf d, sd;
    if ((rc = d.alloc((size_t)n * n * 8))) return rc;
    if ((rc = sd.alloc((size_t)n * 8))) return rc;
    KMG_CUDA_CHECK(cudaMemcpy2DAsync(d.p, (size_t)n * 8, K, (size_t)ldk * 8, (size_t)n * 8, (size_t)n, cudaMemcpyHostToDevice, s));
    if ((rc = kmg_ew_diag_sqrt(d.as<double>(), n, n, sd.as<double>(), s))) return rc;
    if ((rc = kmg_ew_normalize(d.as<double>(), n, n, sd.as<double>(), s))) return rc;
    KMG_CUDA_CHECK(cudaMemcpy2DAsync(K, (size_t)ldk * 8, d.p, (size_t)n * 8, (size_t)n * 8, (size_t)n, cudaMemcpyDeviceToHost, s));
    KMG_CUDA_CHECK(cudaStreamSynchronize(s));
    return 0;
}

int kmg_center_host(const double* K, int64_t n, int64_t ldk, double* out, int64_t ldo) {
    int rc = kmg_rt_require_device();
    if (rc) return rc;
    KMG_REQUIRE(n >= 0 && ldk >= n && ldo >= n, KMG_ERR_ARG, "center: bad arguments");
    if (n == 0) return KMG_OK;
    cudaStream_t s;
    if ((rc = kmg_rt_get_streams(&s, nullptr))) return rc;
    DevBuf d, o, ws;
    if ((rc = d.alloc((size_t)n * n * 8))) return rc;
    if ((rc = o.alloc((size_t)n * n * 8))) return rc;
    if ((rc = ws.alloc((size_t)kmg_ew_center_workspace(n)))) return rc;
    KMG_CUDA_CHECK(cudaMemcpy2DAsync(d.p, (size_t)n * 8, K, (size_t)ldk * 8, (size_t)n * 8, (size_t)n, cudaMemcpyHostToDevice, s));
    if ((rc = kmg_ew_center(d.as<double>(), n, n, o.as<double>(), n, ws.p, s))) return rc;
    KMG_CUDA_CHECK(cudaMemcpy2DAsync(out, (size_t)ldo * 8, o.p, (size_t)n * 8, (size_t)n * 8, (size_t)n, cudaMemcpyDeviceToHost, s));
    KMG_CUDA_CHECK(cudaStreamSynchronize(s));
    return KMG_OK;
}

int kmg_combine_host(const double* const* Ks, int p, int64_t n, const double* u, int degree, int normalize, double* out) {
    int rc = kmg_rt_require_device();
    if (rc) return rc;
    KMG_REQUIRE(p >= 1 && p <= KMG_MAX_COMBINE && n >= 0 && Ks && u && out, KMG_ERR_ARG, "combine: bad arguments");
    if (n == 0) return KMG_OK;
    cudaStream_t s;
    if ((rc = kmg_rt_get_streams(&s, nullptr))) return rc;
    std::vector<DevBuf> bufs(p);
    const double* dptr[KMG_MAX_COMBINE];
    int64_t lds[KMG_MAX_COMBINE];
    for (int m = 0; m < p; ++m) {
        if ((rc = bufs[m].alloc((size_t)n * n * 8))) return rc;
        KMG_CUDA_CHECK(cudaMemcpyAsync(bufs[m].p, Ks[m], (size_t)n * n * 8, cudaMemcpyHostToDevice, s));
        dptr[m] = bufs[m].as<double>();
        lds[m] = n;
    }
    DevBuf o, sd;
    if ((rc = o.alloc((size_t)n * n * 8))) return rc;
    if ((rc = kmg_ew_combine(dptr, lds, u, p, degree, n, n, o.as<double>(), n, s))) return rc;
    if (normalize) {
        double k00 = 0.0;
        KMG_CUDA_CHECK(cudaMemcpyAsync(&k00, o.p, 8, cudaMemcpyDeviceToHost, s));
        KMG_CUDA_CHECK(cudaStreamSynchronize(s));
        if (k00 != 1.0) {
            if ((rc = sd.alloc((size_t)n * 8))) return rc;
            if ((rc = kmg_ew_diag_sqrt(o.as<double>(), n, n, sd.as<double>(), s))) return rc;
            if ((rc = kmg_ew_normalize(o.as<double>(), n, n, sd.as<double>(), s))) return rc;
        }
    }
    KMG_CUDA_CHECK(cudaMemcpyAsync(out, o.p, (size_t)n * n * 8, cudaMemcpyDeviceToHost, s));
    KMG_CUDA_CHECK(cudaStreamSynchronize(s));
    return KMG_OK;
}

int kmg_alignf_stats_host(const double* const* Ks, int p, int64_t n, const int64_t* idx, int64_t nfit, const double* y,
                          double* a, double* M) {
    int rc = kmg_rt_require_device();
    if (rc) return rc;
    KMG_REQUIRE(p >= 1 && p <= KMG_MAX_COMBINE && n >= 0 && nfit >= 0 && Ks && idx && y && a && M, KMG_ERR_ARG, "alignf_stats: bad arguments");
    for (int64_t t = 0; t < nfit; ++t) KMG_REQUIRE(idx[t] >= 0 && idx[t] < n, KMG_ERR_ARG, "alignf_stats: index out of range");
    if (nfit == 0) { for (int i = 0; i < p; ++i) { a[i] = 0; for (int j = 0; j < p; ++j) M[i * p + j] = 0; } return KMG_OK; }
    cudaStream_t s;
    if ((rc = kmg_rt_get_streams(&s, nullptr))) return rc;
    DevBuf sub, dy, ws, part, res;
    std::vector<DevBuf> kc(p);
    if ((rc = sub.alloc((size_t)nfit * nfit * 8))) return rc;
    if ((rc = dy.alloc((size_t)nfit * 8))) return rc;
    if ((rc = ws.alloc((size_t)kmg_ew_center_workspace(nfit)))) return rc;
    if ((rc = part.alloc((size_t)nfit * 8))) return rc;
    if ((rc = res.alloc((size_t)(p + p * p) * 8))) return rc;
    KMG_CUDA_CHECK(cudaMemcpyAsync(dy.p, y, (size_t)nfit * 8, cudaMemcpyHostToDevice, s));
    // The sub-block K[idx][:, idx] (ALIGNF.py:28) is gathered on the host: nfit^2 doubles cross PCIe per kernel instead
    // of the full n^2 matrix (run.py's sizes: 18 MB instead of 72 MB per kernel and data set).
    std::vector<double> hsub((size_t)nfit * nfit);
    for (int i = 0; i < p; ++i) {
        if ((rc = kc[i].alloc((size_t)nfit * nfit * 8))) return rc;
        const double* K = Ks[i];
        for (int64_t a = 0; a < nfit; ++a) {
            const double* row = K + idx[a] * n;
            double* dst = hsub.data() + a * nfit;
            for (int64_t b = 0; b < nfit; ++b) dst[b] = row[idx[b]];
        }
        KMG_CUDA_CHECK(cudaMemcpyAsync(sub.p, hsub.data(), (size_t)nfit * nfit * 8, cudaMemcpyHostToDevice, s));
        if ((rc = kmg_ew_center(sub.as<double>(), nfit, nfit, kc[i].as<double>(), nfit, ws.p, s))) return rc;              // ALIGNF.py:36-41
        // a_i = sum(Kc_i * y y')  (ALIGNF.py:43-48)
        if ((rc = kmg_ew_weighted_dot(kc[i].as<double>(), nfit, nullptr, 0, dy.as<double>(), nfit, part.as<double>(), res.as<double>() + i, s))) return rc;
        KMG_CUDA_CHECK(cudaStreamSynchronize(s));  // `hsub` / `sub` are reused by the next kernel
    }
    for (int i = 0; i < p; ++i)
        for (int j = i; j < p; ++j)  // M_ij = sum(Kc_i * Kc_j)  (ALIGNF.py:50-58)
            if ((rc = kmg_ew_weighted_dot(kc[i].as<double>(), nfit, kc[j].as<double>(), nfit, nullptr, nfit, part.as<double>(),
                                          res.as<double>() + p + i * p + j, s))) return rc;
    std::vector<double> h((size_t)(p + p * p), 0.0);
    KMG_CUDA_CHECK(cudaMemcpyAsync(h.data(), res.p, h.size() * 8, cudaMemcpyDeviceToHost, s));
    KMG_CUDA_CHECK(cudaStreamSynchronize(s));
    for (int i = 0; i < p; ++i) a[i] = h[i];
    for (int i = 0; i < p; ++i)
        for (int j = i; j < p; ++j) M[i * p + j] = M[j * p + i] = h[p + i * p + j];
    return KMG_OK;
}

int kmg_nlck_grad_host(const double* const* Ks_fit, int p, int64_t nfit, const double* u, const double* alpha, int degree,
                       double* grad) {
    int rc = kmg_rt_require_device();
    if (rc) return rc;
    KMG_REQUIRE(p >= 1 && p <= KMG_MAX_COMBINE && nfit >= 0 && Ks_fit && u && alpha && grad && degree >= 1, KMG_ERR_ARG, "nlck_grad: bad arguments");
    if (nfit == 0) { for (int m = 0; m < p; ++m) grad[m] = 0.0; return KMG_OK; }
    cudaStream_t s;
    if ((rc = kmg_rt_get_streams(&s, nullptr))) return rc;
    std::vector<DevBuf> bufs(p);
    const double* dptr[KMG_MAX_COMBINE];
    int64_t lds[KMG_MAX_COMBINE];
    for (int m = 0; m < p; ++m) {
        if ((rc = bufs[m].alloc((size_t)nfit * nfit * 8))) return rc;
        KMG_CUDA_CHECK(cudaMemcpyAsync(bufs[m].p, Ks_fit[m], (size_t)nfit * nfit * 8, cudaMemcpyHostToDevice, s));
        dptr[m] = bufs[m].as<double>();
        lds[m] = nfit;
    }
    DevBuf kt, da, part, res;
    if ((rc = kt.alloc((size_t)nfit * nfit * 8))) return rc;
    if ((rc = da.alloc((size_t)nfit * 8))) return rc;
    if ((rc = part.alloc((size_t)nfit * 8))) return rc;
    if ((rc = res.alloc((size_t)p * 8))) return rc;
    KMG_CUDA_CHECK(cudaMemcpyAsync(da.p, alpha, (size_t)nfit * 8, cudaMemcpyHostToDevice, s));
    // K_t = (sum_m u_m K_m) ** (degree - 1)   (NLCKernels.py:62)
    if ((rc = kmg_ew_combine(dptr, lds, u, p, degree - 1, nfit, nfit, kt.as<double>(), nfit, s))) return rc;
    for (int m = 0; m < p; ++m)  // alpha' (K_t * K_m) alpha  (NLCKernels.py:65)
        if ((rc = kmg_ew_weighted_dot(kt.as<double>(), nfit, dptr[m], nfit, da.as<double>(), nfit, part.as<double>(), res.as<double>() + m, s))) return rc;
    std::vector<double> h((size_t)p);
    KMG_CUDA_CHECK(cudaMemcpyAsync(h.data(), res.p, (size_t)p * 8, cudaMemcpyDeviceToHost, s));
    KMG_CUDA_CHECK(cudaStreamSynchronize(s));
    for (int m = 0; m < p; ++m) grad[m] = -(double)degree * h[m];  // NLCKernels.py:66
    return KMG_OK;
}

// ------------------------------------------------------------------------------------------
// device-pointer entry points
// ------------------------------------------------------------------------------------------
int kmg_pack_dev(const uint8_t* d_seqs, int seq_format, int64_t n, int L, uint32_t* d_planes, int* d_err_flag, void* stream) {
    return kmg_pack_launch(d_seqs, seq_format == KMG_SEQ_ASCII, n, L, d_planes, d_err_flag, (cudaStream_t)stream);
}

int64_t kmg_spectrum_phi_width(const int* ks, int nk) { return kmg_spectrum_padded_width(ks, nk, 0); }

int kmg_spectrum_phi_dev(const uint32_t* d_planes, int64_t n, int L, const int* ks, int nk, int8_t* d_phi, int64_t ld_phi,
                         void* stream) {
    return kmg_spectrum_phi_launch(d_planes, n, L, ks, nk, d_phi, ld_phi, (cudaStream_t)stream);
}

int kmg_gram_i8_dev(const int8_t* d_phi_rows, const int8_t* d_phi_cols, int64_t rows, int64_t cols, int64_t width,
                    int64_t ld_phi, int64_t row_index0, int64_t col_index0, void* d_out, int64_t ldo, int out_dtype,
                    int symmetric, const double* d_sd_rows, const double* d_sd_cols, int m_sub, void* stream) {
    KMG_REQUIRE(out_dtype == KMG_OUT_S32 || out_dtype == KMG_OUT_F64, KMG_ERR_ARG, "gram_i8: bad out_dtype");
    KMG_REQUIRE(!(d_sd_rows && out_dtype != KMG_OUT_F64), KMG_ERR_ARG, "gram_i8: normalisation needs the f64 output");
    KMG_REQUIRE((d_sd_rows == nullptr) == (d_sd_cols == nullptr), KMG_ERR_ARG, "gram_i8: sd_rows and sd_cols go together");
    if (symmetric)
        KMG_REQUIRE(rows == cols && row_index0 == col_index0 && d_phi_rows == d_phi_cols, KMG_ERR_ARG,
                    "gram_i8: symmetric mode needs a square diagonal block of one Phi");
    if (rows == 0 || cols == 0) return KMG_OK;
    GramI8Args a;
    memset(&a, 0, sizeof(a));
    a.phi_rows = d_phi_rows; a.phi_cols = d_phi_cols; a.rows = rows; a.cols = cols; a.Dpad = width; a.ld_phi = ld_phi;
    a.row_index0 = row_index0; a.col_index0 = col_index0; a.out = d_out; a.ldo = ldo; a.out_dtype = out_dtype;
    a.symmetric = symmetric; a.out_t = d_out; a.ldo_t = ldo; a.sd_rows = d_sd_rows; a.sd_cols = d_sd_cols; a.m_sub = m_sub;
    return kmg_gram_i8_launch(&a, (cudaStream_t)stream);
}

// Sharded symmetric spectrum Gram: part `part` of `n_parts` computes its share of the upper-triangle work
// and delivers every tile twice -- into its own block-row and, transposed, into the block-row of the part that owns the
// tile's columns: the mirror of kernels.py:45 is the one exchange step of the path.
//   KMG_EXCH_SINGLE : one launch; the epilogue's threads store the transposed tiles straight into the owners' buffers
//                     (peer memory).  Right for buffers on one device; over NVLink the thread-issued 256-byte stores cap
//                     the kernel (2 GPUs, n = 100 000: 40.5 ms against 30.1 ms with both buffers local).
//   KMG_EXCH_STAGED : one launch per peer block (cyclic distance 1, 2, ...) whose epilogue writes the transposed block
//                     contiguously into local staging (d_stage), each followed by ONE pitched peer copy on the copy
//                     stream while the next block's GEMM runs; the diagonal block (local mirror) goes last.
//   KMG_EXCH_DIRECT : one launch per peer block whose epilogue hands every transposed 32 x 16 piece to the TMA engine
//                     (cp.async.bulk.tensor store through a tensor map over the OWNER's block-row): compute and exchange
//                     are one kernel, tile by tile, with no staging pass and no copy-engine pass through HBM.
namespace {
struct SubBlock { int b; int64_t r_lo, r_hi, c_lo, c_hi; };
// The peer blocks part `a` computes, in launch order: the half block at distance g/2 first, then the full blocks at
// cyclic distance 1, 2, ... < g/2, each cut into KMG_SHARD_SPLIT (default 2) pieces BY COLUMNS.  Every piece is followed
// by the copy (staged exchange) of its transposed image to the owner while the next piece is computed, so what bounds the
// step is  (first launch) + (all copies) -- the copy engine is the critical path (measured per-piece timeline, 8 GPUs,
// n = 200 000, profiles/r2_shard_timeline.txt: a 5 GB block copies in 7.6 ms = 655 GB/s while its launch takes 7.4-8.0 ms,
// and with whole blocks the last copy ended 3.5-4.6 ms after the last launch).  Hence:
//   * pieces are cut by columns: the image of a piece is (columns of the piece) x (rows of the block), so its copy keeps
//     the full 200 KB row width; cutting by rows halves the width and the 2-D copy drops to 430 GB/s (measured: 43.6 ms per
//     step instead of 38.4);
//   * the half block goes first: one of its two sides always has a narrow image (half the rows), its slow copy runs under
//     the following launches instead of at the end;
//   * the last copy is one piece of a full block (2.5 GB) under the diagonal launch (4 ms), which has no exchange.
// At every phase rank a sends to (a + d) mod g -- a permutation, one incoming stream per receiver while the ranks stay in step.
void sharded_plan(int g, const int64_t* bounds, int a, std::vector<SubBlock>* out) {
    static const int split_env = getenv("KMG_SHARD_SPLIT") ? atoi(getenv("KMG_SHARD_SPLIT")) : 2;
    const int split = split_env < 1 ? 1 : (split_env > 8 ? 8 : split_env);
    const int64_t a0 = bounds[a], a1 = bounds[a + 1];
    if (g % 2 == 0 && g > 1) {
        // distance g/2: split by the lower-numbered part's row tiles, as kmg_gram_sharded_takes does
        const int b = (a + g / 2) % g;
        const int64_t b0 = bounds[b], b1 = bounds[b + 1];
        const int lo = a < b ? a : b;
        const int64_t lon = (bounds[lo + 1] - bounds[lo] + 255) / 256;
        const int64_t cut = std::min<int64_t>(bounds[lo] + (lon + 1) / 2 * 256, bounds[lo + 1]);
        if (a < b) { if (cut > a0) out->push_back({b, a0, cut, b0, b1}); }
        else if (cut < b1) out->push_back({b, a0, a1, cut, b1});
    }
    for (int d = 1; d < g; ++d) {
        if (2 * d >= g) break;
        const int b = (a + d) % g;
        const int64_t b0 = bounds[b], b1 = bounds[b + 1];
        const int64_t tiles = (b1 - b0 + 255) / 256;
        int64_t lo = b0;
        for (int q = 1; q <= split && lo < b1; ++q) {
            const int64_t hi = q == split ? b1 : std::min<int64_t>(b1, b0 + (tiles * q / split) * 256);
            if (hi > lo) out->push_back({b, a0, a1, lo, hi});
            lo = hi;
        }
    }
}
size_t stage_align(size_t x) { return (x + 255) & ~(size_t)255; }
}  // namespace

int kmg_gram_sharded_stage_bytes(int n_parts, const int64_t* part_row0, int part, int out_dtype, int64_t* bytes) {
    KMG_REQUIRE(n_parts >= 1 && n_parts <= KMG_MAX_PARTS && part_row0 && part >= 0 && part < n_parts && bytes, KMG_ERR_ARG,
                "gram_sharded_stage_bytes: bad arguments");
    std::vector<SubBlock> plan;
    sharded_plan(n_parts, part_row0, part, &plan);
    size_t total = 0;
    for (const SubBlock& sb : plan)
        total += stage_align((size_t)(sb.r_hi - sb.r_lo) * (size_t)(sb.c_hi - sb.c_lo) * (out_dtype == KMG_OUT_F64 ? 8 : 4));
    *bytes = (int64_t)total;
    return KMG_OK;
}

int kmg_gram_sharded_launches(int n_parts, const int64_t* part_row0, int part, int exchange, int* launches) {
    KMG_REQUIRE(n_parts >= 1 && n_parts <= KMG_MAX_PARTS && part_row0 && part >= 0 && part < n_parts && launches, KMG_ERR_ARG,
                "gram_sharded_launches: bad arguments");
    if (exchange == KMG_EXCH_SINGLE) { *launches = 1; return KMG_OK; }
    std::vector<SubBlock> plan;
    sharded_plan(n_parts, part_row0, part, &plan);
    *launches = (int)plan.size() + 1;  // + the diagonal block
    return KMG_OK;
}

int kmg_gram_i8_sharded_dev(const int8_t* d_phi, int64_t n, int64_t width, int64_t ld_phi, int n_parts, int part,
                            const int64_t* part_row0, void* const* part_out, int64_t ldo, int out_dtype, const double* d_sd,
                            int exchange, void* d_stage, int64_t* computed_entries, void* stream) {
    KMG_REQUIRE(out_dtype == KMG_OUT_S32 || out_dtype == KMG_OUT_F64, KMG_ERR_ARG, "gram_i8_sharded: bad out_dtype");
    KMG_REQUIRE(!(d_sd && out_dtype != KMG_OUT_F64), KMG_ERR_ARG, "gram_i8_sharded: normalisation needs the f64 output");
    KMG_REQUIRE(n_parts >= 1 && n_parts <= KMG_MAX_PARTS && part >= 0 && part < n_parts && part_row0 && part_out, KMG_ERR_ARG,
                "gram_i8_sharded: 1..%d parts", KMG_MAX_PARTS);
    KMG_REQUIRE(ldo >= n, KMG_ERR_ARG, "gram_i8_sharded: ldo < n");
    const bool defer_join = (exchange & KMG_EXCH_DEFER_JOIN) != 0;
    exchange &= ~KMG_EXCH_DEFER_JOIN;
    KMG_REQUIRE(exchange == KMG_EXCH_SINGLE || exchange == KMG_EXCH_DIRECT || (exchange == KMG_EXCH_STAGED && d_stage != nullptr), KMG_ERR_ARG,
                "gram_i8_sharded: exchange must be KMG_EXCH_SINGLE, KMG_EXCH_STAGED (with d_stage) or KMG_EXCH_DIRECT");
    if (computed_entries) *computed_entries = 0;
    if (n == 0) return KMG_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t a0 = part_row0[part], a1 = part_row0[part + 1];
    const int64_t esz = out_dtype == KMG_OUT_F64 ? 8 : 4;
    GramI8Args a;
    const bool staged = exchange == KMG_EXCH_STAGED;
    if (exchange == KMG_EXCH_SINGLE) {
        memset(&a, 0, sizeof(a));
        a.phi_rows = d_phi + a0 * ld_phi; a.phi_cols = d_phi; a.rows = a1 - a0; a.cols = n; a.Dpad = width; a.ld_phi = ld_phi;
        a.row_index0 = a0; a.col_index0 = 0; a.out = part_out[part]; a.ldo = ldo; a.out_dtype = out_dtype;
        a.sd_rows = d_sd ? d_sd + a0 : nullptr; a.sd_cols = d_sd;
        a.n_parts = n_parts; a.part = part; a.part_row0 = part_row0; a.part_out = part_out; a.computed_entries = computed_entries;
        return kmg_gram_i8_launch(&a, s);
    }
    for (int q = 0; q <= n_parts; ++q)
        KMG_REQUIRE((q == n_parts ? part_row0[q] == n : part_row0[q] % 256 == 0) && (q == 0 ? part_row0[0] == 0 : part_row0[q] > part_row0[q - 1]),
                    KMG_ERR_ARG, "gram_i8_sharded: part boundaries must start at 0, increase in multiples of 256 and end at n");
    int rc;
    cudaStream_t s0, copy;
    if ((rc = kmg_rt_get_streams(&s0, &copy))) return rc;
    // KMG_SHARD_COPY_STREAMS=2: consecutive pieces alternate between two copy streams (two copy engines at work)
    static const int n_copy = getenv("KMG_SHARD_COPY_STREAMS") ? atoi(getenv("KMG_SHARD_COPY_STREAMS")) : 1;
    static cudaStream_t copy2[64] = {};
    cudaStream_t copies[2] = {copy, copy};
    if (n_copy >= 2) {
        int dev = 0;
        KMG_CUDA_CHECK(cudaGetDevice(&dev));
        if (copy2[dev & 63] == nullptr) KMG_CUDA_CHECK(cudaStreamCreateWithFlags(&copy2[dev & 63], cudaStreamNonBlocking));
        copies[1] = copy2[dev & 63];
    }
    int piece = 0;
    std::vector<SubBlock> plan;
    sharded_plan(n_parts, part_row0, part, &plan);
    // (Round 1, 8 GPUs, n = 200 000, whole blocks with the half block first: 38.5 ms/step, 35.3 ms of it the GEMM phase
    // including ~2.7 ms of exposed final copy; feeding two peers at once through two copy streams broke the one-stream-
    // per-receiver pattern and cost 7 ms; two copy engines on the same block changed nothing.)
    char* my = static_cast<char*>(part_out[part]);
    char* stage = static_cast<char*>(d_stage);
    int64_t total = 0, got = 0;
    cudaEvent_t ev;
    // KMG_SHARD_TRACE=1 (debug): time every launch and every copy of this call with events and print the timeline
    static const bool trace = getenv("KMG_SHARD_TRACE") != nullptr;
    std::vector<cudaEvent_t> tg, tc;
    auto mark = [&](std::vector<cudaEvent_t>& v, cudaStream_t st) {
        if (!trace) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        v.push_back(e);
    };
    for (const SubBlock& sb : plan) {
        const int64_t rows = sb.r_hi - sb.r_lo, cols = sb.c_hi - sb.c_lo;
        memset(&a, 0, sizeof(a));
        a.phi_rows = d_phi + sb.r_lo * ld_phi; a.phi_cols = d_phi + sb.c_lo * ld_phi; a.rows = rows; a.cols = cols; a.Dpad = width; a.ld_phi = ld_phi;
        a.row_index0 = sb.r_lo; a.col_index0 = sb.c_lo; a.out = my + ((sb.r_lo - a0) * ldo + sb.c_lo) * esz; a.ldo = ldo; a.out_dtype = out_dtype;
        a.sd_rows = d_sd ? d_sd + sb.r_lo : nullptr; a.sd_cols = d_sd ? d_sd + sb.c_lo : nullptr;
        a.mirror_all = 1; a.computed_entries = &got;
        // the transposed block (cols x rows) belongs to rows [c_lo, c_hi) x columns [r_lo, r_hi) of the owner's block-row
        char* dst = static_cast<char*>(part_out[sb.b]) + ((sb.c_lo - part_row0[sb.b]) * ldo + sb.r_lo) * esz;
        if (staged) { a.out_t = stage; a.ldo_t = rows; }
        else { a.out_t = dst; a.ldo_t = ldo; }   // direct: the epilogue's TMA stores go to the owner's memory
        mark(tg, s);
        if ((rc = kmg_gram_i8_launch(&a, s))) return rc;
        mark(tg, s);
        total += got;
        if (!staged) continue;
        // the transposed block (cols x rows, contiguous) -> rows [c_lo, c_hi) x columns [r_lo, r_hi) of the owner's block-row
        cudaStream_t cs = copies[piece++ & 1];
        KMG_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        KMG_CUDA_CHECK(cudaEventRecord(ev, s));
        KMG_CUDA_CHECK(cudaStreamWaitEvent(cs, ev, 0));
        KMG_CUDA_CHECK(cudaEventDestroy(ev));
        mark(tc, cs);
        KMG_CUDA_CHECK(cudaMemcpy2DAsync(dst, (size_t)(ldo * esz), stage, (size_t)(rows * esz), (size_t)(rows * esz), (size_t)cols,
                                         cudaMemcpyDefault, cs));
        mark(tc, cs);
        stage += stage_align((size_t)(rows * cols * esz));
    }
    // diagonal block: the single-GPU symmetric build on this part's own square
    memset(&a, 0, sizeof(a));
    a.phi_rows = d_phi + a0 * ld_phi; a.phi_cols = a.phi_rows; a.rows = a1 - a0; a.cols = a1 - a0; a.Dpad = width; a.ld_phi = ld_phi;
    a.row_index0 = a0; a.col_index0 = a0; a.out = my + a0 * esz; a.ldo = ldo; a.out_dtype = out_dtype;
    a.symmetric = 1; a.out_t = a.out; a.ldo_t = ldo;
    a.sd_rows = d_sd ? d_sd + a0 : nullptr; a.sd_cols = a.sd_rows; a.computed_entries = &got;
    mark(tg, s);
    if ((rc = kmg_gram_i8_launch(&a, s))) return rc;
    mark(tg, s);
    total += got;
    if (trace) {
        cudaStreamSynchronize(s);
        cudaStreamSynchronize(copies[0]);
        cudaStreamSynchronize(copies[1]);
        char line[2048];
        int o = snprintf(line, sizeof(line), "[shard trace part %d] gemm:", part);
        for (size_t i = 0; i + 1 < tg.size(); i += 2) {
            float t0 = 0, t1 = 0;
            cudaEventElapsedTime(&t0, tg[0], tg[i]);
            cudaEventElapsedTime(&t1, tg[0], tg[i + 1]);
            o += snprintf(line + o, sizeof(line) - o, " [%.2f-%.2f]", t0, t1);
        }
        o += snprintf(line + o, sizeof(line) - o, " copy:");
        for (size_t i = 0; i + 1 < tc.size(); i += 2) {
            float t0 = 0, t1 = 0;
            cudaEventElapsedTime(&t0, tg[0], tc[i]);
            cudaEventElapsedTime(&t1, tg[0], tc[i + 1]);
            o += snprintf(line + o, sizeof(line) - o, " [%.2f-%.2f]", t0, t1);
        }
        fprintf(stderr, "%s\n", line);
        for (cudaEvent_t e : tg) cudaEventDestroy(e);
        for (cudaEvent_t e : tc) cudaEventDestroy(e);
    }
    if (staged && !defer_join) {  // `stream` completes only after the peer copies have
        for (int q = 0; q < (copies[1] != copies[0] ? 2 : 1); ++q) {
            KMG_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            KMG_CUDA_CHECK(cudaEventRecord(ev, copies[q]));
            KMG_CUDA_CHECK(cudaStreamWaitEvent(s, ev, 0));
            KMG_CUDA_CHECK(cudaEventDestroy(ev));
        }
    }
    if (computed_entries) *computed_entries = total;
    return KMG_OK;
}

// Makes `stream` wait for the peer copies this thread's staged exchanges have enqueued (KMG_EXCH_DEFER_JOIN): work launched on
// `stream` between the sharded build and the join -- e.g. the plain cross-Gram of the remaining columns -- runs under the copies.
int kmg_gram_sharded_join(void* stream) {
    int rc = kmg_rt_require_device();
    if (rc) return rc;
    cudaStream_t s0, copy;
    if ((rc = kmg_rt_get_streams(&s0, &copy))) return rc;
    cudaEvent_t ev;
    KMG_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    KMG_CUDA_CHECK(cudaEventRecord(ev, copy));
    KMG_CUDA_CHECK(cudaStreamWaitEvent((cudaStream_t)stream, ev, 0));
    KMG_CUDA_CHECK(cudaEventDestroy(ev));
    return KMG_OK;
}

// Marks on the copy stream: kmg_gram_sharded_mark(slot) records "every peer copy enqueued so far" in one of four per-thread
// events; kmg_gram_sharded_wait_mark(slot, stream) makes `stream` wait for that point only -- not for copies enqueued later.
// With two sets of block-row buffers this lets build k+1 start while the copies of build k are still draining, and build
// k+2 (same buffers as k) wait for exactly the copies of build k.
namespace {
thread_local cudaEvent_t g_marks[4] = {nullptr, nullptr, nullptr, nullptr};
thread_local int g_marks_dev = -1;
}
int kmg_gram_sharded_mark(int slot) {
    int rc = kmg_rt_require_device();
    if (rc) return rc;
    KMG_REQUIRE(slot >= 0 && slot < 4, KMG_ERR_ARG, "sharded_mark: slot 0..3");
    cudaStream_t s0, copy;
    if ((rc = kmg_rt_get_streams(&s0, &copy))) return rc;
    int dev = 0;
    KMG_CUDA_CHECK(cudaGetDevice(&dev));
    if (g_marks_dev != dev) { for (int q = 0; q < 4; ++q) g_marks[q] = nullptr; g_marks_dev = dev; }
    if (g_marks[slot] == nullptr) KMG_CUDA_CHECK(cudaEventCreateWithFlags(&g_marks[slot], cudaEventDisableTiming));
    KMG_CUDA_CHECK(cudaEventRecord(g_marks[slot], copy));
    return KMG_OK;
}
int kmg_gram_sharded_wait_mark(int slot, void* stream) {
    KMG_REQUIRE(slot >= 0 && slot < 4, KMG_ERR_ARG, "sharded_wait_mark: slot 0..3");
    if (g_marks[slot] == nullptr) return KMG_OK;  // never marked: nothing to wait for
    KMG_CUDA_CHECK(cudaStreamWaitEvent((cudaStream_t)stream, g_marks[slot], 0));
    return KMG_OK;
}

// Measured int8 tensor-core peak (mma_peak.cu): enqueue `iters` x 4 back-to-back MMAs per CTA pair; the caller times it.
int kmg_mma_peak_i8_dev(int iters, int64_t* ops, void* stream) {
    int rc = kmg_rt_require_device();
    if (rc) return rc;
    return kmg_mma_peak_i8_launch(iters, ops, (cudaStream_t)stream);
}

// Measured issue peak of a CUDA-core pipe (alu_peak.cu): enqueue `iters` x 64 dependent-chain instructions per thread.
int kmg_alu_peak_dev(int kind, int iters, int64_t* ops, void* stream) {
    int rc = kmg_rt_require_device();
    if (rc) return rc;
    return kmg_alu_peak_launch(kind, iters, ops, (cudaStream_t)stream);
}

int kmg_gram_sharded_takes_host(int n_parts, const int64_t* part_row0, int a, int b, int64_t I, int64_t J) {
    KMG_REQUIRE(n_parts >= 1 && n_parts <= KMG_MAX_PARTS && part_row0 && a >= 0 && a < n_parts && b >= 0 && b < n_parts, KMG_ERR_ARG,
                "gram_sharded_takes: bad arguments");
    return kmg_gram_sharded_takes(n_parts, part_row0, a, b, I, J);
}

// ---- CUDA IPC: how the ranks of one node see each other's block-row buffers (cudaMalloc'ed by kmg_dev_malloc) ----
int kmg_ipc_export(const void* d_ptr, uint8_t* handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    KMG_REQUIRE(d_ptr && handle64, KMG_ERR_ARG, "ipc_export: null argument");
    cudaIpcMemHandle_t h;
    KMG_CUDA_CHECK(cudaIpcGetMemHandle(&h, const_cast<void*>(d_ptr)));
    memcpy(handle64, &h, 64);
    return KMG_OK;
}

int kmg_ipc_open(const uint8_t* handle64, void** d_ptr) {
    KMG_REQUIRE(d_ptr && handle64, KMG_ERR_ARG, "ipc_open: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    KMG_CUDA_CHECK(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return KMG_OK;
}

int kmg_ipc_close(void* d_ptr) {
    if (d_ptr) KMG_CUDA_CHECK(cudaIpcCloseMemHandle(d_ptr));
    return KMG_OK;
}

int kmg_gram_i8_simt_dev(const int8_t* d_phi_rows, const int8_t* d_phi_cols, int64_t rows, int64_t cols, int64_t width,
                         int64_t ld_phi, int32_t* d_out, int64_t ldo, void* stream) {
    return kmg_gram_i8_simt_launch(d_phi_rows, d_phi_cols, ld_phi, rows, cols, width, d_out, ldo, (cudaStream_t)stream);
}

int kmg_mismatch_phi_dev(const uint32_t* d_planes, int64_t n, int L, int k, int m, int8_t* d_phi, int64_t ld_phi, void* stream) {
    return kmg_mismatch_phi_launch(d_planes, n, L, k, m, d_phi, ld_phi, (cudaStream_t)stream);
}

int kmg_phi_diag_sqrt_dev(const int8_t* d_phi, int64_t n, int64_t width, int64_t ld_phi, double* d_sd, void* stream) {
    return kmg_phi_diag_sqrt_launch(d_phi, n, width, ld_phi, d_sd, (cudaStream_t)stream);
}

static PairBlock make_block(const uint32_t* pr, const uint32_t* pc, int64_t rows, int64_t cols, int64_t r0, int64_t c0, int L,
                            void* out, int64_t ldo, int dtype, int symmetric, const double* sdr, const double* sdc) {
    PairBlock b;
    memset(&b, 0, sizeof(b));
    b.planes_rows = pr; b.planes_cols = pc; b.rows = rows; b.cols = cols; b.row_index0 = r0; b.col_index0 = c0; b.L = L;
    b.out = out; b.ldo = ldo; b.out_dtype = dtype; b.symmetric = symmetric; b.out_t = out; b.ldo_t = ldo;
    b.sd_rows = sdr; b.sd_cols = sdc;
    return b;
}

int kmg_mismatch_dev(const uint32_t* d_planes_rows, const uint32_t* d_planes_cols, int64_t rows, int64_t cols,
                     int64_t row_index0, int64_t col_index0, int L, int k, int m, void* d_out, int64_t ldo, int out_dtype,
                     int symmetric, const double* d_sd_rows, const double* d_sd_cols, void* stream) {
    KMG_REQUIRE(out_dtype == KMG_OUT_S32 || out_dtype == KMG_OUT_F64, KMG_ERR_ARG, "mismatch: bad out_dtype");
    KMG_REQUIRE(!(d_sd_rows && out_dtype != KMG_OUT_F64), KMG_ERR_ARG, "mismatch: normalisation needs the f64 output");
    KMG_REQUIRE((d_sd_rows == nullptr) == (d_sd_cols == nullptr), KMG_ERR_ARG, "mismatch: sd_rows and sd_cols go together");
    PairBlock b = make_block(d_planes_rows, d_planes_cols, rows, cols, row_index0, col_index0, L, d_out, ldo, out_dtype, symmetric,
                             d_sd_rows, d_sd_cols);
    return kmg_mismatch_launch(&b, k, m, (cudaStream_t)stream);
}

int kmg_mismatch_diag_dev(const uint32_t* d_planes, int64_t n, int L, int k, int m, double* d_sd, void* stream) {
    return kmg_mismatch_diag_launch(d_planes, n, L, k, m, d_sd, (cudaStream_t)stream);
}

int kmg_wd_dev(const uint32_t* d_planes_rows, const uint32_t* d_planes_cols, int64_t rows, int64_t cols, int64_t row_index0,
               int64_t col_index0, int L, int d, double* d_out, int64_t ldo, int symmetric, void* stream) {
    PairBlock b = make_block(d_planes_rows, d_planes_cols, rows, cols, row_index0, col_index0, L, d_out, ldo, KMG_OUT_F64, symmetric,
                             nullptr, nullptr);
    return kmg_wd_launch(&b, d, (cudaStream_t)stream);
}

int kmg_wds_dev(const uint32_t* d_planes_rows, const uint32_t* d_planes_cols, int64_t rows, int64_t cols, int64_t row_index0,
                int64_t col_index0, int L, int d, int S, double* d_out, int64_t ldo, int symmetric, void* stream) {
    PairBlock b = make_block(d_planes_rows, d_planes_cols, rows, cols, row_index0, col_index0, L, d_out, ldo, KMG_OUT_F64, symmetric,
                             nullptr, nullptr);
    return kmg_wds_launch(&b, d, S, (cudaStream_t)stream);
}

int kmg_la_dev(const uint32_t* d_planes_rows, const uint32_t* d_planes_cols, int64_t rows, int64_t cols, int64_t row_index0,
               int64_t col_index0, int L, double e, double d, double beta, int smith, double* d_out, int64_t ldo,
               int symmetric, void* stream) {
    PairBlock b = make_block(d_planes_rows, d_planes_cols, rows, cols, row_index0, col_index0, L, d_out, ldo, KMG_OUT_F64, symmetric,
                             nullptr, nullptr);
    return kmg_la_launch(&b, e, d, beta, smith, (cudaStream_t)stream);
}

int kmg_normalize_dev(double* d_K, int64_t n, int64_t ld, double* d_sd_scratch, void* stream) {
    int rc = kmg_ew_diag_sqrt(d_K, n, ld, d_sd_scratch, (cudaStream_t)stream);
    if (rc) return rc;
    return kmg_ew_normalize(d_K, n, ld, d_sd_scratch, (cudaStream_t)stream);
}

int64_t kmg_center_workspace_bytes(int64_t n) { return kmg_ew_center_workspace(n); }

int kmg_center_dev(const double* d_K, int64_t n, int64_t ld, double* d_out, int64_t ldo, void* d_workspace, void* stream) {
    return kmg_ew_center(d_K, n, ld, d_out, ldo, d_workspace, (cudaStream_t)stream);
}

int kmg_row_sums_dev(const double* d_K, int64_t rows, int64_t cols, int64_t ld, double* d_rs, void* stream) {
    return kmg_ew_row_sums(d_K, rows, cols, ld, d_rs, (cudaStream_t)stream);
}

int64_t kmg_col_sums_workspace_bytes(int64_t rows, int64_t cols) { return kmg_ew_col_sums_workspace(rows, cols); }

int kmg_col_sums_dev(const double* d_K, int64_t rows, int64_t cols, int64_t ld, double* d_cs, void* d_workspace, void* stream) {
    return kmg_ew_col_sums(d_K, rows, cols, ld, d_cs, d_workspace, (cudaStream_t)stream);
}

int kmg_center_apply_dev(const double* d_K, int64_t rows, int64_t cols, int64_t n_total, int64_t ld, const double* d_rs,
                         const double* d_cs, const double* d_g, double* d_out, int64_t ldo, void* stream) {
    return kmg_ew_center_apply(d_K, rows, cols, n_total, ld, d_rs, d_cs, d_g, d_out, ldo, (cudaStream_t)stream);
}

int kmg_gather_dev(const double* d_K, int64_t ld, const int64_t* d_idx, int64_t m, double* d_out, int64_t ldo, void* stream) {
    return kmg_ew_gather(d_K, ld, d_idx, m, d_out, ldo, (cudaStream_t)stream);
}

int kmg_combine_dev(const double* const* d_Ks, const int64_t* lds, const double* u, int p, int degree, int64_t rows, int64_t cols,
                    double* d_out, int64_t ldo, void* stream) {
    return kmg_ew_combine(d_Ks, lds, u, p, degree, rows, cols, d_out, ldo, (cudaStream_t)stream);
}

int kmg_matvec_dev(const double* d_K, int64_t rows, int64_t cols, int64_t ld, const double* d_v, double* d_out, void* stream) {
    return kmg_ew_row_wsums(d_K, rows, cols, ld, d_v, d_out, (cudaStream_t)stream);
}

int64_t kmg_spd_solve_workspace_bytes(int64_t n) { return kmg_solve_workspace(n); }

int kmg_spd_solve_dev(const double* d_K, int64_t n, int64_t ld, const double* d_s, double c, const double* d_b, double* d_x, void* d_work,
                      void* stream) {
    int rc = kmg_rt_require_device();
    if (rc) return rc;
    int* flag = nullptr;
    if ((rc = kmg_spd_solve_launch(d_K, n, ld, d_s, c, d_b, d_x, d_work, &flag, (cudaStream_t)stream))) return rc;
    int h = 0;
    KMG_CUDA_CHECK(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    KMG_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    KMG_REQUIRE(h == 0, KMG_ERR_ARG, "spd_solve: the matrix S K S + c I is not positive definite");
    return KMG_OK;
}

int kmg_spd_solve_host(const double* K, int64_t n, int64_t ldk, const int64_t* idx, int64_t nfit, const double* s, double c,
                       const double* b, double* x) {
    int rc = kmg_rt_require_device();
    if (rc) return rc;
    KMG_REQUIRE(K && b && x && n >= 0 && ldk >= n && (idx != nullptr || nfit == n) && nfit >= 0, KMG_ERR_ARG, "spd_solve: bad arguments");
    if (nfit == 0) return KMG_OK;
    for (int64_t t = 0; idx && t < nfit; ++t) KMG_REQUIRE(idx[t] >= 0 && idx[t] < n, KMG_ERR_ARG, "spd_solve: index out of range");
    cudaStream_t st;
    if ((rc = kmg_rt_get_streams(&st, nullptr))) return rc;
    // K_fit = K[idx][:, idx] (KRR.py:30, KLR.py:67) gathered on the host: nfit^2 doubles cross PCIe, not n^2
    std::vector<double> sub((size_t)nfit * nfit);
    for (int64_t a = 0; a < nfit; ++a) {
        const double* row = K + (idx ? idx[a] : a) * ldk;
        for (int64_t q = 0; q < nfit; ++q) sub[a * nfit + q] = row[idx ? idx[q] : q];
    }
    DevBuf dK, ds, db, dx, work;
    if ((rc = dK.alloc((size_t)nfit * nfit * 8))) return rc;
    if ((rc = db.alloc((size_t)nfit * 8))) return rc;
    if ((rc = dx.alloc((size_t)nfit * 8))) return rc;
    if ((rc = work.alloc((size_t)kmg_solve_workspace(nfit)))) return rc;
    KMG_CUDA_CHECK(cudaMemcpyAsync(dK.p, sub.data(), sub.size() * 8, cudaMemcpyHostToDevice, st));
    KMG_CUDA_CHECK(cudaMemcpyAsync(db.p, b, (size_t)nfit * 8, cudaMemcpyHostToDevice, st));
    if (s) {
        if ((rc = ds.alloc((size_t)nfit * 8))) return rc;
        KMG_CUDA_CHECK(cudaMemcpyAsync(ds.p, s, (size_t)nfit * 8, cudaMemcpyHostToDevice, st));
    }
    rc = kmg_spd_solve_dev(dK.as<double>(), nfit, nfit, s ? ds.as<double>() : nullptr, c, db.as<double>(), dx.as<double>(), work.p, st);
    if (rc) { cudaStreamSynchronize(st); return rc; }
    KMG_CUDA_CHECK(cudaMemcpyAsync(x, dx.p, (size_t)nfit * 8, cudaMemcpyDeviceToHost, st));
    KMG_CUDA_CHECK(cudaStreamSynchronize(st));
    return KMG_OK;
}

int kmg_weighted_dot_dev(const double* d_A, int64_t lda, const double* d_B, int64_t ldb, const double* d_w, int64_t n,
                         double* d_partial, double* d_result, void* stream) {
    return kmg_ew_weighted_dot(d_A, lda, d_B, ldb, d_w, n, d_partial, d_result, (cudaStream_t)stream);
}

}  // extern "C"

