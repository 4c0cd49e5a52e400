// fused.cu -- ALIGNF / NLCK on top of Gram construction WITHOUT the Grams ever visiting the host: the steps the reference
// applies to finished Gram matrices (ALIGNF.py:28-58,91-94; NLCKernels.py:33,36,43-48,52,97-99) run in the epilogues of
// the kernels that produce the entries (epi_ops.cuh) or as passes over device-resident fit sub-blocks.
//
//   kmg_alignf_fused_host   sequences + method list + fit indices + labels -> a (p), M (p x p).  The fit sub-block
//                           K[idx][:, idx] of every method is built directly from the gathered sequences (a Gram entry
//                           depends on its two sequences only, so it IS the sub-block of the full Gram, bit for bit);
//                           the producing kernel's epilogue emits the row sums r = K 1 (the centring statistics) and
//                           K y~; a_m = y~' K_m y~ (= <H K_m H, y y'>_F); M_ij = <H K_i H, H K_j H>_F from
//                           the device-resident sub-blocks, centred on the fly from r.  PCIe: n*L bytes up, (p + p^2)
//                           doubles down -- against p n^2 doubles down and p nfit^2 up for the array-based path.
//   kmg_combine_fused_host  sequences + method list + weights -> Km = (sum_m u_m K_m)**degree [normalised]: one Gram
//                           launch per method accumulates u_m K_m into the one output buffer in its epilogue (product and
//                           sum rounded separately, in method order: bit-identical to numpy's sum over the stacked
//                           kernels); the last launch applies the power and the final normalize_K.  Only Km crosses PCIe.
//   kmg_build_grams_dev     the (optionally normalised, optionally sub-sampled) Grams of a method list as device-resident
//                           matrices for the iterating consumers (NLCK's 50 K-lines and gradients, kmg/resident.py).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../include/kmg.h"
#include "api_internal.h"
#include "elementwise.h"
#include "epi_ops.cuh"
#include "gram_i8.h"
#include "host_link.h"
#include "kmg_common.cuh"
#include "pair_kernels.h"
#include "runtime.h"
#include "seq_kernels.h"

namespace {

__global__ void gather_planes_kernel(const uint32_t* __restrict__ planes, const int64_t* __restrict__ idx, int64_t m,
                                     uint32_t* __restrict__ out) {
    const int64_t t = (blockIdx.x * 256ll + threadIdx.x) >> 3;
    const int w = threadIdx.x & 7;
    if (t < m) out[t * KMG_SEQ_WORDS + w] = planes[idx[t] * KMG_SEQ_WORDS + w];
}

int check_methods(const kmg_method_t* methods, int p, int L) {
    KMG_REQUIRE(methods != nullptr && p >= 1 && p <= KMG_MAX_COMBINE, KMG_ERR_ARG, "fused: between 1 and %d methods", KMG_MAX_COMBINE);
    for (int i = 0; i < p; ++i) {
        const kmg_method_t& m = methods[i];
        switch (m.kind) {
            case KMG_KIND_SP: KMG_REQUIRE(m.k >= 1 && m.k <= KMG_MAX_DENSE_K && m.k <= L, KMG_ERR_UNSUPPORTED, "fused: spectrum k must be 1..%d", KMG_MAX_DENSE_K); break;
            case KMG_KIND_MM: KMG_REQUIRE(m.k >= 1 && m.k <= L && m.m >= 0 && m.m <= KMG_MM_MAX_M, KMG_ERR_UNSUPPORTED, "fused: mismatch (k, m) out of range"); break;
            case KMG_KIND_WD: KMG_REQUIRE(m.d >= 1 && m.d <= 127, KMG_ERR_ARG, "fused: weighted degree d out of range"); break;
            case KMG_KIND_WDS: KMG_REQUIRE(m.d >= 1 && m.d <= 127 && m.S >= 0 && m.S <= 7, KMG_ERR_ARG, "fused: WDS (d, S) out of range"); break;
            case KMG_KIND_LA: KMG_REQUIRE(m.beta > 0.0, KMG_ERR_ARG, "fused: local alignment needs beta > 0"); break;
            default: KMG_REQUIRE(false, KMG_ERR_ARG, "fused: unknown method kind %d", m.kind);
        }
    }
    return KMG_OK;
}

// Everything one method needs to produce blocks of its Gram over a fixed set of sequences.
struct MethodCtx {
    kmg_method_t m;
    int L;
    const uint32_t* planes = nullptr;  // the sequences of the Gram (n)
    int64_t n = 0;
    DevBuf phi, sd;                    // dense feature map (SP, dense MM), sqrt(diag) when the kernel is cosine-normalised
    const double* sdp = nullptr;       // null: stored unnormalised
    bool dense = false;
    bool producer_normalises = true;   // false: normalisation needs a stored copy (WDS, LA under normalize_inputs)
};

// sd of ALL sequences of `full_planes` decides the reference's normalize_K early-out (K[0,0] == 1 leaves the kernel as it is,
// kernels.py:404-405); the context itself may cover a gathered subset (idx != null).
int prepare(MethodCtx* c, const kmg_method_t& m, int L, const uint32_t* planes_sel, int64_t nsel, const uint32_t* planes_full, int64_t nfull,
            const int64_t* d_idx, bool normalize_inputs, cudaStream_t s) {
    c->m = m; c->L = L; c->planes = planes_sel; c->n = nsel;
    int rc;
    const bool mm = m.kind == KMG_KIND_MM;
    const bool want_norm = mm || normalize_inputs;  // get_mismatch_K always normalises (kernels.py:216)
    c->dense = m.kind == KMG_KIND_SP || (mm && m.k <= KMG_MAX_DENSE_K && m.m <= 3 && L - m.k + 1 <= 127);
    if (c->dense) {
        const int ks[1] = {m.k};
        const int64_t width = mm ? ((1ll << (2 * m.k)) + 127) / 128 * 128 : kmg_spectrum_padded_width(ks, 1, L);
        if ((rc = c->phi.alloc((size_t)nsel * width))) return rc;
        rc = mm ? kmg_mismatch_phi_launch(planes_sel, nsel, L, m.k, m.m, c->phi.as<int8_t>(), width, s)
                : kmg_spectrum_phi_launch(planes_sel, nsel, L, ks, 1, c->phi.as<int8_t>(), width, s);
        if (rc) return rc;
    }
    if (!want_norm) return KMG_OK;
    if (m.kind == KMG_KIND_WDS || m.kind == KMG_KIND_LA) { c->producer_normalises = false; return KMG_OK; }
    // sqrt(diag) of the selected sequences + the value for sequence 0 of the FULL set (the early-out test)
    if ((rc = c->sd.alloc((size_t)std::max<int64_t>(nsel, 1) * 8))) return rc;
    double sd0 = 0.0;
    if (m.kind == KMG_KIND_WD) {
        const double diag = (double)(L - 1) + (double)(1 - m.d) / 3.0;  // kernels.py:96
        std::vector<double> h((size_t)nsel, sqrt(diag));
        KMG_CUDA_CHECK(cudaMemcpyAsync(c->sd.p, h.data(), (size_t)nsel * 8, cudaMemcpyHostToDevice, s));
        KMG_CUDA_CHECK(cudaStreamSynchronize(s));
        sd0 = sqrt(diag);
    } else {
        DevBuf first;  // sqrt(diag) of sequence 0 of the full set
        if ((rc = first.alloc(8))) return rc;
        if (c->dense) {
            const int ks[1] = {m.k};
            const int64_t width = mm ? ((1ll << (2 * m.k)) + 127) / 128 * 128 : kmg_spectrum_padded_width(ks, 1, L);
            if ((rc = kmg_phi_diag_sqrt_launch(c->phi.as<int8_t>(), nsel, width, width, c->sd.as<double>(), s))) return rc;
            if (d_idx == nullptr) {
                KMG_CUDA_CHECK(cudaMemcpyAsync(&sd0, c->sd.p, 8, cudaMemcpyDeviceToHost, s));
            } else {
                DevBuf phi0;
                if ((rc = phi0.alloc((size_t)width))) return rc;
                rc = mm ? kmg_mismatch_phi_launch(planes_full, 1, L, m.k, m.m, phi0.as<int8_t>(), width, s)
                        : kmg_spectrum_phi_launch(planes_full, 1, L, ks, 1, phi0.as<int8_t>(), width, s);
                if (rc) return rc;
                if ((rc = kmg_phi_diag_sqrt_launch(phi0.as<int8_t>(), 1, width, width, first.as<double>(), s))) return rc;
                KMG_CUDA_CHECK(cudaMemcpyAsync(&sd0, first.p, 8, cudaMemcpyDeviceToHost, s));
                KMG_CUDA_CHECK(cudaStreamSynchronize(s));
            }
        } else {
            if ((rc = kmg_mismatch_diag_launch(planes_sel, nsel, L, m.k, m.m, c->sd.as<double>(), s))) return rc;
            if ((rc = kmg_mismatch_diag_launch(planes_full, 1, L, m.k, m.m, first.as<double>(), s))) return rc;
            KMG_CUDA_CHECK(cudaMemcpyAsync(&sd0, first.p, 8, cudaMemcpyDeviceToHost, s));
        }
        KMG_CUDA_CHECK(cudaStreamSynchronize(s));
    }
    (void)nfull;
    // normalize_K early-out: raw K[0,0] == 1 <=> sd[0] == 1.  For MM under normalize_inputs the kernel is normalised once by
    // get_mismatch_K; the second normalize_K then meets K[0,0] == 1 and returns it untouched -- the same single normalisation.
    c->sdp = (sd0 == 1.0) ? nullptr : c->sd.as<double>();
    return KMG_OK;
}

// the n x n Gram of the context's method into `out` (row stride ld), plain or symmetric, with optional fused steps
int build(MethodCtx* c, double* out, int64_t ld, bool symmetric, const EpiOps* epi, cudaStream_t s) {
    const int64_t n = c->n;
    if (n == 0) return KMG_OK;
    if (c->dense) {
        const int ks[1] = {c->m.k};
        const int64_t width = c->m.kind == KMG_KIND_MM ? ((1ll << (2 * c->m.k)) + 127) / 128 * 128 : kmg_spectrum_padded_width(ks, 1, c->L);
        GramI8Args a;
        memset(&a, 0, sizeof(a));
        a.phi_rows = c->phi.as<int8_t>(); a.phi_cols = a.phi_rows; a.rows = n; a.cols = n; a.Dpad = width; a.ld_phi = width;
        a.out = out; a.ldo = ld; a.out_dtype = KMG_OUT_F64; a.sd_rows = c->sdp; a.sd_cols = c->sdp;
        const bool fused = epi != nullptr && epi_active(*epi);
        a.symmetric = (symmetric && !fused) ? 1 : 0;  // the fused GEMM variant takes plain blocks (the work is negligible at these sizes)
        a.out_t = out; a.ldo_t = ld; a.epi = fused ? epi : nullptr;
        return kmg_gram_i8_launch(&a, s);
    }
    PairBlock b;
    memset(&b, 0, sizeof(b));
    b.planes_rows = c->planes; b.planes_cols = c->planes; b.rows = n; b.cols = n; b.L = c->L;
    b.out = out; b.ldo = ld; b.out_dtype = KMG_OUT_F64; b.symmetric = symmetric ? 1 : 0; b.out_t = out; b.ldo_t = ld;
    b.sd_rows = c->sdp; b.sd_cols = c->sdp; b.epi = epi;
    switch (c->m.kind) {
        case KMG_KIND_MM: return kmg_mismatch_launch(&b, c->m.k, c->m.m, s);
        case KMG_KIND_WD: return kmg_wd_launch(&b, c->m.d, s);
        case KMG_KIND_WDS: b.sd_rows = b.sd_cols = nullptr; return kmg_wds_launch(&b, c->m.d, c->m.S, s);
        default: b.sd_rows = b.sd_cols = nullptr; b.epi = nullptr; return kmg_la_launch(&b, c->m.e, c->m.dd, c->m.beta, c->m.smith, s);
    }
}

struct Seqs {
    DevBuf planes, sel, didx;
    const uint32_t* full = nullptr;
    const uint32_t* use = nullptr;
    int64_t n = 0, nsel = 0;
};

int upload_and_select(const uint8_t* seqs, int64_t n, int L, int fmt, const int64_t* idx, int64_t nsel, Seqs* q, cudaStream_t s) {
    int rc = kmg_api_upload_planes(seqs, n, L, fmt, &q->planes, s);
    if (rc) return rc;
    q->full = q->planes.as<uint32_t>(); q->n = n;
    if (idx == nullptr) { q->use = q->full; q->nsel = n; return KMG_OK; }
    for (int64_t t = 0; t < nsel; ++t) KMG_REQUIRE(idx[t] >= 0 && idx[t] < n, KMG_ERR_ARG, "fused: index out of range");
    if ((rc = q->didx.alloc((size_t)std::max<int64_t>(nsel, 1) * 8))) return rc;
    if ((rc = q->sel.alloc((size_t)std::max<int64_t>(nsel, 1) * KMG_SEQ_WORDS * 4))) return rc;
    if (nsel > 0) {
        KMG_CUDA_CHECK(cudaMemcpyAsync(q->didx.p, idx, (size_t)nsel * 8, cudaMemcpyHostToDevice, s));
        gather_planes_kernel<<<(unsigned)((nsel * 8 + 255) / 256), 256, 0, s>>>(q->full, q->didx.as<int64_t>(), nsel, q->sel.as<uint32_t>());
        KMG_CUDA_CHECK(cudaGetLastError());
        KMG_CUDA_CHECK(cudaStreamSynchronize(s));  // idx is the caller's
    }
    q->use = q->sel.as<uint32_t>(); q->nsel = nsel;
    return KMG_OK;
}

}  // namespace

extern "C" {

int kmg_alignf_fused_host(const uint8_t* seqs, int64_t n, int L, int seq_format, const kmg_method_t* methods, int p,
                          const int64_t* idx, int64_t nfit, const double* y, double* a, double* M, int64_t* pcie_bytes) {
    int rc = kmg_rt_require_device();
    if (rc) return rc;
    KMG_REQUIRE(n >= 0 && nfit >= 0 && idx && y && a && M && (seqs || n == 0), KMG_ERR_ARG, "alignf_fused: bad arguments");
    if ((rc = check_methods(methods, p, L))) return rc;
    if (pcie_bytes) { pcie_bytes[0] = n * L + nfit * 16; pcie_bytes[1] = (int64_t)(p + p * p) * 8; }
    if (nfit == 0) { for (int i = 0; i < p; ++i) { a[i] = 0; for (int j = 0; j < p; ++j) M[i * p + j] = 0; } return KMG_OK; }
    cudaStream_t s;
    if ((rc = kmg_rt_get_streams(&s, nullptr))) return rc;
    Seqs q;
    if ((rc = upload_and_select(seqs, n, L, seq_format, idx, nfit, &q, s))) return rc;
    // y~ = y - mean(y): <H K H, y y'>_F = (H y)' K (H y)
    std::vector<double> yt((size_t)nfit);
    double mean = 0.0;
    for (int64_t t = 0; t < nfit; ++t) mean += y[t];
    mean /= (double)nfit;
    for (int64_t t = 0; t < nfit; ++t) yt[t] = y[t] - mean;
    const int64_t n_chunks = (nfit + 31) / 32;
    DevBuf dyt, psum, pwsum, part, res;
    std::vector<DevBuf> K(p), r(p), kw(p);
    if ((rc = dyt.alloc((size_t)nfit * 8))) return rc;
    if ((rc = psum.alloc((size_t)nfit * n_chunks * 8))) return rc;
    if ((rc = pwsum.alloc((size_t)nfit * n_chunks * 8))) return rc;
    if ((rc = part.alloc((size_t)nfit * 8))) return rc;
    if ((rc = res.alloc((size_t)(2 * p + p * p) * 8))) return rc;  // [0,p): a   [p,2p): grand sums   [2p, ...): M
    KMG_CUDA_CHECK(cudaMemcpyAsync(dyt.p, yt.data(), (size_t)nfit * 8, cudaMemcpyHostToDevice, s));
    double* d_a = res.as<double>();
    double* d_g = d_a + p;
    double* d_M = d_g + p;
    for (int i = 0; i < p; ++i) {
        MethodCtx c;
        if ((rc = prepare(&c, methods[i], L, q.use, nfit, q.full, n, q.didx.as<int64_t>(), false, s))) return rc;
        if ((rc = K[i].alloc((size_t)nfit * nfit * 8))) return rc;
        if ((rc = r[i].alloc((size_t)nfit * 8))) return rc;
        if ((rc = kw[i].alloc((size_t)nfit * 8))) return rc;
        if (methods[i].kind == KMG_KIND_LA) {
            // FP64-bound producer without a row-statistics epilogue (1e5 flops per entry): statistics from the stored block
            if ((rc = build(&c, K[i].as<double>(), nfit, true, nullptr, s))) return rc;
            if ((rc = kmg_ew_row_sums(K[i].as<double>(), nfit, nfit, nfit, r[i].as<double>(), s))) return rc;
            if ((rc = kmg_ew_row_wsums(K[i].as<double>(), nfit, nfit, nfit, dyt.as<double>(), kw[i].as<double>(), s))) return rc;
        } else {
            EpiOps e;
            memset(&e, 0, sizeof(e));
            e.row_sum_partial = psum.as<double>(); e.row_wsum_partial = pwsum.as<double>(); e.w_cols = dyt.as<double>(); e.n_chunks = n_chunks;
            if ((rc = build(&c, K[i].as<double>(), nfit, false, &e, s))) return rc;
            if ((rc = kmg_ew_partial_rows_reduce(psum.as<double>(), nfit, n_chunks, r[i].as<double>(), s))) return rc;
            if ((rc = kmg_ew_partial_rows_reduce(pwsum.as<double>(), nfit, n_chunks, kw[i].as<double>(), s))) return rc;
        }
        if ((rc = kmg_ew_vec_sum(r[i].as<double>(), nfit, d_g + i, s))) return rc;                           // 1' K 1
        if ((rc = kmg_ew_vec_dot(dyt.as<double>(), kw[i].as<double>(), nfit, d_a + i, s))) return rc;       // a_i = y~' K y~ (ALIGNF.py:43-48)
        KMG_CUDA_CHECK(cudaStreamSynchronize(s));  // the context's feature map and the partial buffers are reused
    }
    for (int i = 0; i < p; ++i)
        for (int j = i; j < p; ++j)  // M_ij = <Kc_i, Kc_j>_F (ALIGNF.py:50-58), both factors centred on the fly
            if ((rc = kmg_ew_centered_dot(K[i].as<double>(), nfit, r[i].as<double>(), d_g + i, K[j].as<double>(), nfit, r[j].as<double>(), d_g + j,
                                          nfit, part.as<double>(), d_M + i * p + j, s))) return rc;
    std::vector<double> h((size_t)(2 * p + p * p), 0.0);
    KMG_CUDA_CHECK(cudaMemcpyAsync(h.data(), res.p, h.size() * 8, cudaMemcpyDeviceToHost, s));
    KMG_CUDA_CHECK(cudaStreamSynchronize(s));
    for (int i = 0; i < p; ++i) a[i] = h[i];
    for (int i = 0; i < p; ++i)
        for (int j = i; j < p; ++j) M[i * p + j] = M[j * p + i] = h[2 * p + i * p + j];
    return KMG_OK;
}

int kmg_combine_fused_host(const uint8_t* seqs, int64_t n, int L, int seq_format, const kmg_method_t* methods, int p,
                           const double* u, int degree, int normalize_inputs, int normalize, double* Km, int64_t ldk, int64_t* pcie_bytes) {
    int rc = kmg_rt_require_device();
    if (rc) return rc;
    KMG_REQUIRE(n >= 0 && u && Km && ldk >= n && degree >= 1 && degree <= 64 && (seqs || n == 0), KMG_ERR_ARG, "combine_fused: bad arguments");
    if ((rc = check_methods(methods, p, L))) return rc;
    if (pcie_bytes) { pcie_bytes[0] = n * L; pcie_bytes[1] = n * n * 8; }
    if (n == 0) return KMG_OK;
    cudaStream_t s;
    if ((rc = kmg_rt_get_streams(&s, nullptr))) return rc;
    Seqs q;
    if ((rc = upload_and_select(seqs, n, L, seq_format, nullptr, 0, &q, s))) return rc;
    DevBuf out, scratch, sds, post;
    if ((rc = out.alloc((size_t)n * n * 8))) return rc;
    // final normalize_K of the combination (NLCKernels.py:99): only meaningful when every input has unit diagonal, where the
    // diagonal of the combination is the constant (sum_m u_m)**degree, summed and raised the way numpy does; its early-out
    // (K[0,0] == 1) leaves the combination as it is
    const double* post_sd = nullptr;
    if (normalize) {
        KMG_REQUIRE(normalize_inputs, KMG_ERR_UNSUPPORTED, "combine_fused: the fused final normalisation needs normalised inputs (unit diagonals)");
        double d = u[0] * 1.0;
        for (int m = 1; m < p; ++m) d = d + u[m] * 1.0;
        d = degree == 1 ? d : (degree == 2 ? d * d : pow(d, (double)degree));
        if (d != 1.0) {
            std::vector<double> h((size_t)n, sqrt(d));
            if ((rc = post.alloc((size_t)n * 8))) return rc;
            KMG_CUDA_CHECK(cudaMemcpyAsync(post.p, h.data(), (size_t)n * 8, cudaMemcpyHostToDevice, s));
            KMG_CUDA_CHECK(cudaStreamSynchronize(s));
            post_sd = post.as<double>();
        }
    }
    for (int m = 0; m < p; ++m) {
        MethodCtx c;
        if ((rc = prepare(&c, methods[m], L, q.use, n, q.full, n, nullptr, normalize_inputs != 0, s))) return rc;
        EpiOps e;
        memset(&e, 0, sizeof(e));
        e.accumulate = m == 0 ? 1 : 2; e.u = u[m];
        if (m == p - 1) { e.post_degree = degree; e.post_sd_rows = post_sd; e.post_sd_cols = post_sd; }
        if (methods[m].kind == KMG_KIND_LA || !c.producer_normalises) {
            // producers without the fused steps (LA) or without a cheap diagonal (WDS / LA under normalize_inputs): build the
            // kernel once into scratch and accumulate it with the same arithmetic in a stored-Gram pass
            if (!scratch.p && (rc = scratch.alloc((size_t)n * n * 8))) return rc;
            if ((rc = build(&c, scratch.as<double>(), n, true, nullptr, s))) return rc;
            const double* sd = nullptr;
            if (!c.producer_normalises) {
                double k00 = 0.0;
                KMG_CUDA_CHECK(cudaMemcpyAsync(&k00, scratch.p, 8, cudaMemcpyDeviceToHost, s));
                KMG_CUDA_CHECK(cudaStreamSynchronize(s));
                if (k00 != 1.0) {
                    if (!sds.p && (rc = sds.alloc((size_t)n * 8))) return rc;
                    if ((rc = kmg_ew_diag_sqrt(scratch.as<double>(), n, n, sds.as<double>(), s))) return rc;
                    sd = sds.as<double>();
                }
            }
            if ((rc = kmg_ew_accumulate(scratch.as<double>(), n, sd, &e, n, out.as<double>(), n, s))) return rc;
        } else {
            if ((rc = build(&c, out.as<double>(), n, true, &e, s))) return rc;
        }
        KMG_CUDA_CHECK(cudaStreamSynchronize(s));
    }
    KMG_CUDA_CHECK(cudaMemcpy2DAsync(Km, (size_t)ldk * 8, out.p, (size_t)n * 8, (size_t)n * 8, (size_t)n, cudaMemcpyDeviceToHost, s));
    KMG_CUDA_CHECK(cudaStreamSynchronize(s));
    return KMG_OK;
}

int kmg_build_grams_dev(const uint8_t* seqs, int64_t n, int L, int seq_format, const kmg_method_t* methods, int p,
                        const int64_t* idx, int64_t nsel, int normalize_inputs, void* const* d_out) {
    int rc = kmg_rt_require_device();
    if (rc) return rc;
    KMG_REQUIRE(n >= 0 && d_out && (seqs || n == 0) && (idx != nullptr || nsel == n), KMG_ERR_ARG, "build_grams: bad arguments");
    if ((rc = check_methods(methods, p, L))) return rc;
    cudaStream_t s;
    if ((rc = kmg_rt_get_streams(&s, nullptr))) return rc;
    Seqs q;
    if ((rc = upload_and_select(seqs, n, L, seq_format, idx, nsel, &q, s))) return rc;
    const int64_t ns = q.nsel;
    if (ns == 0) return KMG_OK;
    DevBuf sds;
    for (int m = 0; m < p; ++m) {
        KMG_REQUIRE(d_out[m] != nullptr, KMG_ERR_ARG, "build_grams: null output buffer");
        MethodCtx c;
        if ((rc = prepare(&c, methods[m], L, q.use, ns, q.full, n, idx ? q.didx.as<int64_t>() : nullptr, normalize_inputs != 0, s))) return rc;
        double* out = static_cast<double*>(d_out[m]);
        if ((rc = build(&c, out, ns, true, nullptr, s))) return rc;
        if (!c.producer_normalises) {
            // WDS / LA under normalize_inputs: normalize_K on the stored block.  The early-out looks at K[0,0] of the FULL
            // kernel; for these kernels the diagonal entry of sequence 0 is only available when sequence 0 is selected.
            KMG_REQUIRE(idx == nullptr, KMG_ERR_UNSUPPORTED, "build_grams: normalised WDS / LA sub-blocks are not supported (build the full kernel)");
            double k00 = 0.0;
            KMG_CUDA_CHECK(cudaMemcpyAsync(&k00, out, 8, cudaMemcpyDeviceToHost, s));
            KMG_CUDA_CHECK(cudaStreamSynchronize(s));
            if (k00 != 1.0) {
                if (!sds.p && (rc = sds.alloc((size_t)ns * 8))) return rc;
                if ((rc = kmg_ew_diag_sqrt(out, ns, ns, sds.as<double>(), s))) return rc;
                if ((rc = kmg_ew_normalize(out, ns, ns, sds.as<double>(), s))) return rc;
            }
        }
        KMG_CUDA_CHECK(cudaStreamSynchronize(s));
    }
    return KMG_OK;
}

}  // extern "C"
