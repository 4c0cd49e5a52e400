/*
 * kmg.h -- C-ABI of libkmg.so: B200-native (sm_100a) Gram / cross-Gram construction for DNA string
 * kernels, the drop-in boundary for the hot path of afiliot/Kernel-Methods-For-Genomics.
 *
 * The reference has no FFI layer: its operator API for this path is the Python module `kernels`
 * (SURVEY.md section 8b).  Each entry point below names the reference function it replaces
 * (file:line in /root/reference); `kernel-methods-for-genomics_b200/kernels.py` is the ctypes
 * binding a maintainer would drop in place of the reference's kernels.py.
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns 0 (KMG_OK) or a negative error code and
 *     leaves a message retrievable with kmg_last_error() (thread local).
 *   - sequences: n x L bytes, row major, no terminators; either ASCII 'A','C','G','T' or integer
 *     codes 0..3 (A<C<G<T as in kernels.py:37,184).  All sequences of one call have the same length
 *     L <= 128 (the challenge data is 101 bp).  Any other byte -> KMG_ERR_ALPHABET.
 *   - Gram matrices: row-major double, leading dimension given in elements.
 *   - `*_host` functions take HOST pointers and do the host<->device copies themselves (this is what
 *     the Python shim calls, one call per Gram); `*_dev` functions take DEVICE pointers plus a
 *     cudaStream_t passed as void* and only enqueue work (block-row construction for the multi-GPU
 *     layer and the benchmark).
 *   - there is no CPU fallback anywhere behind this interface: without a CUDA device every compute
 *     entry point fails with KMG_ERR_CUDA.
 */
#ifndef KMG_H
#define KMG_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KMG_OK 0
#define KMG_ERR_CUDA (-1)
#define KMG_ERR_ARG (-2)
#define KMG_ERR_ALPHABET (-3)
#define KMG_ERR_UNSUPPORTED (-4)
#define KMG_ERR_NOMEM (-5)

#define KMG_OUT_S32 0
#define KMG_OUT_F64 1

#define KMG_SEQ_ASCII 1
#define KMG_SEQ_CODES 0

/* how the mirrored blocks of a sharded symmetric Gram reach their owners (kmg_gram_i8_sharded_dev) */
#define KMG_EXCH_SINGLE 0  /* one launch, thread-issued stores into the owners' buffers (buffers on one device / tests) */
#define KMG_EXCH_STAGED 1  /* per peer block: transposed block into local staging + one pitched peer copy */
#define KMG_EXCH_DIRECT 2  /* per peer block: the epilogue's TMA stores write the owner's buffer over NVLink */
#define KMG_EXCH_DEFER_JOIN 0x100 /* OR-ed to KMG_EXCH_STAGED: do not make the stream wait for the peer copies; the caller enqueues
                                     more work (the plain remainder of a wider block-row) and then calls kmg_gram_sharded_join */

#define KMG_MM_AUTO 0      /* dense feature map + tensor-core GEMM for k <= 8, pairwise bit-vector kernel above */
#define KMG_MM_PAIRWISE 1
#define KMG_MM_DENSE 2

/* ---- library ------------------------------------------------------------------------------- */
int kmg_version(void);
const char* kmg_last_error(void);
int kmg_device_count(void);          /* 0 when no usable CUDA device */
int kmg_set_device(int device);
int kmg_release(void);               /* frees cached device buffers of the current device and cached host blocks */
/* plain device buffers for callers that keep Grams resident between calls (kmg/resident.py: NLCK's 50 iterations
 * reuse the fit sub-blocks instead of re-uploading them); synchronous copies on the default stream */
int kmg_dev_malloc(int64_t bytes, void** ptr);
int kmg_dev_free(void* ptr);
int kmg_dev_upload(void* d_dst, const void* h_src, int64_t bytes);
int kmg_dev_download(void* h_dst, const void* d_src, int64_t bytes);
/* recycled host memory for result arrays (every reference builder allocates its result, e.g. kernels.py:37 np.zeros):
 * 2 MB aligned, huge-page advised, pageable; kmg_host_free returns the block to a bounded cache
 * (KMG_HOST_POOL_BYTES, default 8 GiB; kmg_release empties it) so later results land in memory that is already mapped */
int kmg_host_alloc(int64_t bytes, void** ptr);
int kmg_host_free(void* ptr);
/* How fp64 results reach host arrays that live in kmg_host_alloc blocks (others always take KMG_D2H_WIDEN):
 *   KMG_D2H_WIDEN  narrow integer transport over PCIe (2-4 B/entry) + copy threads that widen to fp64: fastest for one
 *                  process with many cores; every delivered byte is a CPU store.
 *   KMG_D2H_DMA    the kernel writes fp64 on the device, the copy engine writes it straight into the (pinned) array:
 *                  8 B/entry on the link, no CPU in the data path -- scales with the GPUs of a host.
 *   KMG_D2H_MAPPED the kernel stores fp64 through the mapped address of the pinned array (zero copy).
 * Default: env KMG_D2H_MODE (widen | dma | mapped), else widen; kmg/host.py picks per process by timing (auto). */
#define KMG_D2H_WIDEN 0
#define KMG_D2H_DMA 1
#define KMG_D2H_MAPPED 2
int kmg_set_d2h_mode(int mode);
int kmg_get_d2h_mode(void);

/* ---- host-buffer entry points (the reference-facing boundary) ------------------------------ */
/* cols == NULL: symmetric Gram of `rows` (n x n, upper triangle computed and mirrored, as the
 * reference does); otherwise the nr x nc cross-Gram K[i][j] = k(rows[i], cols[j]). */

/* get_spectrum_K (kernels.py:28-47), summed over ks[0..nk) (one k: the reference call). Unnormalised. */
int kmg_spectrum_host(const uint8_t* rows, int64_t nr, const uint8_t* cols, int64_t nc, int L, int seq_format,
                      const int* ks, int nk, double* K, int64_t ldk);
/* get_mismatch_K (kernels.py:196-217).  normalize=1 reproduces the reference (normalize_K applied,
 * symmetric Gram only); normalize=0 returns the raw integer Gram. */
int kmg_mismatch_host(const uint8_t* rows, int64_t nr, const uint8_t* cols, int64_t nc, int L, int seq_format,
                      int k, int m, int normalize, int algo, double* K, int64_t ldk);
/* get_phi_u (kernels.py:12-25) / get_phi_km (kernels.py:161-175) for n sequences: int8 feature rows,
 * columns in product('ACGT', repeat=k) order (segments concatenated over ks), row stride ld >= padded width. */
int kmg_spectrum_phi_host(const uint8_t* seqs, int64_t n, int L, int seq_format, const int* ks, int nk, int8_t* phi, int64_t ld);
int kmg_mismatch_phi_host(const uint8_t* seqs, int64_t n, int L, int seq_format, int k, int m, int8_t* phi, int64_t ld);
/* get_WD_K (kernels.py:84-101).  Symmetric: diagonal is the closed form of kernels.py:96. */
int kmg_wd_host(const uint8_t* rows, int64_t nr, const uint8_t* cols, int64_t nc, int L, int seq_format,
                int d, double* K, int64_t ldk);
/* get_WDShifts_K (kernels.py:138-155): weighted degree with shifts 0..S, delta_s = 1/2/(s+1); fp64 accumulation in
 * the reference's (k, i, s) order (bit-exact).  0 <= S <= 7. */
int kmg_wds_host(const uint8_t* rows, int64_t nr, const uint8_t* cols, int64_t nc, int L, int seq_format,
                 int d, int S, double* K, int64_t ldk);
/* get_LA_K pair loop (kernels.py:287-291) with the INTENDED affine_align / Smith_Waterman
 * (kernels.py:226-270; the reference as written returns 0.0 for every pair -- see DESIGN.md). */
int kmg_la_host(const uint8_t* rows, int64_t nr, const uint8_t* cols, int64_t nc, int L, int seq_format,
                double e, double d, double beta, int smith, double* K, int64_t ldk);

/* normalize_K (kernels.py:398-415): in place.  Returns 1 if K[0][0]==1 (reference early-out, K
 * untouched), 0 after normalising, <0 on error. */
int kmg_normalize_host(double* K, int64_t n, int64_t ldk);
/* center_K (kernels.py:387-395): out = (I-11'/n) K (I-11'/n), closed form. */
int kmg_center_host(const double* K, int64_t n, int64_t ldk, double* out, int64_t ldo);
/* ALIGNF.get_K (ALIGNF.py:93) with degree=1; NLCK (NLCKernels.py:52,97-99): (sum_m u_m K_m)**degree,
 * then normalize_K when normalize=1. */
int kmg_combine_host(const double* const* Ks, int p, int64_t n, const double* u, int degree, int normalize, double* out);
/* ALIGNF.__init__ Gram side (ALIGNF.py:28-29,36-58): sub-block K[idx][:,idx], centre, a_i and M_ij. */
int kmg_alignf_stats_host(const double* const* Ks, int p, int64_t n, const int64_t* idx, int64_t nfit, const double* y,
                          double* a, double* M);
/* NLCK.grad (NLCKernels.py:61-66) on the fit sub-blocks (nfit x nfit, contiguous). */
int kmg_nlck_grad_host(const double* const* Ks_fit, int p, int64_t nfit, const double* u, const double* alpha, int degree,
                       double* grad);

/* ---- fused ALIGNF / NLCK entry points: sequences in, statistics / combination out ------------ */
/* One kernel of a method list -- what select_method's mini-language names (kernels.py:461-505). */
#define KMG_KIND_SP 0   /* SP_k{k}                    get_spectrum_K, unnormalised */
#define KMG_KIND_MM 1   /* MM_k{k}_m{m}               get_mismatch_K, normalised as the reference returns it */
#define KMG_KIND_WD 2   /* WD_d{d} */
#define KMG_KIND_WDS 3  /* WDS_d{d}_s{S} */
#define KMG_KIND_LA 4   /* LA_e{e}_d{dd}_b{beta}_smith{smith}: the intended recursion (see kmg_la_host) */
typedef struct kmg_method {
    int32_t kind, k, m, d, S, smith;
    double e, dd, beta;
} kmg_method_t;
/* ALIGNF.__init__ Gram side (ALIGNF.py:28-29,36-58) from SEQUENCES: for every method the fit sub-block K[idx][:, idx] is
 * built on the device straight from the gathered sequences -- the producing kernel's epilogue emits the centring
 * statistics (row sums) and K y~ -- and a_i = <Kc_i, y y'>_F, M_ij = <Kc_i, Kc_j>_F come back: p + p^2 doubles, no Gram
 * crosses PCIe.  pcie_bytes (nullable, 2 entries): bytes moved host->device / device->host by this call. */
int kmg_alignf_fused_host(const uint8_t* seqs, int64_t n, int L, int seq_format, const kmg_method_t* methods, int p,
                          const int64_t* idx, int64_t nfit, const double* y, double* a, double* M, int64_t* pcie_bytes);
/* ALIGNF.get_K (ALIGNF.py:93; degree 1, normalize_inputs 0, normalize 0) and NLCK.get_K (NLCKernels.py:97-99; every kernel
 * through normalize_K first, then (sum_m u_m K_m)**degree and normalize_K of the result) from SEQUENCES: one Gram launch per
 * method accumulates u_m K_m into the output in its epilogue -- bit-identical to the reference's sum over stacked kernels --
 * and the last one applies the power and the final normalisation; only Km (n x n, row stride ldk) crosses PCIe. */
int kmg_combine_fused_host(const uint8_t* seqs, int64_t n, int L, int seq_format, const kmg_method_t* methods, int p,
                           const double* u, int degree, int normalize_inputs, int normalize, double* Km, int64_t ldk, int64_t* pcie_bytes);
/* The Grams of a method list over the sequences idx[0..nsel) (idx NULL: all n) as DEVICE-resident nsel x nsel fp64 matrices
 * in the caller's buffers d_out[0..p) (kmg_dev_malloc): NLCK's fit sub-blocks (NLCKernels.py:33,36) without an upload;
 * normalize_inputs = 1 applies normalize_K to every kernel as NLCK.normalize_kernels does (NLCKernels.py:43-48). */
int kmg_build_grams_dev(const uint8_t* seqs, int64_t n, int L, int seq_format, const kmg_method_t* methods, int p,
                        const int64_t* idx, int64_t nsel, int normalize_inputs, void* const* d_out);

/* ---- device-pointer entry points (block-row construction) ---------------------------------- */
/* letter_to_num/format (kernels.py:178-193): n x L bytes -> 8 u32 words of bit-planes per sequence.
 * *d_err_flag (device int, zero-initialised by the caller) is set to 1 on a non-ACGT byte. */
int kmg_pack_dev(const uint8_t* d_seqs, int seq_format, int64_t n, int L, uint32_t* d_planes, int* d_err_flag, void* stream);
/* padded feature width of the concatenated spectrum feature map (multiple of 128). */
int64_t kmg_spectrum_phi_width(const int* ks, int nk);
/* get_phi_u (kernels.py:12-25) for every sequence and every k in ks: int8 Phi, n x ld_phi. */
int kmg_spectrum_phi_dev(const uint32_t* d_planes, int64_t n, int L, const int* ks, int nk, int8_t* d_phi, int64_t ld_phi,
                         void* stream);
/* get_phi_km (kernels.py:161-175): dense (k,m)-mismatch feature map, int8, width pad128(4^k), k <= 8, m <= 3. */
int kmg_mismatch_phi_dev(const uint32_t* d_planes, int64_t n, int L, int k, int m, int8_t* d_phi, int64_t ld_phi, void* stream);
/* sd[i] = sqrt(sum_t Phi[i][t]^2) -- the diagonal a normalised spectrum Gram needs. */
int kmg_phi_diag_sqrt_dev(const int8_t* d_phi, int64_t n, int64_t width, int64_t ld_phi, double* d_sd, void* stream);
/* K block = Phi_rows Phi_cols^T on the tensor cores (kernels.py:41-45).  symmetric=1: square diagonal
 * block written in place with mirror stores.  sd_rows/sd_cols (nullable): fused cosine normalisation.
 * m_sub: 0 auto, 1 = 128x256 tiles, 2 = 256x256 tiles. */
int kmg_gram_i8_dev(const int8_t* d_phi_rows, const int8_t* d_phi_cols, int64_t rows, int64_t cols, int64_t width,
                    int64_t ld_phi, int64_t row_index0, int64_t col_index0, void* d_out, int64_t ldo, int out_dtype,
                    int symmetric, const double* d_sd_rows, const double* d_sd_cols, int m_sub, void* stream);
/* Symmetric Gram of ALL n rows of one Phi, cut into n_parts block-rows (one per GPU of a node; part_row0: n_parts+1
 * boundaries, multiples of 256, last = n).  Part `part` computes about half of its block-row -- the blocks at cyclic
 * distance < n_parts/2 plus the upper triangle of its diagonal block -- and delivers every tile twice: into
 * part_out[part] and, transposed, into the block-row buffer of the part that owns the tile's columns (part_out[b]: peer
 * device memory, e.g. from kmg_ipc_open).  After all parts have run (stream sync + barrier) every buffer holds its full
 * rows_p x n block-row: the mirror of kernels.py:45 is the one exchange of the path.
 * exchange: KMG_EXCH_DIRECT -- one GEMM launch per peer block whose epilogue stores every transposed piece through the TMA
 * engine (cp.async.bulk.tensor) straight into the owner's block-row: compute and exchange fused tile by tile;
 * KMG_EXCH_STAGED -- the transposed block goes to d_stage (kmg_gram_sharded_stage_bytes bytes of local staging) and one
 * pitched peer copy per block ships it while the next launch runs; KMG_EXCH_SINGLE -- a single launch whose threads store
 * into the peers' buffers (buffers on one device, or tests).  d_stage is only read for KMG_EXCH_STAGED.
 * d_sd (nullable): sqrt(diag) of all n rows, fused cosine normalisation.  computed_entries (nullable out). */
int kmg_gram_sharded_stage_bytes(int n_parts, const int64_t* part_row0, int part, int out_dtype, int64_t* bytes);
int kmg_gram_sharded_join(void* stream);
/* kmg_gram_sharded_mark(slot): remember "every peer copy enqueued so far" (slot 0..3, per host thread);
 * kmg_gram_sharded_wait_mark(slot, stream): `stream` waits for exactly that point.  With two sets of block-row buffers the
 * next build starts while the previous build's copies drain, and the build after it waits for the right copies only. */
int kmg_gram_sharded_mark(int slot);
int kmg_gram_sharded_wait_mark(int slot, void* stream);
/* GEMM launches one call of kmg_gram_i8_sharded_dev enqueues for this part (host utility, no GPU) */
int kmg_gram_sharded_launches(int n_parts, const int64_t* part_row0, int part, int exchange, int* launches);
int kmg_gram_i8_sharded_dev(const int8_t* d_phi, int64_t n, int64_t width, int64_t ld_phi, int n_parts, int part,
                            const int64_t* part_row0, void* const* part_out, int64_t ldo, int out_dtype, const double* d_sd,
                            int exchange, void* d_stage, int64_t* computed_entries, void* stream);
/* Diagnostic: the int8 tensor-core peak of this GPU as this library can drive it -- iters x 4 back-to-back
 * tcgen05.mma.cta_group::2.kind::i8 (256 x 256 x 32) per CTA pair, operands resident in shared memory, no loads, no
 * epilogue.  Enqueues only; the caller times the stream.  *ops = int8 operations issued.  bench.py's roofline
 * denominator (MEASURED_PEAKS.json has no int8 entry). */
int kmg_mma_peak_i8_dev(int iters, int64_t* ops, void* stream);
/* Diagnostic: issue peak of the CUDA-core pipe that bounds a pairwise kernel -- kind 0 LOP3 / 1 SHF (INT32 ALU pipe:
 * mismatch, weighted degree), 2 POPC (XU pipe), 3 DFMA / 4 DADD / 5 DMUL (FP64 pipe: local alignment) -- as dependent
 * chains of that one instruction on every SM.  Enqueues only; the caller times the stream.  *ops = thread-level
 * instructions executed.  bench.py's denominators for kernels.* (BASELINE.md section 4 asks for microbenchmarks). */
#define KMG_PEAK_LOP3 0
#define KMG_PEAK_SHF 1
#define KMG_PEAK_POPC 2
#define KMG_PEAK_DFMA 3
#define KMG_PEAK_DADD 4
#define KMG_PEAK_DMUL 5
int kmg_alu_peak_dev(int kind, int iters, int64_t* ops, void* stream);
/* the assignment rule itself (host utility, no GPU): 1 when part a computes tile (I, J) of the global 256-grid whose
 * columns belong to part b; exactly one of a:(I,J) and b:(J,I) is 1 for I != J */
int kmg_gram_sharded_takes_host(int n_parts, const int64_t* part_row0, int a, int b, int64_t I, int64_t J);
/* CUDA IPC handles (64 bytes) of buffers from kmg_dev_malloc: one process per GPU, same node */
int kmg_ipc_export(const void* d_ptr, uint8_t* handle64);
int kmg_ipc_open(const uint8_t* handle64, void** d_ptr);
int kmg_ipc_close(void* d_ptr);
/* same contraction on CUDA cores (dp4a), s32 output: validation only, not a product path. */
int kmg_gram_i8_simt_dev(const int8_t* d_phi_rows, const int8_t* d_phi_cols, int64_t rows, int64_t cols, int64_t width,
                         int64_t ld_phi, int32_t* d_out, int64_t ldo, void* stream);
/* pairwise kernels on bit-planes; rows/cols blocks of the same plane array or of two arrays. */
int kmg_mismatch_dev(const uint32_t* d_planes_rows, const uint32_t* d_planes_cols, int64_t rows, int64_t cols,
                     int64_t row_index0, int64_t col_index0, int L, int k, int m, void* d_out, int64_t ldo, int out_dtype,
                     int symmetric, const double* d_sd_rows, const double* d_sd_cols, void* stream);
int kmg_mismatch_diag_dev(const uint32_t* d_planes, int64_t n, int L, int k, int m, double* d_sd, void* stream);
int kmg_wd_dev(const uint32_t* d_planes_rows, const uint32_t* d_planes_cols, int64_t rows, int64_t cols, int64_t row_index0,
               int64_t col_index0, int L, int d, double* d_out, int64_t ldo, int symmetric, void* stream);
int kmg_wds_dev(const uint32_t* d_planes_rows, const uint32_t* d_planes_cols, int64_t rows, int64_t cols, int64_t row_index0,
                int64_t col_index0, int L, int d, int S, double* d_out, int64_t ldo, int symmetric, void* stream);
int kmg_la_dev(const uint32_t* d_planes_rows, const uint32_t* d_planes_cols, int64_t rows, int64_t cols, int64_t row_index0,
               int64_t col_index0, int L, double e, double d, double beta, int smith, double* d_out, int64_t ldo,
               int symmetric, void* stream);
/* stored-Gram passes */
int kmg_normalize_dev(double* d_K, int64_t n, int64_t ld, double* d_sd_scratch /* n */, void* stream);
int64_t kmg_center_workspace_bytes(int64_t n);
int kmg_center_dev(const double* d_K, int64_t n, int64_t ld, double* d_out, int64_t ldo, void* d_workspace, void* stream);
/* pieces of center_K for a block-row sharded Gram (kmg/dist.py all-reduces the column sums and the grand sum):
 * out = K - cs[j]/n - rs[i]/n + g/n^2 with rs = row sums of the block, cs = column sums over ALL n rows, g = grand sum. */
int kmg_row_sums_dev(const double* d_K, int64_t rows, int64_t cols, int64_t ld, double* d_rs, void* stream);
int64_t kmg_col_sums_workspace_bytes(int64_t rows, int64_t cols);
int kmg_col_sums_dev(const double* d_K, int64_t rows, int64_t cols, int64_t ld, double* d_cs, void* d_workspace, void* stream);
int kmg_center_apply_dev(const double* d_K, int64_t rows, int64_t cols, int64_t n_total, int64_t ld, const double* d_rs,
                         const double* d_cs, const double* d_g, double* d_out, int64_t ldo, void* stream);
int kmg_gather_dev(const double* d_K, int64_t ld, const int64_t* d_idx, int64_t m, double* d_out, int64_t ldo, void* stream);
int kmg_combine_dev(const double* const* d_Ks /* host array of device pointers */, const int64_t* lds, const double* u, int p,
                    int degree, int64_t rows, int64_t cols, double* d_out, int64_t ldo, void* stream);
int kmg_weighted_dot_dev(const double* d_A, int64_t lda, const double* d_B, int64_t ldb, const double* d_w, int64_t n,
                         double* d_partial /* n */, double* d_result /* 1 */, void* stream);

/* ---- closed-form solvers on resident Grams ------------------------------- */
/* out = K v (rows x cols, row stride ld): KLR.IRLS's m = K alpha (KLR.py:37) */
int kmg_matvec_dev(const double* d_K, int64_t rows, int64_t cols, int64_t ld, const double* d_v, double* d_out, void* stream);
/* x = inv(S K S + c I) b with S = diag(d_s) (NULL: identity), blocked Cholesky in fp64 on the device.
 *   KRR.fit  (KRR.py:33):   a = inv(K_fit + lambda n I) y              -> s = NULL, c = lambda n, b = y
 *   KLR.WKRR (KLR.py:41-57): alpha = W^1/2 inv(W^1/2 K W^1/2 + n lambda I) W^1/2 z -> s = sqrt(W), c = n lambda, b = s * z, alpha = s * x
 * d_work: kmg_spd_solve_workspace_bytes(n) bytes.  Synchronises the stream; KMG_ERR_ARG when the matrix is not positive definite. */
int64_t kmg_spd_solve_workspace_bytes(int64_t n);
int kmg_spd_solve_dev(const double* d_K, int64_t n, int64_t ld, const double* d_s, double c, const double* d_b, double* d_x, void* d_work,
                      void* stream);
/* the same from a host Gram: K_fit = K[idx][:, idx] (idx NULL: all of K, nfit = n) is gathered on the host and uploaded */
int kmg_spd_solve_host(const double* K, int64_t n, int64_t ldk, const int64_t* idx, int64_t nfit, const double* s, double c,
                       const double* b, double* x);

/* (k,m)-mismatch common-neighbourhood table T[0..k] (host utility). */
int kmg_mismatch_table_host(int k, int m, int64_t* T);

#ifdef __cplusplus
}
#endif
#endif /* KMG_H */
