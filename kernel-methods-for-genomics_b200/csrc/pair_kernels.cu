// pair_kernels.cu -- pairwise bit-vector Gram kernels: (k,m)-mismatch and weighted degree.
//
// Both work on the bit-plane encoding (kmg_common.cuh): the per-base mismatch vector of two
// sequences is ne = (xl^yl)|(xh^yh), one LOP3 per 32 bases.  One thread owns one (row, col) pair;
// a warp owns 32 consecutive columns of one row (the row operand is warp-uniform), a CTA an
// 8 x 32 tile.  Everything lives in registers; HBM traffic is the 64 B of the two sequences and the
// 8 B result.  Both kernels are integer-issue bound (LOP3/SHF/POPC), not HBM bound.
//
// ---- mismatch (kernels.py:161-217) ------------------------------------------------------------
// The reference materialises phi_km over all 4^k k-mers (kernels.py:161-175) and takes dot
// products.  <phi_km(x), phi_km(y)> = sum_{p,q} T[d_H(x[p:p+k], y[q:q+k])] exactly, with T[delta]
// the number of k-mers within Hamming distance m of both of two k-mers at distance delta
// (T[delta] = 0 for delta > 2m).  Window pairs (p, q) are grouped by diagonal q-p: rotating y's
// planes by r brings diagonal r (and, in the wrapped part, diagonal r-128) under x, the Hamming
// distances of all windows on that diagonal are the length-k sliding sums of ne, computed for all
// 128 positions at once with bit-sliced saturating counters (log2(k) doubling steps), and
// N_delta = popc(counter == delta & valid-window mask).  K_raw = sum_delta T[delta] * N_delta.
//
// ---- weighted degree (kernels.py:53-101) ------------------------------------------------------
// c_k = popc(AND_{t<k} (match >> t)) restricted to positions 1..L-k (position 0 is skipped by the
// reference's range(1, L-k+1), kernels.py:78); acc += beta_k * c_k in fp64 with separate multiply
// and add (kernels.py:80) -- __dmul_rn/__dadd_rn keep nvcc from contracting them into an FMA, which
// would change the last bit.  The diagonal is the closed form of kernels.py:96.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "gram_i8.h"
#include "kmg_common.cuh"
#include "pair_kernels.h"

namespace {

constexpr int TILE_R = 8;
constexpr int TILE_C = 32;

struct OutSpec {
    void* out;
    int64_t ldo;
    void* out_t;
    int64_t ldo_t;
    int dtype;
    const double* sd_rows;
    const double* sd_cols;
    int64_t rows, cols, row_index0, col_index0;
    int symmetric;
    int epi;   // any of the fused steps of `e` active
    EpiOps e;
};

OutSpec make_out(const PairBlock* b) {
    OutSpec o;
    memset(&o.e, 0, sizeof(o.e));
    o.epi = 0;
    if (b->epi != nullptr) { o.e = *b->epi; o.epi = epi_active(o.e) ? 1 : 0; }
    o.out = b->out; o.ldo = b->ldo; o.out_t = b->symmetric ? b->out_t : nullptr; o.ldo_t = b->ldo_t;
    o.dtype = b->out_dtype; o.sd_rows = b->sd_rows; o.sd_cols = b->sd_cols;
    o.rows = b->rows; o.cols = b->cols; o.row_index0 = b->row_index0; o.col_index0 = b->col_index0;
    o.symmetric = b->symmetric;
    return o;
}

// tile classification for a symmetric (square, diagonal) block: 0 = compute, no mirror;
// 1 = compute and mirror; 2 = skip (produced by another tile's mirror store).
__device__ __forceinline__ int tile_class(const OutSpec& o, int64_t r0, int64_t c0) {
    if (!o.symmetric) return 0;
    const int64_t rlo = o.row_index0 + r0, rhi = rlo + TILE_R - 1;
    const int64_t clo = o.col_index0 + c0, chi = clo + TILE_C - 1;
    if (chi < rlo) return 2;
    return clo > rhi ? 1 : 0;
}

__device__ __forceinline__ double finish_int(const OutSpec& o, int64_t r, int64_t c, int64_t raw) {
    double v = (double)raw;
    if (o.sd_rows != nullptr) {
        // normalize_K (kernels.py:408-414)
        v = __ddiv_rn(v, __dmul_rn(o.sd_rows[r], o.sd_cols[c]));
        if (o.row_index0 + r == o.col_index0 + c) v = 1.0;
    }
    return v;
}

// fused ALIGNF / NLCK steps on the value of entry (r, c) (block-local indices): returns what is to be stored
__device__ __forceinline__ double apply_epi(const OutSpec& o, int64_t r, int64_t c, double v) {
    if (!o.epi) return v;
    const double prev = o.e.accumulate == 2 ? reinterpret_cast<const double*>(o.out)[r * o.ldo + c] : 0.0;
    const bool nrm = o.e.post_sd_rows != nullptr;
    return epi_finish(o.e, v, prev, o.row_index0 + r == o.col_index0 + c, nrm ? o.e.post_sd_rows[r] : 1.0, nrm ? o.e.post_sd_cols[c] : 1.0);
}

// cosine normalisation of an fp64 kernel value (weighted degree: the diagonal has already been set)
__device__ __forceinline__ double finish_f64(const OutSpec& o, int64_t r, int64_t c, double v) {
    if (o.sd_rows != nullptr) {
        v = __ddiv_rn(v, __dmul_rn(o.sd_rows[r], o.sd_cols[c]));   // normalize_K (kernels.py:408-414)
        if (o.row_index0 + r == o.col_index0 + c) v = 1.0;
    }
    return v;
}

// returns the fp64 value stored (0 for the s32 output, which takes no fused steps)
template <bool EPI>
__device__ __forceinline__ double store_int(const OutSpec& o, int64_t r, int64_t c, int64_t raw, bool mirror) {
    if (o.dtype == KMG_OUT_S32) {
        reinterpret_cast<int32_t*>(o.out)[r * o.ldo + c] = (int32_t)raw;
        if (mirror) reinterpret_cast<int32_t*>(o.out_t)[c * o.ldo_t + r] = (int32_t)raw;
        return 0.0;
    }
    const double v = EPI ? apply_epi(o, r, c, finish_int(o, r, c, raw)) : finish_int(o, r, c, raw);
    reinterpret_cast<double*>(o.out)[r * o.ldo + c] = v;
    if (mirror) reinterpret_cast<double*>(o.out_t)[c * o.ldo_t + r] = v;
    return v;
}

// ------------------------------------------------------------------------------------------
// bit-sliced saturating counters: B bit-planes x 128 positions
// ------------------------------------------------------------------------------------------
template <int B>
struct Cnt {
    uint32_t b[B][4];
};

template <int B>
__device__ __forceinline__ void cnt_zero(Cnt<B>& c) {
#pragma unroll
    for (int i = 0; i < B; ++i)
#pragma unroll
        for (int w = 0; w < 4; ++w) c.b[i][w] = 0u;
}

// r = min(a + b, 2^B - 1) per position
template <int B>
__device__ __forceinline__ void cnt_add_sat(const Cnt<B>& a, const Cnt<B>& b, Cnt<B>& r) {
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        uint32_t c = 0u, s[B];
#pragma unroll
        for (int i = 0; i < B; ++i) {
            const uint32_t x = a.b[i][w], y = b.b[i][w];
            s[i] = x ^ y ^ c;
            c = (x & y) | (c & (x ^ y));
        }
#pragma unroll
        for (int i = 0; i < B; ++i) r.b[i][w] = s[i] | c;
    }
}

// out = in >> s (every plane, 128-bit logical shift; s is a compile-time constant after unrolling)
template <int B>
__device__ __forceinline__ void cnt_shr(const Cnt<B>& in, int s, Cnt<B>& out) {
    const int ws = s >> 5, bs = s & 31;
#pragma unroll
    for (int i = 0; i < B; ++i)
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const uint32_t lo = (w + ws < 4) ? in.b[i][(w + ws) & 3] : 0u;
            const uint32_t hi = (w + ws + 1 < 4) ? in.b[i][(w + ws + 1) & 3] : 0u;
            out.b[i][w] = bs == 0 ? lo : __funnelshift_r(lo, hi, bs);
        }
}

// sliding sums  S[p] = min(sum_{t<K} e[p+t], 2^B-1), compile-time K: doubling over powers of two.
template <int K, int B>
__device__ __forceinline__ void window_sum(const uint32_t (&e)[4], int /*k*/, Cnt<B>& res) {
    Cnt<B> cur;  // sliding sum of width 2^j
    cnt_zero(cur);
#pragma unroll
    for (int w = 0; w < 4; ++w) cur.b[0][w] = e[w];
    cnt_zero(res);
    int off = 0;
#pragma unroll
    for (int j = 0; (1 << j) <= K; ++j) {
        if ((K >> j) & 1) {
            Cnt<B> sh, tmp;
            cnt_shr(cur, off, sh);
            if (off == 0) {
                res = sh;
            } else {
                cnt_add_sat(res, sh, tmp);
                res = tmp;
            }
            off += 1 << j;
        }
        if ((2 << j) <= K) {
            Cnt<B> sh, nxt;
            cnt_shr(cur, 1 << j, sh);
            cnt_add_sat(cur, sh, nxt);
            cur = nxt;
        }
    }
}

// runtime k: k-1 single-bit shifts and saturating increments
template <int B>
__device__ __forceinline__ void window_sum_rt(const uint32_t (&e)[4], int k, Cnt<B>& res) {
    uint32_t sh[4] = {e[0], e[1], e[2], e[3]};
    cnt_zero(res);
#pragma unroll 1
    for (int t = 0; t < k; ++t) {
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            uint32_t c = sh[w], s[B];
#pragma unroll
            for (int i = 0; i < B; ++i) {
                s[i] = res.b[i][w] ^ c;
                c = res.b[i][w] & c;
            }
#pragma unroll
            for (int i = 0; i < B; ++i) res.b[i][w] = s[i] | c;
        }
        kmg_shr1_128(sh);
    }
}

struct MismatchParams {
    int k;
    int L;
    int32_t T[8];      // T[delta], zero beyond 2m
    uint32_t wrap[4];  // bit L-1: where bit 0 re-enters when y is rotated inside its L bits
};

// raw mismatch kernel value of one pair.  y is rotated inside its own L bits (not 128): rotation r puts y[(p+r) mod L]
// under x[p], so one pass serves diagonal +r (windows p <= L-k-r) and diagonal r-L (windows p >= L-r) and L rotations
// cover all (L-k+1)^2 window pairs -- 101 instead of 128 passes at L = 101.  vmask[r] = valid window starts (shared).
// NW = words that can hold a window start, ((L-k) >> 5) + 1: the count loop reads only those, and since everything is
// unrolled straight-line code the compiler drops the upstream shifts / adds of the unused words with them.
template <int K, int B, int NW>
__device__ __forceinline__ int64_t mismatch_pair(const SeqPlanes& x, SeqPlanes y, const MismatchParams& mp,
                                                 const uint32_t (*vmask)[4]) {
    constexpr int ND = (1 << B) - 1;  // distances 0..ND-1 are exact, ND-1 >= 2m
    int32_t N[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) N[d] = 0;
    const uint32_t wrap[4] = {mp.wrap[0], mp.wrap[1], mp.wrap[2], mp.wrap[3]};
#pragma unroll 1
    for (int r = 0; r < mp.L; ++r) {
        uint32_t vm[4];
#pragma unroll
        for (int w = 0; w < 4; ++w) vm[w] = w < NW ? vmask[r][w] : 0u;
        if ((vm[0] | vm[1] | vm[2] | vm[3]) != 0u) {  // block-uniform
            uint32_t e[4];
#pragma unroll
            for (int w = 0; w < 4; ++w) e[w] = (x.lo[w] ^ y.lo[w]) | (x.hi[w] ^ y.hi[w]);
            Cnt<B> s;
            if (K > 0) window_sum<(K > 0 ? K : 1), B>(e, mp.k, s);
            else window_sum_rt<B>(e, mp.k, s);
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                int cnt = 0;
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    uint32_t eq = vm[w];
#pragma unroll
                    for (int i = 0; i < B; ++i) eq &= ((d >> i) & 1) ? s.b[i][w] : ~s.b[i][w];
                    cnt += __popc(eq);
                }
                N[d] += cnt;
            }
        }
        kmg_rotr1_len(y.lo, wrap);
        kmg_rotr1_len(y.hi, wrap);
    }
    int64_t acc = 0;
#pragma unroll
    for (int d = 0; d < ND; ++d) acc += (int64_t)mp.T[d] * (int64_t)N[d];
    return acc;
}

// valid window-start masks per rotation r (period L): a window (p, k) under rotation r pairs x[p..p+k) with
// y[p+r..p+r+k) when p + r + k <= L (diagonal +r: p in [0, L-k-r]) and with y[p+r-L..) when p >= L - r (diagonal r-L:
// p in [L-r, L-k], non-empty iff r >= k); windows that straddle the wrap point are invalid.
__device__ __forceinline__ void build_vmask(uint32_t (*vmask)[4], int L, int k) {
    for (int r = threadIdx.x; r < 128; r += blockDim.x) {
        uint32_t m[4] = {0u, 0u, 0u, 0u};
        if (r < L) {
            if (L - k - r >= 0) kmg_range_mask_128(0, L - k - r, m);
            if (r >= 1 && r >= k) kmg_range_mask_128(L - r, L - k, m);
        }
        vmask[r][0] = m[0]; vmask[r][1] = m[1]; vmask[r][2] = m[2]; vmask[r][3] = m[3];
    }
}

// EPI: the fused ALIGNF / NLCK steps (epi_ops.cuh) are compiled into a separate instantiation: carrying them as dead code
// costs the plain weighted-degree kernel 12 % (measured, same box), so the plain kernels carry none of it.
template <int K, int B, int NW, bool EPI>
__global__ void __launch_bounds__(TILE_R * TILE_C)
mismatch_kernel(const uint32_t* __restrict__ prow, const uint32_t* __restrict__ pcol, const OutSpec o,
                const MismatchParams mp) {
    __shared__ uint32_t vmask[128][4];
    const int64_t r0 = (int64_t)blockIdx.y * TILE_R, c0 = (int64_t)blockIdx.x * TILE_C;
    const int cls = tile_class(o, r0, c0);
    if (cls == 2) return;
    build_vmask(vmask, mp.L, mp.k);
    __syncthreads();
    const int64_t r = r0 + (threadIdx.x >> 5), c = c0 + (threadIdx.x & 31);
    const bool live = r < o.rows && c < o.cols;
    const SeqPlanes x = kmg_load_planes(prow, live ? r : 0);
    const SeqPlanes y = kmg_load_planes(pcol, live ? c : 0);
    const int64_t raw = mismatch_pair<K, B, NW>(x, y, mp, vmask);
    if (!EPI) {
        if (live) store_int<false>(o, r, c, raw, cls == 1);
        return;
    }
    double v = 0.0;
    if (live) v = store_int<true>(o, r, c, raw, cls == 1);
    if (r < o.rows) epi_row_partial_warp(o.e, r, c0, v, live, threadIdx.x & 31);  // warp-uniform: r, c0
}

template <int K, int B, int NW>
__global__ void __launch_bounds__(256)
mismatch_diag_kernel(const uint32_t* __restrict__ planes, int64_t n, const MismatchParams mp, double* __restrict__ sd) {
    __shared__ uint32_t vmask[128][4];
    build_vmask(vmask, mp.L, mp.k);
    __syncthreads();
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const SeqPlanes x = kmg_load_planes(planes, i < n ? i : 0);
    const int64_t raw = mismatch_pair<K, B, NW>(x, x, mp, vmask);
    if (i < n) sd[i] = sqrt((double)raw);  // np.sqrt(np.diag(K)), kernels.py:408 (IEEE, correctly rounded)
}

// ------------------------------------------------------------------------------------------
// weighted degree
// ------------------------------------------------------------------------------------------
struct WdParams {
    int d;
    int L;
    uint32_t posmask[4]; // bits 1..L-1 (position 0 is skipped by the reference, kernels.py:78)
    double diag;        // L - 1 + (1 - d) / 3, evaluated on the host like kernels.py:96
    double beta[128];   // beta_k = 2*(d-k+1)/d/(d+1), evaluated on the host like kernels.py:61
};

// Each thread owns WD_PPT pairs (one row, columns c, c+32, c+64, c+96): the index arithmetic, the row operand, the
// loop control, the warp vote and the beta_k load are shared (the kernel is issue bound: ncu 87 % of issue slots at
// one pair per thread, a third of them outside the k loop).  CTA tile 8 rows x 128 columns.
constexpr int WD_PPT = 4;
constexpr int WD_TILE_C = TILE_C * WD_PPT;


// One k of the weighted-degree sum for the WD_PPT pairs of a thread: m &= (match >> (k-1)), c_k = popc(m),
// acc += beta_k * c_k (kernels.py:74-80).  NW = words of the 128-bit vectors that can still be non-zero.
// Returns true when no pair of the warp has a run of length k left (the remaining terms add +0.0).
// Measured alternatives, same box, d = 10, 32768^2 (tools/ab_pair.sh): this form 1.285e11 entries/s; the run vector as its
// own shifted operand (m_k = m_{k-1} & (m_{k-1} >> 1): no second array, 16 registers less, but shift and AND become one
// serial chain) 1.135e11; voting on the population counts instead of OR-ing the run-vector words 1.19e11, the same
// software-pipelined on the previous step's counts 1.11e11; retiring the four 32-pair slices one by one 1.09e11.
template <int NW>
__device__ __forceinline__ bool wd_step(uint32_t (&m)[WD_PPT][4], uint32_t (&sh)[WD_PPT][4], double (&acc)[WD_PPT], int k, double bk) {
    uint32_t any = 0u;
#pragma unroll
    for (int j = 0; j < WD_PPT; ++j) {
        if (k > 1) {
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const uint32_t hi = (w + 1 < NW) ? sh[j][w + 1 < 4 ? w + 1 : 3] : 0u;  // words >= NW are zero by now
                sh[j][w] = __funnelshift_r(sh[j][w], hi, 1);
                m[j][w] &= sh[j][w];
            }
        }
#pragma unroll
        for (int w = 0; w < NW; ++w) any |= m[j][w];
    }
    if (__all_sync(0xffffffffu, any == 0u)) return true;
#pragma unroll
    for (int j = 0; j < WD_PPT; ++j) {
        // The XU pipe (POPC, I2F) is the busiest one after the ALU: a carry-save adder over three of the four words
        // trades one POPC for two LOP3, and the exact int -> double conversion is one FP64 add of 2^52.
        const uint32_t s3 = m[j][0] ^ m[j][1] ^ m[j][2];
        const uint32_t cy = (m[j][0] & m[j][1]) | (m[j][2] & (m[j][0] ^ m[j][1]));
        int cnt = __popc(s3) + 2 * __popc(cy);
        if (NW == 4) cnt += __popc(m[j][3]);
        const double cd = __hiloint2double(0x43300000, cnt) - 4503599627370496.0;
        acc[j] = __dadd_rn(acc[j], __dmul_rn(bk, cd));
    }
    return false;
}

template <bool EPI>
__global__ void __launch_bounds__(TILE_R * TILE_C)
wd_kernel(const uint32_t* __restrict__ prow, const uint32_t* __restrict__ pcol, const OutSpec o, const WdParams wp) {
    __shared__ double tile[TILE_R][WD_TILE_C + 1];
    const int64_t r0 = (int64_t)blockIdx.y * TILE_R, c0 = (int64_t)blockIdx.x * WD_TILE_C;
    int cls = 0;
    if (o.symmetric) {
        const int64_t rlo = o.row_index0 + r0, rhi = rlo + TILE_R - 1;
        const int64_t clo = o.col_index0 + c0, chi = clo + WD_TILE_C - 1;
        if (chi < rlo) return;          // produced by another tile's mirror store
        cls = clo > rhi ? 1 : 0;        // strictly above the diagonal: mirror
    }
    const int tr = threadIdx.x >> 5, tc = threadIdx.x & 31;
    const int64_t r = r0 + tr;
    const bool row_ok = r < o.rows;
    const SeqPlanes x = kmg_load_planes(prow, row_ok ? r : 0);
    uint32_t m[WD_PPT][4], sh[WD_PPT][4];
#pragma unroll
    for (int j = 0; j < WD_PPT; ++j) {
        const int64_t c = c0 + tc + 32 * j;
        const SeqPlanes y = kmg_load_planes(pcol, c < o.cols ? c : 0);
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            m[j][w] = ~((x.lo[w] ^ y.lo[w]) | (x.hi[w] ^ y.hi[w])) & wp.posmask[w];  // positions 1..L-1
            sh[j][w] = m[j][w];
        }
    }
    double acc[WD_PPT];
#pragma unroll
    for (int j = 0; j < WD_PPT; ++j) acc[j] = 0.0;
    // The shifted match vector of iteration k is match >> (k-1), bits <= L - k: after the iteration k = L - 95 its fourth
    // word is empty (and so is the run vector's), so the later iterations work on three words (L = 101: k >= 7, a quarter of the shift / and / popc work of those iterations).
    const int k4 = min(wp.d, wp.L - 95);  // last k done on all four words (<= 0: none)
    bool done = false;
#pragma unroll 1
    for (int k = 1; k <= k4; ++k) {
        if (wd_step<4>(m, sh, acc, k, wp.beta[k - 1])) { done = true; break; }
    }
    if (!done) {
#pragma unroll 1
        for (int k = max(k4, 0) + 1; k <= wp.d; ++k) {
            if (wd_step<3>(m, sh, acc, k, wp.beta[k - 1])) break;
        }
    }
#pragma unroll
    for (int j = 0; j < WD_PPT; ++j) {
        const int64_t c = c0 + tc + 32 * j;
        const bool live = row_ok && c < o.cols;
        if (o.row_index0 + r == o.col_index0 + c) acc[j] = wp.diag;
        if (EPI && live) acc[j] = apply_epi(o, r, c, finish_f64(o, r, c, acc[j]));
        if (live) reinterpret_cast<double*>(o.out)[r * o.ldo + c] = acc[j];
        if (cls == 1) tile[tr][tc + 32 * j] = acc[j];
        if (EPI && o.epi && row_ok && c0 + 32 * j < o.cols) epi_row_partial_warp(o.e, r, c0 + 32 * j, acc[j], live, tc);  // warp-uniform test
    }
    if (cls == 1) {  // mirror through shared memory so each column receives 8 consecutive doubles
        __syncthreads();
#pragma unroll
        for (int j = 0; j < WD_PPT; ++j) {
            const int mc = (threadIdx.x >> 3) + 32 * j, mr = threadIdx.x & 7;
            if (r0 + mr < o.rows && c0 + mc < o.cols)
                reinterpret_cast<double*>(o.out_t)[(c0 + mc) * o.ldo_t + r0 + mr] = tile[mr][mc];
        }
    }
}

// ------------------------------------------------------------------------------------------
// weighted degree with shifts (kernels.py:106-155)
//   c_t = sum_k beta_k * c_st(k),   c_st(k) = sum_{i=1}^{L-k} sum_{s=0}^{S} [s+i<L] delta_s *
//            ([x[i+s:i+s+k]==y[i:i+k]] + [x[i:i+k]==y[i+s:i+s+k]]),   delta_s = 1/2/(s+1)
// The reference accumulates c_st sequentially over (i, s) in fp64; delta_2 = 1/6 is inexact, so
// the order matters for bit-exactness.  Zero terms add +0.0 (a no-op), so only the set bits of the
// run vectors are visited, in ascending i then ascending s.
// ------------------------------------------------------------------------------------------
struct WdsParams {
    int d, S, L;
    double beta[128];
    double delta[8];
};

template <int SMAX, bool EPI>
__global__ void __launch_bounds__(TILE_R * TILE_C)
wds_kernel(const uint32_t* __restrict__ prow, const uint32_t* __restrict__ pcol, const OutSpec o, const WdsParams wp) {
    __shared__ double tile[TILE_R][TILE_C + 1];
    const int64_t r0 = (int64_t)blockIdx.y * TILE_R, c0 = (int64_t)blockIdx.x * TILE_C;
    const int cls = tile_class(o, r0, c0);
    if (cls == 2) return;
    const int tr = threadIdx.x >> 5, tc = threadIdx.x & 31;
    const int64_t r = r0 + tr, c = c0 + tc;
    const bool live = r < o.rows && c < o.cols;
    const SeqPlanes x = kmg_load_planes(prow, live ? r : 0);
    const SeqPlanes y = kmg_load_planes(pcol, live ? c : 0);
    uint32_t ca[SMAX + 1][4], cb[SMAX + 1][4], sa[SMAX + 1][4], sb[SMAX + 1][4];
    {
        uint32_t xl[4] = {x.lo[0], x.lo[1], x.lo[2], x.lo[3]}, xh[4] = {x.hi[0], x.hi[1], x.hi[2], x.hi[3]};
        uint32_t yl[4] = {y.lo[0], y.lo[1], y.lo[2], y.lo[3]}, yh[4] = {y.hi[0], y.hi[1], y.hi[2], y.hi[3]};
#pragma unroll
        for (int s = 0; s <= SMAX; ++s) {
            uint32_t vm[4] = {0u, 0u, 0u, 0u};
            if (s <= wp.S && wp.L - 1 - s >= 0) kmg_range_mask_128(0, wp.L - 1 - s, vm);
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                ca[s][w] = ~((xl[w] ^ y.lo[w]) | (xh[w] ^ y.hi[w])) & vm[w];  // bit i: x[i+s] == y[i]
                cb[s][w] = ~((x.lo[w] ^ yl[w]) | (x.hi[w] ^ yh[w])) & vm[w];  // bit i: x[i] == y[i+s]
                sa[s][w] = ca[s][w];
                sb[s][w] = cb[s][w];
            }
            kmg_shr1_128(xl); kmg_shr1_128(xh); kmg_shr1_128(yl); kmg_shr1_128(yh);
        }
    }
    double c_t = 0.0;
#pragma unroll 1
    for (int k = 1; k <= wp.d; ++k) {
        if (k > 1) {
#pragma unroll
            for (int s = 0; s <= SMAX; ++s) {
                kmg_shr1_128(sa[s]);
                kmg_shr1_128(sb[s]);
#pragma unroll
                for (int w = 0; w < 4; ++w) { ca[s][w] &= sa[s][w]; cb[s][w] &= sb[s][w]; }
            }
        }
        uint32_t st[4] = {0u, 0u, 0u, 0u};
        if (wp.L - k >= 1) kmg_range_mask_128(1, wp.L - k, st);  // window starts i = 1 .. L-k
        double c_st = 0.0;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            uint32_t u = 0u;
#pragma unroll
            for (int s = 0; s <= SMAX; ++s) u |= ca[s][w] | cb[s][w];
            u &= st[w];
            while (u) {
                const int bit = __ffs(u) - 1;
                u &= u - 1;
#pragma unroll
                for (int s = 0; s <= SMAX; ++s) {
                    const int m = (int)((ca[s][w] >> bit) & 1u) + (int)((cb[s][w] >> bit) & 1u);
                    if (m) c_st = __dadd_rn(c_st, __dmul_rn(wp.delta[s], (double)m));
                }
            }
        }
        c_t = __dadd_rn(c_t, __dmul_rn(wp.beta[k - 1], c_st));
    }
    if (EPI && live) c_t = apply_epi(o, r, c, finish_f64(o, r, c, c_t));
    if (live) reinterpret_cast<double*>(o.out)[r * o.ldo + c] = c_t;
    if (EPI && o.epi && r < o.rows) epi_row_partial_warp(o.e, r, c0, c_t, live, tc);
    if (cls == 1) {
        tile[tr][tc] = c_t;
        __syncthreads();
        const int mc = threadIdx.x >> 3, mr = threadIdx.x & 7;
        if (r0 + mr < o.rows && c0 + mc < o.cols)
            reinterpret_cast<double*>(o.out_t)[(c0 + mc) * o.ldo_t + r0 + mr] = tile[mr][mc];
    }
}

int check_block(const PairBlock* b) {
    KMG_REQUIRE(b->rows >= 0 && b->cols >= 0, KMG_ERR_ARG, "negative block shape");
    KMG_REQUIRE(b->L >= 1 && b->L <= KMG_MAX_L, KMG_ERR_UNSUPPORTED, "sequence length %d not supported (1..%d)", b->L, KMG_MAX_L);
    KMG_REQUIRE((b->rows + TILE_R - 1) / TILE_R <= 65535, KMG_ERR_ARG, "row block too tall (max %d rows per call)", 65535 * TILE_R);
    if (b->symmetric) {
        KMG_REQUIRE(b->rows == b->cols && b->row_index0 == b->col_index0 && b->out_t != nullptr, KMG_ERR_ARG,
                    "symmetric mode needs a square diagonal block and a mirror destination");
    }
    if (b->epi != nullptr && epi_active(*b->epi)) {
        KMG_REQUIRE(b->out_dtype == KMG_OUT_F64, KMG_ERR_ARG, "fused epilogue steps need the fp64 output");
        KMG_REQUIRE(!(b->epi->row_sum_partial || b->epi->row_wsum_partial) || (!b->symmetric && b->epi->n_chunks >= (b->cols + 31) / 32), KMG_ERR_ARG,
                    "row partial sums need a plain (non-symmetric) block and n_chunks >= ceil(cols / 32)");
        KMG_REQUIRE(!b->epi->row_wsum_partial || b->epi->w_cols, KMG_ERR_ARG, "weighted row partial sums need w_cols");
        KMG_REQUIRE((b->epi->post_sd_rows == nullptr) == (b->epi->post_sd_cols == nullptr), KMG_ERR_ARG, "post_sd_rows and post_sd_cols go together");
    }
    return KMG_OK;
}

// words that can hold a window start: 3 when L - k + 1 <= 96 (the 101-bp data of the reference at every k >= 6), else 4
inline int mm_result_words(const MismatchParams& mp) { return ((mp.L - mp.k) >> 5) + 1 <= 3 ? 3 : 4; }

template <int B, int NW>
int launch_mismatch_k(const PairBlock* b, const OutSpec& o, const MismatchParams& mp, cudaStream_t stream) {
    dim3 grid((unsigned)((b->cols + TILE_C - 1) / TILE_C), (unsigned)((b->rows + TILE_R - 1) / TILE_R));
    dim3 block(TILE_R * TILE_C);
    if (o.epi) {  // fused ALIGNF / NLCK steps: the runtime-k variant only (these calls are small)
        mismatch_kernel<0, B, NW, true><<<grid, block, 0, stream>>>(b->planes_rows, b->planes_cols, o, mp);
        KMG_CUDA_CHECK(cudaGetLastError());
        return KMG_OK;
    }
#define KMG_MM_CASE(KK) \
    case KK: mismatch_kernel<KK, B, NW, false><<<grid, block, 0, stream>>>(b->planes_rows, b->planes_cols, o, mp); break;
    switch (mp.k) {
        KMG_MM_CASE(1) KMG_MM_CASE(2) KMG_MM_CASE(3) KMG_MM_CASE(4) KMG_MM_CASE(5) KMG_MM_CASE(6) KMG_MM_CASE(7)
        KMG_MM_CASE(8) KMG_MM_CASE(9) KMG_MM_CASE(10) KMG_MM_CASE(11) KMG_MM_CASE(12) KMG_MM_CASE(13) KMG_MM_CASE(14)
        KMG_MM_CASE(15) KMG_MM_CASE(16)
        default: mismatch_kernel<0, B, NW, false><<<grid, block, 0, stream>>>(b->planes_rows, b->planes_cols, o, mp); break;
    }
#undef KMG_MM_CASE
    KMG_CUDA_CHECK(cudaGetLastError());
    return KMG_OK;
}

template <int B, int NW>
int launch_mismatch_diag_k(const uint32_t* planes, int64_t n, const MismatchParams& mp, double* sd, cudaStream_t stream) {
    const unsigned grid = (unsigned)((n + 255) / 256);
#define KMG_MM_CASE(KK) \
    case KK: mismatch_diag_kernel<KK, B, NW><<<grid, 256, 0, stream>>>(planes, n, mp, sd); break;
    switch (mp.k) {
        KMG_MM_CASE(1) KMG_MM_CASE(2) KMG_MM_CASE(3) KMG_MM_CASE(4) KMG_MM_CASE(5) KMG_MM_CASE(6) KMG_MM_CASE(7)
        KMG_MM_CASE(8) KMG_MM_CASE(9) KMG_MM_CASE(10) KMG_MM_CASE(11) KMG_MM_CASE(12) KMG_MM_CASE(13) KMG_MM_CASE(14)
        KMG_MM_CASE(15) KMG_MM_CASE(16)
        default: mismatch_diag_kernel<0, B, NW><<<grid, 256, 0, stream>>>(planes, n, mp, sd); break;
    }
#undef KMG_MM_CASE
    KMG_CUDA_CHECK(cudaGetLastError());
    return KMG_OK;
}

int fill_mismatch_params(int k, int m, int L, MismatchParams* mp, int* bits) {
    KMG_REQUIRE(k >= 1 && k <= L, KMG_ERR_ARG, "mismatch: need 1 <= k <= L (k=%d, L=%d)", k, L);
    KMG_REQUIRE(m >= 0 && m <= KMG_MM_MAX_M, KMG_ERR_UNSUPPORTED, "mismatch: m=%d not supported (0..%d)", m, KMG_MM_MAX_M);
    int64_t T[KMG_MAX_L + 1];
    kmg_mismatch_table(k, m, T);
    mp->k = k; mp->L = L;
    for (int d = 0; d < 8; ++d) mp->T[d] = 0;
    for (int d = 0; d <= k && d <= 2 * m; ++d) {
        KMG_REQUIRE(T[d] < (1ll << 31), KMG_ERR_UNSUPPORTED, "mismatch: neighbourhood table overflows int32");
        mp->T[d] = (int32_t)T[d];
    }
    for (int w = 0; w < 4; ++w) mp->wrap[w] = (w == ((L - 1) >> 5)) ? (1u << ((L - 1) & 31)) : 0u;
    *bits = (m == 0) ? 1 : (m == 1 ? 2 : 3);
    return KMG_OK;
}

}  // namespace

// T[delta] = #{b in {A,C,G,T}^k : d_H(u,b) <= m and d_H(v,b) <= m} for d_H(u,v) = delta.  On the
// delta differing positions b may agree with u (a of them), with v (bb) or with neither (c, 2
// choices each); on the k-delta common positions it differs in t places (3 choices each).
int kmg_mismatch_table(int k, int m, int64_t* T) {
    auto binom = [](int n, int r) -> int64_t {
        if (r < 0 || r > n) return 0;
        int64_t v = 1;
        for (int i = 1; i <= r; ++i) v = v * (n - r + i) / i;
        return v;
    };
    auto ipow = [](int64_t b, int e) -> int64_t { int64_t v = 1; while (e-- > 0) v *= b; return v; };
    for (int delta = 0; delta <= k; ++delta) {
        int64_t tot = 0;
        for (int a = 0; a <= delta; ++a)
            for (int bb = 0; a + bb <= delta; ++bb) {
                const int c = delta - a - bb;
                const int64_t multi = binom(delta, a) * binom(delta - a, bb);
                for (int t = 0; t <= k - delta; ++t)
                    if (bb + c + t <= m && a + c + t <= m) tot += multi * ipow(2, c) * binom(k - delta, t) * ipow(3, t);
            }
        T[delta] = tot;
    }
    return KMG_OK;
}

int kmg_mismatch_launch(const PairBlock* b, int k, int m, cudaStream_t stream) {
    int rc = check_block(b);
    if (rc) return rc;
    MismatchParams mp;
    int bits = 0;
    rc = fill_mismatch_params(k, m, b->L, &mp, &bits);
    if (rc) return rc;
    const int64_t W = b->L - k + 1;
    KMG_REQUIRE(b->out_dtype == KMG_OUT_F64 || W * W * (int64_t)mp.T[0] < (1ll << 31), KMG_ERR_UNSUPPORTED,
                "mismatch: raw values may overflow int32, use the f64 output");
    if (b->rows == 0 || b->cols == 0) return KMG_OK;
    const OutSpec o = make_out(b);
    if (mm_result_words(mp) == 3) {
        if (bits == 1) return launch_mismatch_k<1, 3>(b, o, mp, stream);
        if (bits == 2) return launch_mismatch_k<2, 3>(b, o, mp, stream);
        return launch_mismatch_k<3, 3>(b, o, mp, stream);
    }
    if (bits == 1) return launch_mismatch_k<1, 4>(b, o, mp, stream);
    if (bits == 2) return launch_mismatch_k<2, 4>(b, o, mp, stream);
    return launch_mismatch_k<3, 4>(b, o, mp, stream);
}

int kmg_mismatch_diag_launch(const uint32_t* planes, int64_t n, int L, int k, int m, double* sd, cudaStream_t stream) {
    KMG_REQUIRE(L >= 1 && L <= KMG_MAX_L, KMG_ERR_UNSUPPORTED, "sequence length %d not supported (1..%d)", L, KMG_MAX_L);
    MismatchParams mp;
    int bits = 0;
    int rc = fill_mismatch_params(k, m, L, &mp, &bits);
    if (rc) return rc;
    if (n <= 0) return KMG_OK;
    if (mm_result_words(mp) == 3) {
        if (bits == 1) return launch_mismatch_diag_k<1, 3>(planes, n, mp, sd, stream);
        if (bits == 2) return launch_mismatch_diag_k<2, 3>(planes, n, mp, sd, stream);
        return launch_mismatch_diag_k<3, 3>(planes, n, mp, sd, stream);
    }
    if (bits == 1) return launch_mismatch_diag_k<1, 4>(planes, n, mp, sd, stream);
    if (bits == 2) return launch_mismatch_diag_k<2, 4>(planes, n, mp, sd, stream);
    return launch_mismatch_diag_k<3, 4>(planes, n, mp, sd, stream);
}

int kmg_wd_launch(const PairBlock* b, int d, cudaStream_t stream) {
    int rc = check_block(b);
    if (rc) return rc;
    KMG_REQUIRE(d >= 1 && d <= 127, KMG_ERR_ARG, "weighted degree: need 1 <= d <= 127 (d=%d)", d);
    KMG_REQUIRE(b->out_dtype == KMG_OUT_F64, KMG_ERR_ARG, "weighted degree Gram is fp64 (kernels.py:74-81)");
    KMG_REQUIRE((b->sd_rows == nullptr) == (b->sd_cols == nullptr), KMG_ERR_ARG, "weighted degree: sd_rows and sd_cols go together");
    if (b->rows == 0 || b->cols == 0) return KMG_OK;
    WdParams wp;
    wp.d = d; wp.L = b->L;
    // Python: L - 1 + (1 - d) / 3  -> int + (int / int -> float)
    wp.diag = (double)(b->L - 1) + (double)(1 - d) / 3.0;
    for (int k = 1; k <= 127; ++k)
        wp.beta[k - 1] = k <= d ? (double)(2 * (d - k + 1)) / (double)d / (double)(d + 1) : 0.0;  // kernels.py:61
    wp.beta[127] = 0.0;
    const OutSpec o = make_out(b);
    for (int w = 0; w < 4; ++w) wp.posmask[w] = 0u;
    if (b->L >= 2) kmg_range_mask_128(1, b->L - 1, wp.posmask);
    dim3 grid((unsigned)((b->cols + WD_TILE_C - 1) / WD_TILE_C), (unsigned)((b->rows + TILE_R - 1) / TILE_R));
    if (o.epi || o.sd_rows != nullptr) wd_kernel<true><<<grid, TILE_R * TILE_C, 0, stream>>>(b->planes_rows, b->planes_cols, o, wp);
    else wd_kernel<false><<<grid, TILE_R * TILE_C, 0, stream>>>(b->planes_rows, b->planes_cols, o, wp);
    KMG_CUDA_CHECK(cudaGetLastError());
    return KMG_OK;
}

int kmg_wds_launch(const PairBlock* b, int d, int S, cudaStream_t stream) {
    int rc = check_block(b);
    if (rc) return rc;
    KMG_REQUIRE(d >= 1 && d <= 127, KMG_ERR_ARG, "weighted degree with shifts: need 1 <= d <= 127 (d=%d)", d);
    KMG_REQUIRE(S >= 0 && S <= 7, KMG_ERR_UNSUPPORTED, "weighted degree with shifts: 0 <= S <= 7 supported (S=%d)", S);
    KMG_REQUIRE(b->out_dtype == KMG_OUT_F64 && (b->sd_rows == nullptr) == (b->sd_cols == nullptr), KMG_ERR_ARG, "weighted degree with shifts Gram is fp64");
    if (b->rows == 0 || b->cols == 0) return KMG_OK;
    WdsParams wp;
    wp.d = d; wp.S = S; wp.L = b->L;
    for (int k = 1; k <= 128; ++k)
        wp.beta[k - 1] = k <= d ? (double)(2 * (d - k + 1)) / (double)d / (double)(d + 1) : 0.0;  // kernels.py:61
    for (int s = 0; s < 8; ++s) wp.delta[s] = 1.0 / 2.0 / (double)(s + 1);                        // kernels.py:112
    const OutSpec o = make_out(b);
    dim3 grid((unsigned)((b->cols + TILE_C - 1) / TILE_C), (unsigned)((b->rows + TILE_R - 1) / TILE_R));
    const bool epi = o.epi || o.sd_rows != nullptr;
    if (epi) {
        if (S <= 1) wds_kernel<1, true><<<grid, TILE_R * TILE_C, 0, stream>>>(b->planes_rows, b->planes_cols, o, wp);
        else if (S <= 3) wds_kernel<3, true><<<grid, TILE_R * TILE_C, 0, stream>>>(b->planes_rows, b->planes_cols, o, wp);
        else wds_kernel<7, true><<<grid, TILE_R * TILE_C, 0, stream>>>(b->planes_rows, b->planes_cols, o, wp);
    } else if (S <= 1) wds_kernel<1, false><<<grid, TILE_R * TILE_C, 0, stream>>>(b->planes_rows, b->planes_cols, o, wp);
    else if (S <= 3) wds_kernel<3, false><<<grid, TILE_R * TILE_C, 0, stream>>>(b->planes_rows, b->planes_cols, o, wp);
    else wds_kernel<7, false><<<grid, TILE_R * TILE_C, 0, stream>>>(b->planes_rows, b->planes_cols, o, wp);
    KMG_CUDA_CHECK(cudaGetLastError());
    return KMG_OK;
}
