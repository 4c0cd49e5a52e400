// la_kernel.cu -- Vert-Saigo local-alignment kernel, anti-diagonal wavefront DP in registers.
//
// Reference: affine_align / Smith_Waterman / get_LA_K (kernels.py:226-302).  AS WRITTEN the
// reference returns exactly 0.0 for every pair (its five DP matrices are one aliased array,
// kernels.py:238, and the loops never reach the cell that is read); that behaviour
// is reproduced on the host side (kernels.LA_REFERENCE_COMPAT) without a kernel.  This file
// implements the INTENDED recursion, which is the measurable work:
//
//   M [i,j] = e^{b s(x_i,y_j)} (1 + X[i-1,j-1] + Y[i-1,j-1] + M[i-1,j-1])
//   X [i,j] = e^{b d} M[i-1,j] + e^{b e} X[i-1,j]
//   Y [i,j] = e^{b d} (M[i,j-1] + X[i,j-1]) + e^{b e} Y[i,j-1]
//   X2[i,j] = M[i-1,j] + X2[i-1,j]
//   Y2[i,j] = M[i,j-1] + X2[i,j-1] + Y2[i,j-1]
//   K(x,y)  = (1/b) ln(1 + X2 + Y2 + M)[n_x, n_y]
//
// A group of LP lanes (16 or 32) owns one pair; lane l owns rows l*RPL+1 .. l*RPL+RPL and walks the
// columns skewed by one step per lane (lane l is on column t-l at step t), so the group sweeps
// anti-diagonal bands; the row above a lane's strip arrives by __shfl_up from the lane that computed
// it one step earlier.  For the challenge's L = 101 two pairs share a warp (16 lanes x 7 rows):
// 116 steps of 7 cells per lane instead of 132 steps of 4 cells with 6 lanes idle.
//
// Numerics: with the reference's default parameters (e=11, d=1, beta=0.5, taken literally as
// e^{+b e}) the linear-space values overflow fp64 within ~90 cells, so the state is kept in fp64
// scaled by a group-wide power of two 2^-E: whenever the largest live value exceeds 2^64 every live
// value is multiplied by an exact power of two and E is bumped (error-free).  The check runs every
// `check_every` steps, chosen on the host from the worst-case growth per step so that nothing can
// overflow in between.  The result is (ln(sum_scaled) + E ln 2)/b -- mathematically the log-space
// evaluation the north star asks for at ~1/50th of the flops of a literal log-sum-exp per cell.
// Smith-Waterman (max-plus) runs directly in log space (adds and maxes only).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "gram_i8.h"
#include "kmg_common.cuh"
#include "pair_kernels.h"

namespace {

constexpr int LA_THREADS = 128;
constexpr int LA_YPAD = 32;  // >= LP - 1: the column index of a lane runs from -(LP-1) to L+LP-2

struct LaParams {
    int L;
    int smith;
    int check_every;    // steps between rescale checks (affine)
    double inv_beta;
    double ed, ee;      // e^{beta d}, e^{beta e}    (affine)
    double bd, be;      // beta d, beta e            (smith)
    double sub[16];     // affine: e^{beta S[a][b]};  smith: beta S[a][b]   (S = kernels.py:223, index [x][y])
    int64_t rows, cols, row_index0, col_index0;
    int symmetric;
    double* out;
    int64_t ldo;
    double* out_t;
    int64_t ldo_t;
};

__device__ __forceinline__ int code_at(const SeqPlanes& s, int pos) {
    // pos is lane-dependent: select the word without dynamic register indexing
    const int w = pos >> 5, b = pos & 31;
    const uint32_t lo = w == 0 ? s.lo[0] : (w == 1 ? s.lo[1] : (w == 2 ? s.lo[2] : s.lo[3]));
    const uint32_t hi = w == 0 ? s.hi[0] : (w == 1 ? s.hi[1] : (w == 2 ? s.hi[2] : s.hi[3]));
    return (int)(((lo >> b) & 1u) | (((hi >> b) & 1u) << 1));
}

// The two half-warps of a warp run the same number of steps, so the shuffles use the FULL mask with a segment width of
// LP: with the half-warp mask the compiler cannot prove the mask uniform and wraps every shuffle in MATCH.ANY / REDUX /
// BRA.DIV (round 1: ~15 of the 153 instructions of a step, and the top stall of the loop).
template <int LP>
__device__ __forceinline__ double shfl_up_d(double v) { return __shfl_up_sync(0xffffffffu, v, 1, LP); }

__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

// substitution value of row q of this lane's strip against the column code the address was formed for
template <int OFF>
__device__ __forceinline__ double lds_f64_off(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(OFF));
    return v;
}

template <int LP, int RPL, int MIN_BLOCKS>
__global__ void __launch_bounds__(LA_THREADS, MIN_BLOCKS)
la_kernel(const uint32_t* __restrict__ prow, const uint32_t* __restrict__ pcol, const LaParams p) {
    constexpr int GROUPS = LA_THREADS / LP;  // pairs per CTA
    // column codes of y with LA_YPAD entries of the "no base" code 4 on either side: a lane that has not reached column 0
    // yet (or is past column L-1) looks up column 4 of its substitution table -- zeros (affine) / -inf (max-plus) -- which
    // keeps an all-zero (all -inf) state unchanged, so the step needs no "is my column valid" branch
    __shared__ uint8_t ycode_s[GROUPS][KMG_MAX_L + 2 * LA_YPAD];
    // Every thread owns a private copy of the substitution values of ITS rows: tab[q][y code] = sub[x code of row q][y].
    // The address of a step is (thread base + 8 * y code); the row is an immediate offset of the load: one address
    // computation per step instead of one per cell (FP64 instructions take two issue slots each and everything else one,
    // so every other instruction removed from the step is a slot the FP64 pipe gets: ncu, profiles/r2_la_wd_ncu.txt).
    extern __shared__ double tab_s[];  // [LA_THREADS][RPL][5]
    const int group = threadIdx.x / LP, gl = threadIdx.x % LP;  // group in CTA, lane in group
    int64_t pair = (int64_t)blockIdx.x * GROUPS + group;
    const int64_t npairs = p.rows * p.cols;
    const bool in_range = pair < npairs;
    if (!in_range) pair = npairs - 1;  // keep the group alive for the shuffles; its result is not stored
    const int64_t r = pair / p.cols, c = pair % p.cols;
    const int64_t gr = p.row_index0 + r, gc = p.col_index0 + c;
    const bool skip = p.symmetric && gc < gr;  // produced by the mirror store of (c, r)
    // kernels.py:289-291: K[i,j] is evaluated with x = row i, y = row j for j >= i, then mirrored
    const bool swap = gc < gr;
    const SeqPlanes xs = swap ? kmg_load_planes(pcol, c) : kmg_load_planes(prow, r);
    const SeqPlanes ys = swap ? kmg_load_planes(prow, r) : kmg_load_planes(pcol, c);
    const int L = p.L;
    for (int jj = gl; jj < KMG_MAX_L + 2 * LA_YPAD; jj += LP) {
        const int j = jj - LA_YPAD;
        ycode_s[group][jj] = (j >= 0 && j < L) ? (uint8_t)code_at(ys, j) : (uint8_t)4;
    }
    double* tab = tab_s + (size_t)threadIdx.x * (RPL * 5);
    const double pad = p.smith ? -INFINITY : 0.0;
#pragma unroll
    for (int q = 0; q < RPL; ++q) {
        const int xcode = code_at(xs, (gl * RPL + q) & 127);
#pragma unroll
        for (int yc = 0; yc < 4; ++yc) tab[q * 5 + yc] = p.sub[xcode * 4 + yc];
        tab[q * 5 + 4] = pad;
    }
    __syncthreads();
    // a whole warp whose groups have nothing to do can leave (warp-uniform test, after the only block-wide barrier)
    if (__all_sync(0xffffffffu, skip || !in_range)) return;
    const uint32_t tab_base = (uint32_t)__cvta_generic_to_shared(tab);
    const uint8_t* ycol = &ycode_s[group][LA_YPAD - gl];  // ycol[t] = code of this lane's column at step t

    const int last_lane = (L - 1) / RPL, last_q = (L - 1) % RPL;
    const int steps = L + LP - 1;
    double result = 0.0;

    // Per cell the state is kept as the sums the recursion actually consumes (rows of a lane's strip):
    //   T = M + X       (Y [i,j] = e^{bd} T[i,j-1] + e^{be} Y[i,j-1])
    //   S = T + Y       (M [i,j] = e^{b s} (1 + S[i-1,j-1]))
    //   U = M + X2      (X2[i,j] = U[i-1,j];  Y2[i,j] = U[i,j-1] + Y2[i,j-1];  result = ln(1 + U + Y2))
    // 10 FP64 operations per cell instead of 12 and five live arrays.  T and U reproduce the written recursion bit for
    // bit; S associates (M + X) + Y instead of (X + Y) + M (far inside the 1e-12 of the parity tests).  In the max-plus
    // (Smith-Waterman) variant every one of these regroupings is exact.
    if (!p.smith) {
        // ---------------- affine_align, scaled linear space
        double T[RPL], Y[RPL], U[RPL], Y2[RPL], S[RPL];
#pragma unroll
        for (int q = 0; q < RPL; ++q) T[q] = Y[q] = U[q] = Y2[q] = S[q] = 0.0;
        // What the lane below needs from this lane's bottom row: X of ITS first row at the same column,
        // X[i+1,j] = e^{bd} M[i,j] + e^{be} X[i,j] -- formed here, so one value travels instead of two -- plus S and U.
        double xdown = 0.0;
        double dS = 0.0;            // S of the row above the strip, previous column
        int E = 0;                  // stored = true * 2^-E
        double one_s = 1.0;
        double tot = 1.0;           // (1 + U + Y2)[row L-1 of this lane's strip, column L-1], scaled by 2^-Etot
        int Etot = 0;
        const double ed = p.ed, ee = p.ee;
        // The first lane of a group has no lane above: what its shuffles return (its own values) must count as zero.
        // Folded into the arithmetic that consumes them (a multiply-add by z0 instead of an add): no select instructions.
        const double z0 = gl == 0 ? 0.0 : 1.0;
        int until_check = p.check_every;
#pragma unroll 1
        for (int t = 0; t < steps; ++t) {
            const double uX = shfl_up_d<LP>(xdown);        // X of this strip's first row at this lane's column
            const double uS = shfl_up_d<LP>(S[RPL - 1]);
            const double uU = shfl_up_d<LP>(U[RPL - 1]);
            const uint32_t col_addr = tab_base + 8u * ycol[t];
            double aM = 0.0, aX = 0.0, aU = 0.0;           // "up"   : (i-1, j), set by the cell above
            double gS = dS;                                // "diag" : (i-1, j-1)
#pragma unroll
            for (int q = 0; q < RPL; ++q) {
                double a;
                switch (q) {  // the row is an immediate offset of the load
                    case 0: a = lds_f64_off<0>(col_addr); break;
                    case 1: a = lds_f64_off<40>(col_addr); break;
                    case 2: a = lds_f64_off<80>(col_addr); break;
                    case 3: a = lds_f64_off<120>(col_addr); break;
                    case 4: a = lds_f64_off<160>(col_addr); break;
                    case 5: a = lds_f64_off<200>(col_addr); break;
                    case 6: a = lds_f64_off<240>(col_addr); break;
                    case 7: a = lds_f64_off<280>(col_addr); break;
                    case 8: a = lds_f64_off<320>(col_addr); break;
                    case 9: a = lds_f64_off<360>(col_addr); break;
                    case 10: a = lds_f64_off<400>(col_addr); break;
                    case 11: a = lds_f64_off<440>(col_addr); break;
                    default: a = lds_f64_off<480>(col_addr); break;
                }
                const double lT = T[q], lY = Y[q], lU = U[q], lY2 = Y2[q], lS = S[q];  // "left": (i, j-1)
                const double nM = q == 0 ? a * fma(z0, gS, one_s) : a * (one_s + gS);
                const double nX = q == 0 ? z0 * uX : ed * aM + ee * aX;
                const double nY = ed * lT + ee * lY;
                const double nY2 = lU + lY2;
                const double nU = q == 0 ? fma(z0, uU, nM) : nM + aU;
                const double nT = nM + nX;
                gS = lS;
                aM = nM; aX = nX; aU = nU;
                T[q] = nT; Y[q] = nY; U[q] = nU; Y2[q] = nY2; S[q] = nT + nY;
            }
            xdown = ed * aM + ee * aX;
            dS = uS;
            const int j = t - gl;  // 0-based column of this lane
            if (j == L - 1) {
                // column n_y of this lane's rows: keep the cell of row last_q (only lane last_lane's is the result)
                double u = U[0], y2 = Y2[0];
#pragma unroll
                for (int q = 1; q < RPL; ++q)
                    if (q == last_q) { u = U[q]; y2 = Y2[q]; }
                tot = (one_s + u) + y2;
                Etot = E;
            }
            if (--until_check == 0) {
                until_check = p.check_every;
                // ---- group-wide power-of-two rescale (exact).  Lanes past their last column hold values nobody reads:
                // they stay out of the maximum (they may have overflowed) but are scaled along with the rest.
                int hi = 0;
#pragma unroll
                for (int q = 0; q < RPL; ++q) {
                    hi = max(hi, __double2hiint(S[q]));   // S >= T, Y, M, X
                    hi = max(hi, __double2hiint(U[q]));
                    hi = max(hi, __double2hiint(Y2[q]));
                }
                hi = max(hi, __double2hiint(xdown));
                if (j >= L) hi = 0;
                // maximum over the LP lanes of the group: full-mask butterfly inside segments of LP lanes
#pragma unroll
                for (int o = LP / 2; o > 0; o >>= 1) hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o, LP));
                const int ex = (hi >> 20) - 1023;  // exponent of the largest live value (all values >= 0)
                if (ex > 64) {                     // uniform within the group
                    const double sc = __hiloint2double((1023 - ex) << 20, 0);  // 2^-ex
#pragma unroll
                    for (int q = 0; q < RPL; ++q) {
                        T[q] *= sc; Y[q] *= sc; U[q] *= sc; Y2[q] *= sc; S[q] *= sc;
                    }
                    xdown *= sc; dS *= sc;
                    E += ex;
                    one_s = (E < 1000) ? __hiloint2double((1023 - E) << 20, 0) : 0.0;  // 2^-E (negligible beyond)
                }
            }
        }
        if (gl == last_lane) result = p.inv_beta * (log(tot) + (double)Etot * 0.6931471805599453094);
    } else {
        // ---------------- Smith_Waterman (max-plus), log space: T = max(M, X), S = max(T, Y), U = max(M, X2)
        const double NI = -INFINITY;
        double T[RPL], Y[RPL], U[RPL], Y2[RPL], S[RPL];
#pragma unroll
        for (int q = 0; q < RPL; ++q) T[q] = Y[q] = U[q] = Y2[q] = S[q] = NI;
        double bM = NI, bX = NI, dS = NI;
        double best = 0.0;
        const double bd = p.bd, be = p.be;
#pragma unroll 1
        for (int t = 0; t < steps; ++t) {
            double uM = shfl_up_d<LP>(bM);
            double uX = shfl_up_d<LP>(bX);
            double uS = shfl_up_d<LP>(S[RPL - 1]);
            double uU = shfl_up_d<LP>(U[RPL - 1]);
            if (gl == 0) uM = uX = uS = uU = NI;
            const double* col = tab + ycol[t];
            double aM = uM, aX = uX, aU = uU;
            double gS = dS;
#pragma unroll
            for (int q = 0; q < RPL; ++q) {
                const double s = col[q * 5];
                const double lT = T[q], lY = Y[q], lU = U[q], lY2 = Y2[q], lS = S[q];
                const double nM = s + fmax(0.0, gS);
                const double nX = fmax(bd + aM, be + aX);
                const double nY = fmax(bd + lT, be + lY);   // bd + max(M, X) == max(bd + M, bd + X): rounding is monotone
                const double nY2 = fmax(lU, lY2);
                const double nU = fmax(nM, aU);
                const double nT = fmax(nM, nX);
                gS = lS;
                aM = nM; aX = nX; aU = nU;
                T[q] = nT; Y[q] = nY; U[q] = nU; Y2[q] = nY2; S[q] = fmax(nT, nY);
            }
            bM = aM; bX = aX;
            dS = uS;
            if (t - gl == L - 1) {
                double u = U[0], y2 = Y2[0];
#pragma unroll
                for (int q = 1; q < RPL; ++q)
                    if (q == last_q) { u = U[q]; y2 = Y2[q]; }
                best = fmax(0.0, fmax(u, y2));
            }
        }
        if (gl == last_lane) result = p.inv_beta * best;
    }
    if (gl == last_lane && in_range && !skip) {
        p.out[r * p.ldo + c] = result;
        if (p.symmetric && gc != gr) p.out_t[c * p.ldo_t + r] = result;
    }
}

template <int LP, int RPL>
int launch_la(const uint32_t* prow, const uint32_t* pcol, const LaParams& p, cudaStream_t stream) {
    constexpr int GROUPS = LA_THREADS / LP;
    const int64_t pairs = p.rows * p.cols;
    const int64_t blocks = (pairs + GROUPS - 1) / GROUPS;
    KMG_REQUIRE(blocks < (1ll << 31), KMG_ERR_ARG, "local alignment: block too large for one launch");
    // Resident blocks per SM: occupancy is what this kernel needs most (FP64 chains down a lane's rows; ncu: "wait" is the
    // top stall).  8 x 13 strips: 3 blocks of 128 threads at 168 registers (44 bytes of spills) run 2048 x 4096 pairs in
    // 68.8 ms against 81.0 ms with 2 blocks at 182 registers (4 blocks: 128 registers, 944 bytes of spills, 296 ms);
    // 16 x 7 strips: 5 blocks at 96 registers (36 bytes of spills) 80.6 ms, 4 blocks at 120 registers 84.1 ms.
    constexpr int MINB_DEFAULT = RPL <= 8 ? 5 : 3;
    constexpr size_t smem = (size_t)LA_THREADS * RPL * 5 * sizeof(double);
    static const int minb_env = getenv("KMG_LA_MIN_BLOCKS") ? atoi(getenv("KMG_LA_MIN_BLOCKS")) : 0;
    static bool attr_set[64][8] = {};
    int dev = 0;
    KMG_CUDA_CHECK(cudaGetDevice(&dev));
#define KMG_LA_LAUNCH(MB)                                                                                                              \
    do {                                                                                                                               \
        if (!attr_set[dev & 63][MB]) {                                                                                                 \
            KMG_CUDA_CHECK(cudaFuncSetAttribute(la_kernel<LP, RPL, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
            attr_set[dev & 63][MB] = true;                                                                                             \
        }                                                                                                                              \
        la_kernel<LP, RPL, MB><<<(unsigned)blocks, LA_THREADS, smem, stream>>>(prow, pcol, p);                                       \
    } while (0)
    const int minb = minb_env ? minb_env : MINB_DEFAULT;
    if (minb <= 2) KMG_LA_LAUNCH(2);
    else if (minb == 3) KMG_LA_LAUNCH(3);
    else if (minb == 4) KMG_LA_LAUNCH(4);
    else KMG_LA_LAUNCH(5);
#undef KMG_LA_LAUNCH
    KMG_CUDA_CHECK(cudaGetLastError());
    return KMG_OK;
}

}  // namespace

int kmg_la_launch(const PairBlock* b, double e, double d, double beta, int smith, cudaStream_t stream) {
    KMG_REQUIRE(b->L >= 1 && b->L <= KMG_MAX_L, KMG_ERR_UNSUPPORTED, "sequence length %d not supported (1..%d)", b->L, KMG_MAX_L);
    KMG_REQUIRE(b->out_dtype == KMG_OUT_F64 && b->sd_rows == nullptr, KMG_ERR_ARG, "local alignment Gram is raw fp64");
    KMG_REQUIRE(beta > 0.0 && isfinite(beta) && isfinite(e) && isfinite(d), KMG_ERR_ARG, "local alignment: need finite e, d and beta > 0");
    if (b->symmetric)
        KMG_REQUIRE(b->rows == b->cols && b->row_index0 == b->col_index0 && b->out_t != nullptr, KMG_ERR_ARG,
                    "symmetric mode needs a square diagonal block and a mirror destination");
    const int L = b->L;
    // lanes per pair / rows per lane.  The wavefront keeps LP lanes busy for L + LP - 1 steps of RPL cells: 10 201 of the
    // 16 x 116 x 7 = 12 992 cell slots do work at L = 101 (78.5 %); four pairs per warp with 8 lanes x 13 rows fill
    // 8 x 108 x 13 = 11 232 (90.8 %).  2048 x 4096 pairs, same box: 68.8 ms (8 x 13, 3 blocks per SM) against 84.1 ms (16 x 7) and
    // 90.3 ms for the round-1 kernel.  KMG_LA_SHAPE=16 selects 16 x 7.
    static const int shape = getenv("KMG_LA_SHAPE") ? atoi(getenv("KMG_LA_SHAPE")) : 8;
    const bool wide = shape == 8 && L > 96 && L <= 104;
    const int lp = wide ? 8 : ((L > 96) ? 16 : 32);
    const int rpl = wide ? 13 : ((L > 112) ? 8 : (L > 96 ? 7 : 4));
    // worst-case growth of any state value per wavefront step, in bits: each of the RPL cells of a lane's column can
    // multiply by at most 3*max(e^{b s}) or e^{b d} + e^{b e}
    const double cell_bits = 1.4427 * beta * fmax(fmax(e, d), 9.0) + 2.0;
    const double step_bits = rpl * fmax(cell_bits, 2.0);
    KMG_REQUIRE(smith || step_bits < 450.0, KMG_ERR_UNSUPPORTED, "local alignment: beta*max(e,d,9) too large for the scaled fp64 recursion");
    if (b->rows == 0 || b->cols == 0) return KMG_OK;
    static const int S[4][4] = {{4, 0, 0, 0}, {0, 9, -3, -1}, {0, -3, 6, 2}, {0, -1, -2, 5}};  // kernels.py:223
    LaParams p;
    p.L = L; p.smith = smith; p.inv_beta = 1.0 / beta;
    // values stay below 2^64 * 2^(check_every * step_bits) < 2^1000 between checks
    int ce = (int)floor(900.0 / step_bits);
    p.check_every = ce < 1 ? 1 : (ce > 8 ? 8 : ce);
    p.ed = exp(beta * d); p.ee = exp(beta * e); p.bd = beta * d; p.be = beta * e;
    for (int a = 0; a < 4; ++a)
        for (int c = 0; c < 4; ++c) p.sub[a * 4 + c] = smith ? beta * (double)S[a][c] : exp(beta * (double)S[a][c]);
    p.rows = b->rows; p.cols = b->cols; p.row_index0 = b->row_index0; p.col_index0 = b->col_index0;
    p.symmetric = b->symmetric;
    p.out = reinterpret_cast<double*>(b->out); p.ldo = b->ldo;
    p.out_t = reinterpret_cast<double*>(b->out_t); p.ldo_t = b->ldo_t;
    if (lp == 8) return launch_la<8, 13>(b->planes_rows, b->planes_cols, p, stream);
    if (lp == 16 && rpl == 7) return launch_la<16, 7>(b->planes_rows, b->planes_cols, p, stream);
    if (lp == 16 && rpl == 8) return launch_la<16, 8>(b->planes_rows, b->planes_cols, p, stream);
    return launch_la<32, 4>(b->planes_rows, b->planes_cols, p, stream);
}
