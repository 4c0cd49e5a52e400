/*
 * kmg_oracle.c -- plain-C CPU restatement of the reference's Gram construction
 * (afiliot/Kernel-Methods-For-Genomics, kernels.py).
 *
 * *** TEST INFRASTRUCTURE ONLY *** -- never linked into, loaded by, or called from the product
 * (kernel-methods-for-genomics_b200/). Used by tests/ (as the checker at sizes numpy is too slow
 * for), by __graft_entry__.smoke() and by bench.py's cpu_baseline / `--impl reference` legs (as the
 * timed CPU baseline, kind "port").  It is itself checked against oracle/oracle_np.py and against
 * the golden vectors produced by the unmodified reference (tests/test_oracle_golden.py,
 * tests/test_oracle_c.py).
 *
 * All functions compute a rectangular block  out[r*ldo + c] = K(rowseq[r], colseq[c])  so that the
 * tests can check sampled tiles of Grams that do not fit in host memory.
 * Sequences are uint8 codes A=0,C=1,G=2,T=3 (kernels.py:184 uses 1..4), row-major n x L.
 *
 * Build: make -C oracle   (gcc -O2 -pthread; threads via a small pthread parallel-for, no OpenMP here)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

/* ---- tiny pthread parallel-for (this image has no libgomp) --------------------------------- */
static int g_threads = 0;
int orc_num_threads(void) {
    if (g_threads > 0) return g_threads;
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
void orc_set_threads(int t) { g_threads = t > 0 ? t : 0; }

typedef void (*orc_body)(int64_t i, void *ctx);
typedef struct { atomic_llong next; int64_t n; int64_t chunk; orc_body body; void *ctx; } orc_job;
static void *orc_worker(void *p) {
    orc_job *j = (orc_job *)p;
    for (;;) {
        int64_t b = atomic_fetch_add(&j->next, j->chunk);
        if (b >= j->n) break;
        int64_t e = b + j->chunk < j->n ? b + j->chunk : j->n;
        for (int64_t i = b; i < e; ++i) j->body(i, j->ctx);
    }
    return NULL;
}
static void parallel_for(int64_t n, int64_t chunk, orc_body body, void *ctx) {
    int T = orc_num_threads();
    if (T > n) T = (int)(n > 0 ? n : 1);
    orc_job job; atomic_init(&job.next, 0); job.n = n; job.chunk = chunk > 0 ? chunk : 1; job.body = body; job.ctx = ctx;
    if (T <= 1) { orc_worker(&job); return; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)T);
    for (int t = 1; t < T; ++t) pthread_create(&th[t], NULL, orc_worker, &job);
    orc_worker(&job);
    for (int t = 1; t < T; ++t) pthread_join(th[t], NULL);
    free(th);
}

/* ------------------------------------------------------------------------------------------
 * spectrum (kernels.py:12-47): Phi[u][b] = #windows of u equal to k-mer b; K = Phi Phi^T.
 * Dense feature vectors and dot products, as the reference does; summed over the list ks
 * (BASELINE.json config 3 is the sum for k=1..7).  Counts <= 101 fit uint8.
 * ------------------------------------------------------------------------------------------ */
static void spectrum_phi_row(const uint8_t *x, int L, const int *ks, int nk, uint8_t *phi) {
    int64_t off = 0;
    for (int q = 0; q < nk; ++q) {
        int k = ks[q];
        int64_t D = 1LL << (2 * k);
        for (int p = 0; p + k <= L; ++p) {
            int64_t idx = 0;
            for (int t = 0; t < k; ++t) idx = idx * 4 + x[p + t];
            phi[off + idx]++;
        }
        off += D;
    }
}

typedef struct { const uint8_t *rows, *cols; int64_t nr, nc, D, ldo; int L, nk; const int *ks; uint8_t *PR, *PC; double *out; } sp_ctx;
static void sp_phi_r(int64_t i, void *p) { sp_ctx *c = (sp_ctx *)p; spectrum_phi_row(c->rows + i * c->L, c->L, c->ks, c->nk, c->PR + i * c->D); }
static void sp_phi_c(int64_t j, void *p) { sp_ctx *c = (sp_ctx *)p; spectrum_phi_row(c->cols + j * c->L, c->L, c->ks, c->nk, c->PC + j * c->D); }
static void sp_dot(int64_t i, void *p) {
    sp_ctx *c = (sp_ctx *)p;
    const uint8_t *a = c->PR + i * c->D;
    for (int64_t j = 0; j < c->nc; ++j) {
        const uint8_t *b = c->PC + j * c->D;
        int32_t acc = 0;
        for (int64_t t = 0; t < c->D; ++t) acc += (int32_t)a[t] * (int32_t)b[t];
        c->out[i * c->ldo + j] = (double)acc;
    }
}
int orc_spectrum_block(const uint8_t *rows, int64_t nr, const uint8_t *cols, int64_t nc, int L,
                       const int *ks, int nk, double *out, int64_t ldo) {
    int64_t D = 0;
    for (int q = 0; q < nk; ++q) D += 1LL << (2 * ks[q]);
    uint8_t *PR = (uint8_t *)calloc((size_t)nr * D, 1), *PC = (uint8_t *)calloc((size_t)nc * D, 1);
    if (!PR || !PC) { free(PR); free(PC); return -1; }
    sp_ctx c = {rows, cols, nr, nc, D, ldo, L, nk, ks, PR, PC, out};
    parallel_for(nr, 8, sp_phi_r, &c);
    parallel_for(nc, 8, sp_phi_c, &c);
    parallel_for(nr, 2, sp_dot, &c);
    free(PR); free(PC);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * mismatch (kernels.py:161-217): raw K(x,y) = sum_{p,q} T[d_H(x[p:p+k], y[q:q+k])] with
 * T the common-neighbourhood-size table (oracle_np.mismatch_table); W = L-k+1 windows
 * (the reference hard-codes 101-k+1, kernels.py:171).  Exact integers, returned as int64.
 * ------------------------------------------------------------------------------------------ */
typedef struct { const uint8_t *rows, *cols; int64_t nr, nc, ldo; int L, k; const int64_t *T; int64_t *out; } mm_ctx;
static void mm_pair(int64_t ij, void *p) {
    mm_ctx *c = (mm_ctx *)p;
    int L = c->L, k = c->k, W = L - k + 1;
    int64_t i = ij / c->nc, j = ij % c->nc;
    uint8_t ne[128 * 128];
    uint8_t *buf = (L <= 128) ? ne : (uint8_t *)malloc((size_t)L * L);
    const uint8_t *x = c->rows + i * L, *y = c->cols + j * L;
    for (int a = 0; a < L; ++a)
        for (int b = 0; b < L; ++b) buf[a * L + b] = x[a] != y[b];
    int64_t acc = 0;
    for (int a = 0; a < W; ++a)
        for (int b = 0; b < W; ++b) {
            int h = 0;
            for (int t = 0; t < k; ++t) h += buf[(a + t) * L + b + t];
            acc += c->T[h];
        }
    c->out[i * c->ldo + j] = acc;
    if (buf != ne) free(buf);
}
int orc_mismatch_raw_block(const uint8_t *rows, int64_t nr, const uint8_t *cols, int64_t nc, int L,
                           int k, const int64_t *T, int64_t *out, int64_t ldo) {
    if (L - k + 1 <= 0) return -1;
    mm_ctx c = {rows, cols, nr, nc, ldo, L, k, T, out};
    parallel_for(nr * nc, 16, mm_pair, &c);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * weighted degree (kernels.py:53-101).  Off-diagonal pairs: for k=1..d, c_k = #{l in [1,L-k]:
 * x[l:l+k]==y[l:l+k]} (position 0 skipped, kernels.py:78), acc += beta_k*c_k in fp64, one
 * multiply then one add (kernels.py:80).  `diag_closed_form` selects kernels.py:96 for pairs
 * whose global indices coincide (row_index0+r == col_index0+c).
 * Compile with -ffp-contract=off so the multiply-add is not fused.
 * ------------------------------------------------------------------------------------------ */
double orc_wd_beta(int d, int k) { return (double)(2 * (d - k + 1)) / (double)d / (double)(d + 1); }

double orc_wd_pair(const uint8_t *x, const uint8_t *y, int d, int L) {
    double c_t = 0.0;
    for (int k = 1; k <= d; ++k) {
        double beta_k = orc_wd_beta(d, k);
        int c_st = 0;
        for (int l = 1; l < L - k + 1; ++l) c_st += (memcmp(x + l, y + l, (size_t)k) == 0);
        volatile double prod = beta_k * (double)c_st;
        c_t = c_t + prod;
    }
    return c_t;
}

typedef struct { const uint8_t *rows, *cols; int64_t nr, nc, r0, c0, ldo; int L, d; double *out; } wd_ctx;
static void wd_row(int64_t i, void *p) {
    wd_ctx *c = (wd_ctx *)p;
    double diag = (double)(c->L - 1) + (double)(1 - c->d) / 3.0; /* kernels.py:96: L - 1 + (1 - d) / 3 */
    for (int64_t j = 0; j < c->nc; ++j)
        c->out[i * c->ldo + j] = (c->r0 + i == c->c0 + j) ? diag : orc_wd_pair(c->rows + i * c->L, c->cols + j * c->L, c->d, c->L);
}
int orc_wd_block(const uint8_t *rows, int64_t nr, int64_t row_index0, const uint8_t *cols, int64_t nc,
                 int64_t col_index0, int L, int d, double *out, int64_t ldo) {
    wd_ctx c = {rows, cols, nr, nc, row_index0, col_index0, ldo, L, d, out};
    parallel_for(nr, 4, wd_row, &c);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * local alignment, INTENDED semantics (the reference itself returns 0.0 for every
 * pair because its five DP matrices alias one array, kernels.py:238).  Log-space fp64.
 *   smith=0: affine_align (kernels.py:226-246);  smith=1: Smith_Waterman (kernels.py:249-270).
 * out[r][c] is computed with x = the sequence of smaller global index (kernels.py:289-291
 * fills j>=i and mirrors; the recursion is not symmetric in (x,y)).
 * ------------------------------------------------------------------------------------------ */
static const int S_LA[4][4] = {{4, 0, 0, 0}, {0, 9, -3, -1}, {0, -3, 6, 2}, {0, -1, -2, 5}}; /* kernels.py:223 */

static double lse2(double a, double b) {
    double m = a > b ? a : b;
    if (m == -INFINITY) return m;
    return m + log(exp(a - m) + exp(b - m));
}
static double lse3(double a, double b, double c) {
    double m = a > b ? a : b; m = m > c ? m : c;
    if (m == -INFINITY) return m;
    return m + log(exp(a - m) + exp(b - m) + exp(c - m));
}
static double lse4(double a, double b, double c, double d) {
    double m = a > b ? a : b; m = m > c ? m : c; m = m > d ? m : d;
    if (m == -INFINITY) return m;
    return m + log(exp(a - m) + exp(b - m) + exp(c - m) + exp(d - m));
}
static double max2(double a, double b) { return a > b ? a : b; }

double orc_la_pair(const uint8_t *x, int nx, const uint8_t *y, int ny, double e, double d, double beta, int smith) {
    int ld = ny + 1;
    size_t sz = (size_t)(nx + 1) * ld;
    double *M = (double *)malloc(5 * sz * sizeof(double));
    double *X = M + sz, *Y = X + sz, *X2 = Y + sz, *Y2 = X2 + sz;
    for (size_t t = 0; t < 5 * sz; ++t) M[t] = -INFINITY;
    double bd = beta * d, be = beta * e;
    for (int i = 1; i <= nx; ++i)
        for (int j = 1; j <= ny; ++j) {
            double s = beta * (double)S_LA[x[i - 1]][y[j - 1]];
            int c = i * ld + j, up = (i - 1) * ld + j, lf = i * ld + j - 1, dg = (i - 1) * ld + j - 1;
            if (!smith) {
                M[c] = s + lse4(0.0, X[dg], Y[dg], M[dg]);
                X[c] = lse2(bd + M[up], be + X[up]);
                Y[c] = lse3(bd + M[lf], bd + X[lf], be + Y[lf]);
                X2[c] = lse2(M[up], X2[up]);
                Y2[c] = lse3(M[lf], X2[lf], Y2[lf]);
            } else {
                M[c] = s + max2(max2(0.0, X[dg]), max2(Y[dg], M[dg]));
                X[c] = max2(bd + M[up], be + X[up]);
                Y[c] = max2(max2(bd + M[lf], bd + X[lf]), be + Y[lf]);
                X2[c] = max2(M[up], X2[up]);
                Y2[c] = max2(max2(M[lf], X2[lf]), Y2[lf]);
            }
        }
    int c = nx * ld + ny;
    double r = smith ? max2(max2(0.0, X2[c]), max2(Y2[c], M[c])) : lse4(0.0, X2[c], Y2[c], M[c]);
    free(M);
    return (1.0 / beta) * r;
}

typedef struct { const uint8_t *rows, *cols; int64_t nr, nc, r0, c0, ldo; int L, smith; double e, d, beta; double *out; } la_ctx;
static void la_one(int64_t ij, void *p) {
    la_ctx *c = (la_ctx *)p;
    int64_t i = ij / c->nc, j = ij % c->nc;
    const uint8_t *a = c->rows + i * c->L, *b = c->cols + j * c->L;
    if (c->r0 + i > c->c0 + j) { const uint8_t *t = a; a = b; b = t; }
    c->out[i * c->ldo + j] = orc_la_pair(a, c->L, b, c->L, c->e, c->d, c->beta, c->smith);
}
int orc_la_block(const uint8_t *rows, int64_t nr, int64_t row_index0, const uint8_t *cols, int64_t nc,
                 int64_t col_index0, int L, double e, double d, double beta, int smith, double *out, int64_t ldo) {
    la_ctx c = {rows, cols, nr, nc, row_index0, col_index0, ldo, L, smith, e, d, beta, out};
    parallel_for(nr * nc, 1, la_one, &c);
    return 0;
}
