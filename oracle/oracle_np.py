"""
oracle_np.py -- CPU restatement (numpy) of the Gram-construction hot path of
afiliot/Kernel-Methods-For-Genomics (`kernels.py`, plus the Gram-side algebra of
`ALIGNF.py` / `NLCKernels.py`).

*** TEST INFRASTRUCTURE ONLY ***  Nothing under `oracle/` is part of the product.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
reference` legs may import it, and there only as the checker / the timed CPU baseline.
The product path (`kernel-methods-for-genomics_b200/`) never imports this module and has
no CPU fallback.

Parity status (see DESIGN.md section "Oracle"):
  * spectrum, weighted degree, mismatch (raw + normalised), normalize_K, center_K are PINNED:
    `oracle/gen_golden.py` ran the unmodified reference (`/root/reference/kernels.py`) in the
    build container and `tests/test_oracle_golden.py` checks this file against those vectors
    bit-for-bit (center_K: normwise 1e-12), plus the SHA-256 known answers of SURVEY.md App. B.
  * local alignment: the reference returns exactly 0.0 for every pair (five aliased arrays,
    `kernels.py:238`), which `la_reference_compat` reproduces and the golden vectors pin.
    The *intended* Vert-Saigo recursion (`la_affine_intended`, `la_smith_intended`) is
    PARITY UNPINNED -- the reference cannot produce a value for it; it is pinned only against
    an independent high-precision (mpmath) evaluation of the same recursion.

Every function cites the reference file:line it follows.
"""
from itertools import product as _product
from math import comb as _comb, factorial as _fact

import numpy as np

ALPHABET = "ACGT"  # kernels.py:37,184 -- lexicographic A<C<G<T, codes 0..3


# --------------------------------------------------------------------------------------
# encoding  (kernels.py:178-193 `letter_to_num`/`format`: A,C,G,T -> 1..4; we use 0..3)
# --------------------------------------------------------------------------------------
def encode(seqs):
    """list/Series of equal-length ACGT strings -> uint8 codes (n, L), A=0 C=1 G=2 T=3."""
    seqs = list(seqs)
    n = len(seqs)
    if n == 0:
        return np.zeros((0, 0), np.uint8)
    L = len(seqs[0])
    raw = np.frombuffer("".join(seqs).encode("ascii"), dtype=np.uint8)
    if raw.size != n * L:
        raise ValueError("sequences must all have the same length")
    raw = raw.reshape(n, L)
    lut = np.full(256, 255, np.uint8)
    for c, ch in enumerate(ALPHABET):
        lut[ord(ch)] = c
    codes = lut[raw]
    if (codes == 255).any():
        raise ValueError("non-ACGT character in sequence")
    return codes


def decode(codes):
    lut = np.frombuffer(ALPHABET.encode(), dtype=np.uint8)
    return ["".join(map(chr, lut[r])) for r in np.asarray(codes)]


def synthetic_codes(n, L=101, seed=0):
    """uniform iid bases, PCG64(seed)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.integers(0, 4, size=(n, L), dtype=np.uint8)


# --------------------------------------------------------------------------------------
# spectrum kernel  (kernels.py:12-47)
# --------------------------------------------------------------------------------------
def kmer_index(codes, k):
    """index of every window's k-mer in product('ACGT', repeat=k) order  -> int64 (n, L-k+1)."""
    codes = np.asarray(codes, np.int64)
    n, L = codes.shape
    W = L - k + 1
    idx = np.zeros((n, max(W, 0)), np.int64)
    for t in range(k):
        idx = idx * 4 + codes[:, t:t + W]
    return idx


def spectrum_phi(codes, k):
    """get_phi_u (kernels.py:12-25): counts of each of the 4^k k-mers over L-k+1 windows."""
    n = codes.shape[0]
    idx = kmer_index(codes, k)
    phi = np.zeros((n, 4 ** k), np.int64)
    rows = np.repeat(np.arange(n), idx.shape[1])
    np.add.at(phi, (rows, idx.ravel()), 1)
    return phi


def spectrum_gram(codes, ks):
    """get_spectrum_K (kernels.py:28-47) for one k, or the sum over several k (BASELINE config 3).
    Unnormalised; float64 holding exact integers."""
    if np.isscalar(ks):
        ks = [int(ks)]
    n = codes.shape[0]
    K = np.zeros((n, n), np.int64)
    for k in ks:
        phi = spectrum_phi(codes, k).astype(np.float64)  # exact: products <= 101^2, sums < 2^53
        K += (phi @ phi.T).astype(np.int64)
    return K.astype(np.float64)


# --------------------------------------------------------------------------------------
# normalise / centre  (kernels.py:387-415)
# --------------------------------------------------------------------------------------
def normalize_K(K):
    """normalize_K (kernels.py:398-415). In place; returns the same object.
    Early-out iff K[0,0]==1; else K_ij /= sqrt(K_ii)*sqrt(K_jj) for j>i, mirrored, diag := 1."""
    if K.shape[0] == 0:
        return K
    if K[0, 0] == 1:
        return K
    dg = np.sqrt(np.diag(K))
    den = dg[:, None] * dg[None, :]
    iu = np.triu_indices(K.shape[0], 1)
    up = K[iu] / den[iu]
    K[iu] = up
    K.T[iu] = up
    np.fill_diagonal(K, 1.0)
    return K


def center_K(K):
    """center_K (kernels.py:387-395): (I-11'/n) K (I-11'/n) through multi_dot, as the reference does."""
    n = K.shape[0]
    B = np.eye(n) - np.ones((n, n)) / n
    return np.linalg.multi_dot([B, K, B])


# --------------------------------------------------------------------------------------
# weighted degree  (kernels.py:53-101)
# --------------------------------------------------------------------------------------
def wd_beta(d, k):
    """beta (kernels.py:61), Python evaluation order."""
    return 2 * (d - k + 1) / d / (d + 1)


def wd_counts(codes, d):
    """c_k[i,j] = #{l in [1, L-k] : x_i[l:l+k]==x_j[l:l+k]}  (kernels.py:78-79; position 0 skipped).
    Returns int64 (d, n, n)."""
    codes = np.asarray(codes)
    n, L = codes.shape
    eq = (codes[:, None, 1:] == codes[None, :, 1:])  # (n,n,L-1) positions 1..L-1
    out = np.zeros((d, n, n), np.int64)
    run = np.ones((n, n, L - 1), bool)
    for k in range(1, d + 1):
        # run[l] = eq[l..l+k-1] all true, l index 0 <-> position 1
        W = L - 1 - (k - 1)
        if W <= 0:
            break
        run = run[:, :, :W] & eq[:, :, k - 1:k - 1 + W]
        out[k - 1] = run.sum(-1)
    return out


def wd_gram(codes, d):
    """get_WD_K (kernels.py:84-101): off-diagonal = sequential fp64 sum_k beta_k*c_k (kernels.py:74-81),
    diagonal = closed form L-1+(1-d)/3 (kernels.py:96)."""
    codes = np.asarray(codes)
    n, L = codes.shape
    K = np.zeros((n, n))
    B = 64
    for i0 in range(0, n, B):
        ci = codes[i0:i0 + B]
        for j0 in range(0, n, B):
            cj = codes[j0:j0 + B]
            eq = (ci[:, None, 1:] == cj[None, :, 1:])
            acc = np.zeros((ci.shape[0], cj.shape[0]))
            run = np.ones(eq.shape, bool)
            for k in range(1, d + 1):
                W = L - 1 - (k - 1)
                if W > 0:
                    run = run[:, :, :W] & eq[:, :, k - 1:k - 1 + W]
                    c = run.sum(-1).astype(np.float64)
                else:
                    c = np.zeros(acc.shape)
                acc = acc + wd_beta(d, k) * c  # c_t += beta_k * c_st (one mul, one add, fp64)
            K[i0:i0 + B, j0:j0 + B] = acc
    np.fill_diagonal(K, L - 1 + (1 - d) / 3)
    return K


def wds_pair(x, y, d, S, L):
    """get_WDShifts_d (kernels.py:115-135) on byte strings: slices that run past the end are shorter and
    therefore unequal, exactly as with the reference's Python strings; fp64 accumulation in the same order."""
    c_t = 0
    for k in range(1, d + 1):
        beta_k = wd_beta(d, k)
        c_st = 0
        for i in range(1, L - k + 1):
            for s in range(0, S + 1):
                if s + i < L:
                    c_st += (1 / 2 / (s + 1)) * ((x[i + s:i + s + k] == y[i:i + k]) + (x[i:i + k] == y[i + s:i + s + k]))
        c_t += beta_k * c_st
    return c_t


def wds_gram(codes, d, S):
    """get_WDShifts_K (kernels.py:138-155): j >= i computed (diagonal included, no closed form), mirrored."""
    codes = np.asarray(codes, np.uint8)
    n, L = codes.shape
    seqs = [bytes(r) for r in codes]
    K = np.zeros((n, n))
    for i in range(n):
        for j in range(i, n):
            K[i, j] = wds_pair(seqs[i], seqs[j], d, S, L)
            K[j, i] = K[i, j]
    return K


# --------------------------------------------------------------------------------------
# mismatch kernel  (kernels.py:161-217)
# --------------------------------------------------------------------------------------
def mismatch_table(k, m, A=4):
    """T[delta] = #{b in alphabet^k : d_H(u,b)<=m and d_H(v,b)<=m} for any u,v with d_H(u,v)=delta.
    (Exact pairwise identity for <phi_km(x), phi_km(y)> of kernels.py:161-175)
    On the delta differing positions b may equal u (a of them), equal v (bb) or neither (c);
    on the k-delta agreeing positions b differs in t places."""
    T = []
    for delta in range(k + 1):
        tot = 0
        for a in range(delta + 1):
            for bb in range(delta - a + 1):
                c = delta - a - bb
                multi = _fact(delta) // (_fact(a) * _fact(bb) * _fact(c))
                for t in range(k - delta + 1):
                    # d(u,b) = bb + c + t ; d(v,b) = a + c + t
                    if bb + c + t <= m and a + c + t <= m:
                        tot += multi * (A - 2) ** c * _comb(k - delta, t) * (A - 1) ** t
        T.append(tot)
    return T


def mismatch_phi(codes, k, m):
    """get_phi_km (kernels.py:161-175), dense over all 4^k k-mers. Only feasible for small k."""
    codes = np.asarray(codes)
    n, L = codes.shape
    W = L - k + 1
    betas = np.array(list(_product(range(4), repeat=k)), np.uint8)  # (4^k, k)
    phi = np.zeros((n, betas.shape[0]), np.int64)
    for i in range(W):
        win = codes[:, i:i + k]
        ham = (win[:, None, :] != betas[None, :, :]).sum(-1)
        phi += (ham <= m)
    return phi


def mismatch_gram_raw_dense(codes, k, m):
    phi = mismatch_phi(codes, k, m).astype(np.float64)
    return phi @ phi.T


def mismatch_gram_raw(codes, k, m):
    """K_raw(x,y) = sum_{p,q} T[d_H(x[p:p+k], y[q:q+k])] -- int64 (n,n). Any k."""
    codes = np.asarray(codes)
    n, L = codes.shape
    W = L - k + 1
    T = np.array(mismatch_table(k, m), np.int64)
    K = np.zeros((n, n), np.int64)
    B = 16
    for i0 in range(0, n, B):
        ci = codes[i0:i0 + B]
        for j0 in range(0, n, B):
            cj = codes[j0:j0 + B]
            ne = (ci[:, None, :, None] != cj[None, :, None, :]).astype(np.int8)  # (bi,bj,L,L)
            ham = np.zeros((ci.shape[0], cj.shape[0], W, W), np.int8)
            for t in range(k):
                ham += ne[:, :, t:t + W, t:t + W]
            K[i0:i0 + B, j0:j0 + B] = T[ham].sum((-1, -2))
    return K


def mismatch_gram(codes, k, m):
    """get_mismatch_K (kernels.py:196-217): raw Gram then normalize_K."""
    return normalize_K(mismatch_gram_raw(codes, k, m).astype(np.float64))


# --------------------------------------------------------------------------------------
# local alignment  (kernels.py:223-302)
# --------------------------------------------------------------------------------------
S_LA = np.array([[4, 0, 0, 0], [0, 9, -3, -1], [0, -3, 6, 2], [0, -1, -2, 5]])  # kernels.py:223 (asymmetric)


def la_reference_compat(codes, e=11, d=1, beta=0.5, smith=0):
    """What kernels.py:226-270 actually evaluates: the five DP matrices are ONE aliased array
    (kernels.py:238,262) whose last write per cell is A[i,j]=3*A[i,j-1] with A[i,0]=0, and the loops
    never touch [n_x,n_y] -- every pair gives (1/beta)*log(1+0) = 0.0."""
    n = np.asarray(codes).shape[0]
    return np.zeros((n, n))


def _logaddexp_many(*v):
    v = np.array(v, dtype=np.float64)
    mx = v.max()
    if mx == -np.inf:
        return -np.inf
    return mx + np.log(np.exp(v - mx).sum())


def la_affine_intended(x, y, e, d, beta):
    """Intended semantics of affine_align (kernels.py:226-246) with five DISTINCT matrices, loops
    i=1..n_x, j=1..n_y, s = S[x[i-1], y[j-1]].  Log-space fp64.  PARITY UNPINNED."""
    nx, ny = len(x), len(y)
    NI = -np.inf
    M = np.full((nx + 1, ny + 1), NI); X = M.copy(); Y = M.copy(); X2 = M.copy(); Y2 = M.copy()
    bd, be = beta * d, beta * e
    for i in range(1, nx + 1):
        for j in range(1, ny + 1):
            s = beta * S_LA[x[i - 1], y[j - 1]]
            M[i, j] = s + _logaddexp_many(0.0, X[i - 1, j - 1], Y[i - 1, j - 1], M[i - 1, j - 1])
            X[i, j] = _logaddexp_many(bd + M[i - 1, j], be + X[i - 1, j])
            Y[i, j] = _logaddexp_many(bd + M[i, j - 1], bd + X[i, j - 1], be + Y[i, j - 1])
            X2[i, j] = _logaddexp_many(M[i - 1, j], X2[i - 1, j])
            Y2[i, j] = _logaddexp_many(M[i, j - 1], X2[i, j - 1], Y2[i, j - 1])
    return (1 / beta) * _logaddexp_many(0.0, X2[nx, ny], Y2[nx, ny], M[nx, ny])


def la_smith_intended(x, y, e, d, beta):
    """Intended semantics of Smith_Waterman (kernels.py:249-270): sums replaced by max. Log-space."""
    nx, ny = len(x), len(y)
    NI = -np.inf
    M = np.full((nx + 1, ny + 1), NI); X = M.copy(); Y = M.copy(); X2 = M.copy(); Y2 = M.copy()
    bd, be = beta * d, beta * e
    for i in range(1, nx + 1):
        for j in range(1, ny + 1):
            s = beta * S_LA[x[i - 1], y[j - 1]]
            M[i, j] = s + max(0.0, X[i - 1, j - 1], Y[i - 1, j - 1], M[i - 1, j - 1])
            X[i, j] = max(bd + M[i - 1, j], be + X[i - 1, j])
            Y[i, j] = max(bd + M[i, j - 1], bd + X[i, j - 1], be + Y[i, j - 1])
            X2[i, j] = max(M[i - 1, j], X2[i - 1, j])
            Y2[i, j] = max(M[i, j - 1], X2[i, j - 1], Y2[i, j - 1])
    return (1 / beta) * max(0.0, X2[nx, ny], Y2[nx, ny], M[nx, ny])


def la_gram_intended(codes, e=11, d=1, beta=0.5, smith=0):
    """get_LA_K (kernels.py:273-291) with the intended pair function: K[i,j] computed with x=row i,
    y=row j for j>=i and mirrored (the recursion is not symmetric in (x,y))."""
    codes = np.asarray(codes)
    n = codes.shape[0]
    f = la_smith_intended if smith else la_affine_intended
    K = np.zeros((n, n))
    for i in range(n):
        for j in range(i, n):
            K[i, j] = f(codes[i], codes[j], e, d, beta)
            K[j, i] = K[i, j]
    return K


# --------------------------------------------------------------------------------------
# ALIGNF / NLCK Gram-side algebra  (ALIGNF.py:28-58,91-94 ; NLCKernels.py:43-66,97-99)
# --------------------------------------------------------------------------------------
def alignf_stats(kernels, idx, y):
    """ALIGNF.__init__ Gram side: sub-block K[idx][:,idx] (ALIGNF.py:28), centre each (:36-41),
    a_i = sum(Kc_i * yy') (:43-48), M_ij = sum(Kc_i * Kc_j) (:50-58)."""
    y = np.asarray(y, np.float64)
    Y = np.outer(y, y)
    Kc = [center_K(K[idx][:, idx]) for K in kernels]
    p = len(Kc)
    a = np.array([(k * Y).sum() for k in Kc])
    M = np.zeros((p, p))
    for i in range(p):
        for j in range(i, p):
            M[i, j] = (Kc[i] * Kc[j]).sum()
            M[j, i] = M[i, j]
    return a, M


def combine(kernels, u, degree=1):
    """ALIGNF.get_K (ALIGNF.py:93) for degree=1; NLCK (NLCKernels.py:52,97): (sum_m u_m K_m)**degree."""
    u = np.asarray(u, np.float64)
    Km = np.sum(np.asarray(kernels) * u[:, None, None], axis=0)
    return Km ** degree if degree != 1 else Km


def nlck_grad(kernels_fit, u, alpha, degree):
    """NLCK.grad (NLCKernels.py:61-66)."""
    u = np.asarray(u, np.float64)
    K_t = np.sum(np.asarray(kernels_fit) * u[:, None, None], axis=0) ** (degree - 1)
    g = np.array([alpha.T.dot(K_t * Km).dot(alpha) for Km in kernels_fit])
    return -degree * g
