"""
kernels.py -- drop-in replacement for the reference's `kernels.py`
(afiliot/Kernel-Methods-For-Genomics), backed by hand-written sm_100a CUDA through libkmg.so.

Same function names, arguments and numpy float64 Gram-matrix return values as the reference, so
`utils.py` (`km.select_method(X, method)`, utils.py:153), `ALIGNF.py` and `NLCKernels.py`
(`center_K`, `normalize_K`) and everything above them (SVM/KRR/KLR, run.py, main.py) run unchanged.
Put this directory on `sys.path` ahead of the reference's and `import kernels`.

Every function names the reference lines it replaces.  There is NO CPU fallback: without the built
extension `import kernels` still works but the first call raises ImportError; without a GPU it
raises kmg.KmgError.

Deliberate differences from the reference (all documented in DESIGN.md):
  * `select_method` raises NotImplementedError for an unknown method string (the reference
    constructs the exception without raising it and dies with UnboundLocalError, kernels.py:504).
  * local alignment: the reference's affine_align / Smith_Waterman return exactly 0.0 for every
    pair (aliased DP matrices, kernels.py:238) and get_LA_K(eig=1) then dies in ARPACK.  By default
    this module evaluates the INTENDED Vert-Saigo recursion; set LA_REFERENCE_COMPAT = True (or
    env KMG_LA_REFERENCE_COMPAT=1) to reproduce the reference's output bit for bit.
  * sequences must all have the same length L <= 128 and contain only A, C, G, T (ValueError
    otherwise; the challenge data is 101 bp ACGT).
  * SS_ (substring) and GP_ (gappy) kernels are outside the hot path and
    raise NotImplementedError.
"""
import os

import numpy as np

from kmg import host as _host

LA_REFERENCE_COMPAT = os.environ.get("KMG_LA_REFERENCE_COMPAT", "0") == "1"

# substitution matrix extracted from BLOSUM62 (kernels.py:223) -- asymmetric, indexed [x, y]
S = np.array([[4, 0, 0, 0], [0, 9, -3, -1], [0, -3, 6, 2], [0, -1, -2, 5]])


def _seqs(X):
    """The reference reads `X.loc[:, 'seq']` positionally (kernels.py:39,94,208,287)."""
    if hasattr(X, "loc"):
        return X.loc[:, "seq"]
    return X


# ------------------------------------------------------------------------------------------------
# Spectrum kernel (kernels.py:12-47)
# ------------------------------------------------------------------------------------------------
def _beta_index(b):
    """column of k-mer `b` (string over ACGT or `format`-ted ints 1..4) in product('ACGT', repeat=k) order"""
    idx = 0
    for ch in b:
        idx = idx * 4 + ("ACGT".index(ch) if isinstance(ch, str) else int(ch) - 1)
    return idx


def get_phi_u(x, k, betas):
    """kernels.py:12-25 -- k-mer count vector of one sequence, entries ordered like `betas`."""
    phi = _host.spectrum_phi([x], int(k))[0].astype(np.float64)
    return phi[[_beta_index(b) for b in betas]] if len(betas) else np.zeros(0)


def get_spectrum_K(X, k):
    """kernels.py:28-47 -- K = Phi Phi^T, unnormalised, float64 holding exact integers."""
    return _host.spectrum_gram(_seqs(X), int(k))


def get_spectrum_sum_K(X, ks):
    """Sum of spectrum kernels over several k in one concatenated-feature GEMM (BASELINE config 3)."""
    return _host.spectrum_gram(_seqs(X), [int(k) for k in ks])


# ------------------------------------------------------------------------------------------------
# Weighted degree kernel (kernels.py:53-101)
# ------------------------------------------------------------------------------------------------
def beta(d, k):
    """kernels.py:53-61."""
    return 2 * (d - k + 1) / d / (d + 1)


def get_WD_d(x, y, d, L):
    """kernels.py:64-81 -- one pair, loop value (no closed-form diagonal)."""
    if len(x) != L or len(y) != L:
        raise ValueError("get_WD_d: sequences must have length L")
    return float(_host.wd_gram([x], int(d), cols=[y])[0, 0])


def get_WD_K(X, d):
    """kernels.py:84-101."""
    return _host.wd_gram(_seqs(X), int(d))


# ------------------------------------------------------------------------------------------------
# Mismatch kernel (kernels.py:161-217)
# ------------------------------------------------------------------------------------------------
_DIGITS = str.maketrans('ACGT', '1234')


def letter_to_num(x):
    """'ACGT' -> '1234', every other character left as it is (kernels.py:178-184)."""
    return x.translate(_DIGITS)


def format(x):
    """Sequence -> int array over 1..4; a non-ACGT character fails in int(), as in the reference (kernels.py:187-193)."""
    return np.array([int(c) for c in letter_to_num(x)], dtype=int)


def get_phi_km(x, k, m, betas):
    """kernels.py:161-175 -- x and betas are `format`-ted integer arrays (A,C,G,T -> 1..4)."""
    xs = "".join("ACGT"[int(c) - 1] for c in x)
    phi = _host.mismatch_phi([xs], int(k), int(m))[0].astype(np.float64)
    return phi[[_beta_index(b) for b in betas]] if len(betas) else np.zeros(0)


def get_mismatch_K(X, k, m):
    """kernels.py:196-217 -- raw Gram then normalize_K."""
    return _host.mismatch_gram(_seqs(X), int(k), int(m), normalize=True)


# ------------------------------------------------------------------------------------------------
# Local alignment kernel (kernels.py:220-302)
# ------------------------------------------------------------------------------------------------
def affine_align(x, y, e, d, beta):
    """kernels.py:226-246."""
    if LA_REFERENCE_COMPAT:
        return 0.0
    return float(_host.la_gram([x], e, d, beta, 0, cols=[y])[0, 0])


def Smith_Waterman(x, y, e=11, d=1, beta=0.5):
    """kernels.py:249-270."""
    if LA_REFERENCE_COMPAT:
        return 0.0
    return float(_host.la_gram([x], e, d, beta, 1, cols=[y])[0, 0])


def get_LA_K(X, e=11, d=1, beta=0.5, smith=0, eig=1):
    """kernels.py:273-302.  The reference computes a post-processed copy K1 (eigenvalue shift or
    empirical kernel map) and then returns the unprocessed K (kernels.py:302); so do we."""
    seqs = _seqs(X)
    if LA_REFERENCE_COMPAT:
        n = len(seqs)
        if eig == 1 and n > 0:
            from scipy.sparse.linalg import eigs  # reference behaviour: ARPACK fails on the zero matrix
            eigs(np.zeros((n, n)))
        return np.zeros((n, n))
    return _host.la_gram(seqs, e, d, beta, int(smith))


# ------------------------------------------------------------------------------------------------
# Normalize / centre (kernels.py:385-415)
# ------------------------------------------------------------------------------------------------
def center_K(K):
    """kernels.py:387-395 -- returns a new array."""
    return _host.center(np.asarray(K, dtype=np.float64))


def normalize_K(K):
    """kernels.py:398-415 -- in place, returns the same object; prints on the early-out."""
    if _host.normalize_inplace(K):
        print('Kernel already normalized')
    return K


# ------------------------------------------------------------------------------------------------
# Select method (kernels.py:461-505)
# ------------------------------------------------------------------------------------------------
def delta(s):
    """kernels.py:106-112."""
    return 1/2/(s+1)


def get_WDShifts_d(x, y, d, S, L):
    """kernels.py:115-135 -- one pair."""
    if len(x) != L or len(y) != L:
        raise ValueError("get_WDShifts_d: sequences must have length L")
    return float(_host.wds_gram([x], int(d), int(S), cols=[y])[0, 0])


def get_WDShifts_K(X, d, S):
    """kernels.py:138-155 -- weighted degree with shifts (first 'next' row of SURVEY.md section 8f)."""
    return _host.wds_gram(_seqs(X), int(d), int(S))


def get_string_K(X, lbda, k):
    raise NotImplementedError("SS (substring) kernel is outside the hot path (SURVEY.md section 2)")


def get_gappy_K(X, k, g):
    raise NotImplementedError("GP (gappy) kernel is outside the hot path (SURVEY.md section 2)")


# method prefix -> (builder, how to read the '_'-separated fields: (field number, characters to drop, type)).
# 'WDS' is listed before 'WD' because the latter is a prefix of the former.
def _dsl():
    one_int = ((1, 1, int),)
    two_ints = ((1, 1, int), (2, 1, int))
    return (
        ('WDS', get_WDShifts_K, two_ints),
        ('SP', get_spectrum_K, one_int),
        ('WD', get_WD_K, one_int),
        ('MM', get_mismatch_K, two_ints),
        ('LA', get_LA_K, ((1, 1, float), (2, 1, float), (3, 1, float), (4, len('smith'), int), (5, len('eig'), int))),
        ('SS', get_string_K, ((1, 1, float), (2, 1, int))),
        ('GP', get_gappy_K, two_ints),
    )


def select_method(X, method):
    """The reference's method mini-language (kernels.py:461-505): SP_k{x}, WD_d{x}, WDS_d{x}_s{y}, MM_k{x}_m{y},
    LA_e{x}_d{y}_b{z}_smith{X}_eig{Y}, SS_l{x}_k{y}, GP_k{x}_g{y}.  Every field carries a one-letter tag (or the words
    'smith' / 'eig') in front of its value; the tag is dropped, not checked, exactly as the reference does."""
    fields = method.split('_')
    for prefix, build, spec in _dsl():
        if method.startswith(prefix):
            if prefix == 'WD':
                print(fields)  # the reference echoes the split method string for this kernel (kernels.py:487)
            return build(X, *[kind(fields[pos][drop:]) for pos, drop, kind in spec])
    raise NotImplementedError('Method not implemented. Please refer to the documentation for choosing among available methods')
