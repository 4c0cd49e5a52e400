"""
ALIGNF.py -- drop-in for the reference's ALIGNF.py with the Gram-side algebra on the GPU.

Public surface as in the reference (ALIGNF.py:8-113): class `ALIGNF(X, y, ID, kernels)` with `center`, `get_a`,
`get_M`, `loss`, `jac`, `callbackF`, `get_v`, `get_K`, attributes `a`, `M`, `u_star`, and `aligned_kernels(methods)`.

What runs in libkmg.so:
  * the sub-block selection K[idx][:, idx] and the centring of every kernel (ALIGNF.py:28-29, 36-41),
  * a_i = <Kc_i, y y'>_F (ALIGNF.py:43-48) and M_ij = <Kc_i, Kc_j>_F (ALIGNF.py:50-58): one call,
    `kmg_alignf_stats_host`; the centred sub-blocks and their p^2 pairwise products never leave the device,
  * the final combination sum_i u*_i K_i over the full kernels (ALIGNF.py:91-94): `kmg_combine_host`.
What stays on the host: the p-dimensional (p <= ~10) non-negative quadratic programme  min_v v'Mv - 2 v'a, v >= 0
(ALIGNF.py:60-89) -- solver code, outside the hot path.  It is solved here with the same
method and settings the reference uses (SciPy's L-BFGS-B from a standard-normal start, pgtol 1e-6, one progress
line per iteration), written independently.
"""
import numpy as np
from scipy.optimize import fmin_l_bfgs_b

from kmg import host as _host


def _rows_of(ID, wanted):
    """Position in `ID` of every Id in `wanted` (ALIGNF.py:27); duplicates in ID resolve as np.where does."""
    pos = [np.flatnonzero(ID == w) for w in wanted]
    return np.array(pos).squeeze()


class ALIGNF():
    """Centred-alignment kernel learning (Cortes, Mohri, Rostamizadeh): weights u* maximising the alignment of
    sum_i u_i K_i with the label kernel y y', from the statistics a and M of the centred training sub-blocks."""

    def __init__(self, X, y, ID, kernels):
        """X, y: training features / labels (DataFrames with 'Id', 'Bound'); ID: Ids in kernel order; kernels: list of
        (n, n) float64 Grams."""
        self.X, self.ID, self.kernels = X, ID, kernels
        self.y = y.loc[:, 'Bound']
        self.Id_X = X.loc[:, 'Id'].to_numpy()
        self.idx = _rows_of(self.ID, self.Id_X)
        self.p = len(kernels)
        self.Nfeval = 1
        for stage in ('Centering kernels...', 'Computing vector a...', 'Computing matrix M...'):
            print(stage)
        a, M = _host.alignf_stats(kernels, np.atleast_1d(self.idx), np.asarray(self.y, dtype=np.float64))
        self.a, self.M = a.T, M
        self.u_star = self.get_v()

    @classmethod
    def from_sequences(cls, seqs, methods, idx, y):
        """The same object from SEQUENCES: `seqs` are all n sequences in kernel order, `methods` the reference's method
        strings, `idx` the fit rows, `y` their labels.  a and M come from one fused call (kmg.fused.alignf_stats: the fit
        sub-blocks are built, centred and reduced on the device; p + p^2 doubles come back), get_K from another
        (kmg.fused.combine: every Gram launch adds u_i K_i to the result in its epilogue).  No n x n kernel is built on
        the host at any point."""
        from kmg import fused as _fused
        self = cls.__new__(cls)
        self.X = self.ID = self.kernels = None
        self._seqs, self._methods = seqs, list(methods)
        self.idx = np.atleast_1d(np.asarray(idx))
        self.y = np.asarray(y, dtype=np.float64)
        self.p = len(self._methods)
        self.Nfeval = 1
        for stage in ('Centering kernels...', 'Computing vector a...', 'Computing matrix M...'):
            print(stage)
        a, M = _fused.alignf_stats(seqs, self._methods, self.idx, self.y)
        self.a, self.M = a.T, M
        self.u_star = self.get_v()
        return self

    # ---- pieces of the reference's interface that callers may still reach for
    @property
    def Y(self):
        """y y' (ALIGNF.py:23); only materialised on request -- a uses it in fused form on the device."""
        y = np.asarray(self.y, dtype=np.float64)
        return y[:, None] * y[None, :]

    def center(self, kernels):
        """Centred copies of `kernels` (ALIGNF.py:36-41); the constructor itself uses the fused statistics call."""
        print('Centering kernels...')
        return [_host.center(np.ascontiguousarray(K, dtype=np.float64)) for K in kernels]

    def get_a(self):
        return self.a

    def get_M(self):
        return self.M

    # ---- the small QP over the weights
    def loss(self, v):
        """v'Mv - 2 v'a (ALIGNF.py:60-61)."""
        return float(v @ self.M @ v - 2.0 * (v @ self.a))

    def jac(self, v):
        """Gradient of `loss` (ALIGNF.py:63-64)."""
        return 2.0 * (self.M @ v - self.a)

    def callbackF(self, Xi, Yi=0):
        """One progress line per L-BFGS-B iteration: the loss and, from the second one on, its change."""
        now = self.loss(Xi)
        line = 'Iteration {0:2.0f} : loss={1:8.4f}'.format(self.Nfeval, now)
        if self.Nfeval > 1:
            line += ', tol={0:8.4f}'.format(abs(self.L - now))
        print(line)
        self.L = now
        self.Nfeval += 1

    def get_v(self):
        """argmin of `loss` over v >= 0, scaled to unit Euclidean norm (ALIGNF.py:81-89)."""
        print('Gradient descent...')
        start = np.random.randn(self.p)
        v, _, _ = fmin_l_bfgs_b(self.loss, start, fprime=self.jac, bounds=[(0.0, None)] * self.p, pgtol=1e-6,
                                callback=self.callbackF)
        return v / np.sqrt(v @ v)

    def get_K(self):
        """Km = sum_i u*_i K_i over the full, uncentred kernels (ALIGNF.py:91-94)."""
        print('Alignment vector : ', self.u_star, '\n' + '-' * 61)
        if self.kernels is None:  # built from sequences: accumulate u_i K_i in the Gram epilogues
            from kmg import fused as _fused
            return _fused.combine(self._seqs, self._methods, self.u_star, degree=1)
        return _host.combine(self.kernels, self.u_star, degree=1)


def aligned_kernels(methods):
    """One ALIGNF combination per data set (ALIGNF.py:97-113).  Loading the data is the reference's `utils` module's
    business (out of scope here): it has to be importable."""
    import utils
    data, data1, data2, data3, kernels, ID = utils.get_all_data(methods)
    combined = [ALIGNF(d[0], d[1], ID, kernels).get_K() for d in (data1, data2, data3)]
    return data, data1, data2, data3, combined, ID
