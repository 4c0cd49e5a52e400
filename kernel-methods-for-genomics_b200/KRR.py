"""
KRR.py -- drop-in for the reference's KRR.py (kernel ridge regression on a precomputed Gram) with the K_fit algebra on the
GPU: the dual coefficients a = inv(K_fit + lambda n I) y (KRR.py:33) come from a blocked fp64 Cholesky on the device
(kmg_spd_solve_host: the fit sub-block K[idx][:, idx] is gathered on the host, nfit^2 doubles go up, nfit come back).
Same class, arguments and attributes as the reference (KRR.py:4-66); the selection of support vectors, the intercept and
the predictions are a handful of host-side products on the coefficients (kmg/_dual.py) and stay in numpy.
SURVEY.md section 8(f) row 2; the reference's own KRR.py runs unchanged on the Grams of this package as well.
"""
import numpy as np

from kmg import host as _host
from kmg._dual import DualClassifier


class KRR(DualClassifier):
    """Kernel ridge regression used as a classifier (labels -1 / 1)."""

    def __init__(self, K, ID, eps=1e-5, lbda=0.1, solver=None):
        """K: (n, n) Gram; ID: Ids in kernel order; eps: threshold under which a coefficient is dropped; lbda: ridge."""
        self.K, self.ID = K, ID
        self.eps, self.lbda, self.solver = eps, lbda, solver

    def fit(self, X, y):
        """Dual coefficients on the rows of K named by X.Id (KRR.py:22-41)."""
        self._start_fit(X, y)
        rhs = np.asarray(self.y_fit, dtype=np.float64)
        self.K_fit = None  # the sub-block lives on the device only for the solve
        self._finish_fit(_host.spd_solve(self.K, rhs, self.lbda * self.n, idx=self.idx_fit))
