"""
KRR.py -- drop-in for the reference's KRR.py (kernel ridge regression on a precomputed Gram) with the K_fit algebra on the
GPU: the dual coefficients a = inv(K_fit + lambda n I) y (KRR.py:33) come from a blocked fp64 Cholesky on the device
(kmg_spd_solve_host: the fit sub-block K[idx][:, idx] is gathered on the host, nfit^2 doubles go up, nfit come back).
Same class, arguments and attributes as the reference (KRR.py:4-66); the selection of support vectors, the intercept and
the predictions are a handful of host-side dot products on the coefficients and stay in numpy.
SURVEY.md section 8(f) row 2; the reference's own KRR.py runs unchanged on the Grams of this package as well.
"""
import numpy as np

from kmg import host as _host


def _rows_of(ID, wanted):
    return np.array([np.flatnonzero(ID == w) for w in wanted]).squeeze()


class KRR():
    """Kernel ridge regression used as a classifier (labels -1 / 1)."""

    def __init__(self, K, ID, eps=1e-5, lbda=0.1, solver=None):
        """K: (n, n) Gram; ID: Ids in kernel order; eps: threshold under which a coefficient is dropped; lbda: ridge."""
        self.K, self.ID, self.eps, self.lbda, self.solver = K, ID, eps, lbda, solver

    def fit(self, X, y):
        """Dual coefficients on the rows of K named by X.Id (KRR.py:22-41)."""
        self.Id_fit = np.array(X.loc[:, 'Id'])
        self.idx_fit = _rows_of(self.ID, self.Id_fit)
        self.y_fit, self.X_fit = np.array(y.loc[:, 'Bound']), X
        self.n = self.idx_fit.size
        self.a = _host.spd_solve(self.K, np.asarray(self.y_fit, dtype=np.float64), self.lbda * self.n, idx=self.idx_fit)
        keep = np.where(np.abs(self.a) > self.eps)
        self.y_fit, self.a = self.y_fit[keep], self.a[keep]
        self.idx_sv = self.idx_fit[keep]
        self.K_fit = None  # the sub-block lives on the device only for the solve
        self.y_hat = np.array([np.dot(self.a, self.K[self.idx_sv, i]).squeeze() for i in self.idx_sv])
        self.b = np.mean(self.y_fit - self.y_hat)

    def predict(self, X):
        """Signs of a' K[sv, i] + b for the rows named by X.Id (KRR.py:43-55)."""
        self.Id_pred = np.array(X.loc[:, 'Id'])
        self.idx_pred = _rows_of(self.ID, self.Id_pred)
        return np.array([np.sign(np.dot(self.a, self.K[self.idx_sv, i].squeeze()) + self.b) for i in np.atleast_1d(self.idx_pred)])

    def score(self, pred, y):
        """Fraction of correct predictions (KRR.py:57-66)."""
        label = np.array(y.loc[:, 'Bound']) if not isinstance(y, np.ndarray) else y
        assert 0 not in np.unique(label), "Labels must be -1 or 1, not 0 or 1"
        return np.mean(pred == label)
