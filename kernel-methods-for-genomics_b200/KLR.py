"""
KLR.py -- drop-in for the reference's KLR.py (kernel logistic regression by IRLS on a precomputed Gram) with the K_fit
algebra on the GPU: the fit sub-block is uploaded once and stays resident (kmg.resident.DeviceGram); every IRLS iteration
does m = K alpha on the device (KLR.py:37) and solves the weighted ridge system
alpha = W^1/2 inv(W^1/2 K W^1/2 + n lambda I) W^1/2 z (KLR.py:41-57) by a blocked fp64 Cholesky on the device.  The
n-vectors W and z are formed on the host with the reference's own formulas (KLR.py:28-39).
Same class, arguments and attributes as the reference (KLR.py:4-111).  SURVEY.md section 8(f) row 2.
"""
import numpy as np

from kmg import resident as _res


def _rows_of(ID, wanted):
    return np.array([np.flatnonzero(ID == w) for w in wanted]).squeeze()


class KLR():
    """Kernel logistic regression (labels -1 / 1)."""

    def __init__(self, K, ID, eps=1e-5, lbda=0.1, tol=1e-5, maxiter=50, solver=None):
        self.K, self.ID, self.eps, self.lbda, self.tol, self.solver, self.maxiter = K, ID, eps, lbda, tol, solver, maxiter

    def sigmoid(self, x):
        return 1 / (1 + np.exp(-x))

    def IRLS(self, K, y, alpha):
        """Weights and working response of one IRLS step (KLR.py:28-39); K is a resident Gram or an array."""
        m = K.matvec(alpha) if isinstance(K, _res.DeviceGram) else np.dot(K, alpha)
        W = self.sigmoid(m) * self.sigmoid(-m)
        z = m + y / self.sigmoid(-y * m)
        return W, z

    def WKRR(self, K, W, z):
        """New alpha of the weighted kernel ridge problem (KLR.py:41-57)."""
        ws = np.sqrt(W)
        if not isinstance(K, _res.DeviceGram):
            K = _res.DeviceGram.from_host(K)
        return ws * K.spd_solve(ws * z, self.n * self.lbda, s=ws)

    def fit(self, X, y):
        """IRLS until the update is below tol or maxiter is reached (KLR.py:59-86)."""
        self.Id_fit = np.array(X.loc[:, 'Id'])
        self.idx_fit = _rows_of(self.ID, self.Id_fit)
        self.y_fit, self.X_fit = np.array(y.loc[:, 'Bound']), X
        self.n = self.idx_fit.size
        K_fit = _res.DeviceGram.from_host(np.ascontiguousarray(self.K[self.idx_fit][:, self.idx_fit]))  # one upload for all iterations
        self.K_fit = K_fit
        alpha_prev = np.zeros(self.n)
        diff = np.inf
        for _ in range(self.maxiter):
            if diff > self.tol:
                W, z = self.IRLS(K_fit, self.y_fit, alpha_prev)
                alpha = self.WKRR(K_fit, W, z)
                diff = np.linalg.norm(alpha - alpha_prev, ord=2)
                alpha_prev = alpha.copy()
        self.a = alpha_prev
        keep = np.where(np.abs(self.a) > self.eps)
        self.y_fit, self.a = self.y_fit[keep], self.a[keep]
        self.idx_sv = self.idx_fit[keep]
        self.y_hat = np.array([np.dot(self.a, self.K[self.idx_sv, i]).squeeze() for i in self.idx_sv])
        self.b = np.mean(self.y_fit - self.y_hat)

    def predict(self, X):
        self.Id_pred = np.array(X.loc[:, 'Id'])
        self.idx_pred = _rows_of(self.ID, self.Id_pred)
        return np.array([np.sign(np.dot(self.a, self.K[self.idx_sv, i].squeeze()) + self.b) for i in np.atleast_1d(self.idx_pred)])

    def score(self, pred, y):
        label = np.array(y.loc[:, 'Bound']) if not isinstance(y, np.ndarray) else y
        assert 0 not in np.unique(label), "Labels must be -1 or 1, not 0 or 1"
        return np.mean(pred == label)
