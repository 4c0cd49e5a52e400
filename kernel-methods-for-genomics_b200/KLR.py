"""
KLR.py -- drop-in for the reference's KLR.py (kernel logistic regression by IRLS on a precomputed Gram) with the K_fit
algebra on the GPU: the fit sub-block is uploaded once and stays resident (kmg.resident.DeviceGram); every IRLS iteration
does m = K alpha on the device (KLR.py:37) and solves the weighted ridge system
alpha = W^1/2 inv(W^1/2 K W^1/2 + n lambda I) W^1/2 z (KLR.py:41-57) by a blocked fp64 Cholesky on the device.  The
n-vectors W and z are formed on the host with the reference's own formulas (KLR.py:28-39).
Same class, arguments and attributes as the reference (KLR.py:4-111); support selection, intercept, predictions and
score are kmg/_dual.py.  SURVEY.md section 8(f) row 2.
"""
import numpy as np

from kmg import resident as _res
from kmg._dual import DualClassifier


class KLR(DualClassifier):
    """Kernel logistic regression (labels -1 / 1)."""

    def __init__(self, K, ID, eps=1e-5, lbda=0.1, tol=1e-5, maxiter=50, solver=None):
        self.K, self.ID = K, ID
        self.eps, self.lbda, self.tol, self.maxiter, self.solver = eps, lbda, tol, maxiter, solver

    def sigmoid(self, x):
        return 1.0 / (1.0 + np.exp(-x))

    def IRLS(self, K, y, alpha):
        """Weights and working response of one IRLS step (KLR.py:28-39); K is a resident Gram or an array."""
        m = K.matvec(alpha) if isinstance(K, _res.DeviceGram) else np.dot(K, alpha)
        return self.sigmoid(m) * self.sigmoid(-m), m + y / self.sigmoid(-y * m)

    def WKRR(self, K, W, z):
        """New alpha of the weighted kernel ridge problem (KLR.py:41-57)."""
        ws = np.sqrt(W)
        if not isinstance(K, _res.DeviceGram):
            K = _res.DeviceGram.from_host(K)
        return ws * K.spd_solve(ws * z, self.n * self.lbda, s=ws)

    def fit(self, X, y):
        """IRLS until the update is below tol or maxiter is reached (KLR.py:59-86)."""
        self._start_fit(X, y)
        # one upload for all iterations
        self.K_fit = _res.DeviceGram.from_host(np.ascontiguousarray(self.K[self.idx_fit][:, self.idx_fit]))
        alpha = np.zeros(self.n)
        for _ in range(self.maxiter):
            W, z = self.IRLS(self.K_fit, self.y_fit, alpha)
            nxt = self.WKRR(self.K_fit, W, z)
            step = np.sqrt(((nxt - alpha) ** 2).sum())
            alpha = nxt
            if not step > self.tol:  # the reference keeps looping without updating once the step is <= tol (KLR.py:69-77)
                break
        self._finish_fit(alpha)
