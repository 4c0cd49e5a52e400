"""
NLCKernels.py -- drop-in for the reference's NLCKernels.py with the Gram-side algebra on the GPU.

Same class, constructor and method names as the reference (NLCKernels.py:12-100).  What moves to libkmg.so:
  * normalize_kernels: normalize_K of every kernel, in place (NLCKernels.py:43-48),
  * the K-line of svm_step, (sum_m u_m K_m) ** degree on the fit sub-blocks (NLCKernels.py:52),
  * grad: -degree * alpha' ((sum u K)^(degree-1) o K_m) alpha for every m (NLCKernels.py:61-66),
    -- both on fit sub-blocks uploaded once and kept resident in HBM (kmg/resident.py) for all iterations,
  * get_K: (sum_m u*_m K_m) ** degree followed by normalize_K over the full kernels (NLCKernels.py:97-99).
What stays as in the reference: the cvxopt QP of svm_step (NLCKernels.py:53-59) and the projected-gradient loop
(NLCKernels.py:68-92) -- solver code, outside the hot path.  cvxopt is imported lazily, so
everything except svm_step/fit works without it.
"""
import numpy as np

from kernels import normalize_K
from kmg import host as _host
from kmg import resident as _res


class NLCK():
    """
    Implementation of NLCK algorithm.
    Reference : "Learning Non-Linear Combinations of Kernels", Cortes et al. (2009)
    """
    def __init__(self, X, y, ID, kernels, C=1e-5, eps=1e-8, degree=2):
        self.X = X
        self.y = y.loc[:, 'Bound']
        self.n = y.shape[0]
        self.ID = ID
        self.kernels = self.normalize_kernels(kernels)
        self.Id_X = np.array(X.loc[:, 'Id'])
        self.idx = np.array([np.where(self.ID == self.Id_X[i])[0] for i in range(len(self.Id_X))]).squeeze()
        self.kernels_fit = [np.ascontiguousarray(K[self.idx][:, self.idx]) for K in self.kernels]  # NLCKernels.py:36
        self.p = len(self.kernels_fit)
        self.C = C
        self.lbda = 1 / (2 * self.C * self.n)
        self.eps = eps
        self.degree = degree
        self._fit_dev = None  # fit sub-blocks resident in HBM for the whole projected-gradient loop

    def _resident(self):
        """Upload the p fit sub-blocks once; svm_step and grad then re-use them every iteration."""
        if self._fit_dev is None:
            grams = [_res.DeviceGram.from_host(K) for K in self.kernels_fit]
            self._fit_dev = (grams, _res.QuadForms(grams), _res.DeviceGram(self.kernels_fit[0].shape[0]))
        return self._fit_dev

    def normalize_kernels(self, kernels):
        """NLCKernels.py:43-48 (normalize_K mutates its argument and returns it)."""
        new_kernels = []
        for k, K in enumerate(kernels):
            print('Normalizing kernel {}...'.format(k + 1))
            new_kernels.append(normalize_K(K))
        return new_kernels

    def svm_step(self, u):
        """NLCKernels.py:50-59 -- the Gram line on the GPU, the QP in cvxopt as in the reference."""
        from cvxopt import matrix, spmatrix, solvers
        solvers.options['show_progress'] = False
        r, o, z = np.arange(self.n), np.ones(self.n), np.zeros(self.n)
        grams, _, out = self._resident()
        K = _res.combine(grams, u, degree=self.degree, out=out).to_host()
        P = matrix(K.astype(float), tc='d')
        q = matrix(-self.y, tc='d')
        G = spmatrix(np.r_[self.y, -self.y], np.r_[r, r + self.n], np.r_[r, r], tc='d')
        h = matrix(np.r_[o * self.C, z], tc='d')
        sol = solvers.qp(P, q, G, h)
        return np.ravel(sol['x'])

    def grad(self, u, alpha):
        """NLCKernels.py:61-66."""
        return self._resident()[1].grad(u, alpha, self.degree)

    def normalize(self, u, u0, fnorm):
        """NLCKernels.py:68-72."""
        u_s = (u - u0)
        u_s_norm = u_s / np.sqrt(np.sum(u_s**2))
        u_s = u_s_norm * fnorm
        return u_s + u0

    def fit(self, u0=0, fnorm=10, n_iter=20, eta=1):
        """NLCKernels.py:74-92 -- unchanged."""
        u = np.ones(self.p)
        u = self.normalize(u, u0, fnorm)
        u = np.array([0 if u[i] < 0 else u[i] for i in range(self.p)])
        score_prev = np.inf
        for k in range(n_iter):
            print('Iteration {}, u={}, score={:0.5f}'.format(k, u, score_prev))
            alpha = self.svm_step(u)
            g = self.grad(u, alpha)
            u_next = self.normalize(u - eta * g, u0, fnorm)
            u_next = np.array([0 if u_next[i] < 0 else u_next[i] for i in range(self.p)])
            score = np.linalg.norm(u_next - u, np.inf)
            if score > score_prev:
                eta *= 0.8
            if score < self.eps:
                return u_next
            u = u_next
            score_prev = score.copy()
        return u_next

    def get_K(self, u0=0, fnorm=1, n_iter=50, eta=1):
        """NLCKernels.py:94-100."""
        u_star = self.fit(u0, fnorm, n_iter, eta)
        print('Alignment vector : ', u_star)
        print('Normalizing final kernel...')
        # combination, power and normalize_K (including its K[0,0]==1 early-out) in one device pass
        return _host.combine(self.kernels, u_star, degree=self.degree, normalize=True)
