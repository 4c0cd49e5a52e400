"""
NLCKernels.py -- drop-in for the reference's NLCKernels.py with the Gram-side algebra on the GPU.

Public surface as in the reference (NLCKernels.py:12-150): class `NLCK(X, y, ID, kernels, C, eps, degree)` with
`normalize_kernels`, `svm_step`, `grad`, `normalize`, `fit`, `get_K`, and the module-level `cross_validation` harness
(host glue, NLCKernels.py:103-150).

What runs in libkmg.so:
  * normalize_kernels: normalize_K of every kernel, in place (NLCKernels.py:43-48),
  * the K-line of svm_step, (sum_m u_m K_m) ** degree on the fit sub-blocks (NLCKernels.py:52),
  * grad: -degree * alpha' ((sum u K)^(degree-1) o K_m) alpha for every m (NLCKernels.py:61-66),
    -- both on fit sub-blocks uploaded once and kept resident in HBM (kmg/resident.py) for all iterations,
  * get_K: (sum_m u*_m K_m) ** degree followed by normalize_K over the full kernels (NLCKernels.py:97-99).
What stays on the host: the C-SVM dual of svm_step (a cvxopt QP, NLCKernels.py:53-59) and the projected-gradient loop
over the p weights (NLCKernels.py:68-92) -- solver code, outside the hot path, written here
independently with the reference's update rule and settings.  cvxopt is imported lazily, so everything except
svm_step / fit works without it.
"""
import numpy as np

from kernels import normalize_K
from kmg import host as _host
from kmg import resident as _res


class NLCK():
    """Non-linear (polynomial) combination of kernels (Cortes, Mohri, Rostamizadeh): projected gradient on the weights
    u of K_u = (sum_m u_m K_m)^degree, every step solving the C-SVM dual for the current K_u."""

    def __init__(self, X, y, ID, kernels, C=1e-5, eps=1e-8, degree=2):
        """X, y: training features / labels (DataFrames with 'Id', 'Bound'); ID: Ids in kernel order; kernels: list of
        (n, n) float64 Grams (normalised in place); C: SVM box constant; eps: stopping threshold on the weight update;
        degree: order of the polynomial combination."""
        self.X, self.ID = X, ID
        self.y = y.loc[:, 'Bound']
        self.n = y.shape[0]
        self.kernels = self.normalize_kernels(kernels)
        self.Id_X = X.loc[:, 'Id'].to_numpy()
        self.idx = np.array([np.flatnonzero(self.ID == w) for w in self.Id_X]).squeeze()  # NLCKernels.py:35
        self.kernels_fit = [np.ascontiguousarray(K[self.idx][:, self.idx]) for K in self.kernels]  # NLCKernels.py:36
        self.p = len(self.kernels_fit)
        self.C, self.eps, self.degree = C, eps, degree
        self.lbda = 1 / (2 * self.C * self.n)
        self._fit_dev = None  # fit sub-blocks resident in HBM for the whole projected-gradient loop

    @classmethod
    def from_sequences(cls, seqs, methods, idx, y, C=1e-5, eps=1e-8, degree=2):
        """The same object from SEQUENCES: `seqs` are all n sequences in kernel order, `methods` the reference's method
        strings, `idx` the fit rows, `y` their labels.  The normalised fit sub-blocks (NLCKernels.py:33,36) are built on the
        device and stay there for every iteration (kmg.fused.resident_grams), and get_K accumulates u_m K_m, the power and
        the final normalisation in the Gram epilogues (kmg.fused.combine): no n x n kernel is built on the host."""
        from kmg import fused as _fused
        self = cls.__new__(cls)
        self.X = self.ID = self.kernels = self.kernels_fit = None
        self._seqs, self._methods = seqs, list(methods)
        self.idx = np.atleast_1d(np.asarray(idx))
        self.y = np.asarray(y, dtype=np.float64)
        self.n = self.y.shape[0]
        self.p = len(self._methods)
        self.C, self.eps, self.degree = C, eps, degree
        self.lbda = 1 / (2 * self.C * self.n)
        grams = _fused.resident_grams(seqs, self._methods, self.idx, normalize_inputs=True)
        self._fit_dev = (grams, _res.QuadForms(grams), _res.DeviceGram(self.n))
        return self

    def _resident(self):
        """Upload the p fit sub-blocks once; svm_step and grad then re-use them every iteration."""
        if self._fit_dev is None:
            grams = [_res.DeviceGram.from_host(K) for K in self.kernels_fit]
            self._fit_dev = (grams, _res.QuadForms(grams), _res.DeviceGram(self.kernels_fit[0].shape[0]))
        return self._fit_dev

    def normalize_kernels(self, kernels):
        """normalize_K of every kernel (it mutates its argument and returns it, NLCKernels.py:43-48)."""
        out = []
        for number, K in enumerate(kernels, start=1):
            print('Normalizing kernel {}...'.format(number))
            out.append(normalize_K(K))
        return out

    def svm_step(self, u):
        """Dual variables of the C-SVM on K_u (NLCKernels.py:50-59): min 1/2 a'K_u a - y'a  s.t. 0 <= y_i a_i <= C.
        K_u comes from the device; the QP is cvxopt's, as in the reference."""
        from cvxopt import matrix, solvers, spmatrix
        solvers.options['show_progress'] = False
        grams, _, out = self._resident()
        K_u = _res.combine(grams, u, degree=self.degree, out=out).to_host()
        y = np.asarray(self.y, dtype=float)
        n, pos = self.n, np.arange(self.n)
        # the 2n box rows  y_i a_i <= C  and  -y_i a_i <= 0  as one sparse matrix
        box = spmatrix(np.concatenate((y, -y)), np.concatenate((pos, pos + n)), np.concatenate((pos, pos)), tc='d')
        rhs = matrix(np.concatenate((np.full(n, float(self.C)), np.zeros(n))), tc='d')
        sol = solvers.qp(matrix(K_u, tc='d'), matrix(-y, tc='d'), box, rhs)
        return np.asarray(sol['x']).ravel()

    def grad(self, u, alpha):
        """d/du_m of the dual objective: -degree * alpha' (K_t o K_m) alpha, K_t = (sum u K)^(degree-1) (NLCKernels.py:61-66)."""
        return self._resident()[1].grad(u, alpha, self.degree)

    def normalize(self, u, u0, fnorm):
        """Put u on the sphere of radius fnorm around u0 (NLCKernels.py:68-72)."""
        d = u - u0
        return u0 + d * (fnorm / np.sqrt((d * d).sum()))

    def fit(self, u0=0, fnorm=10, n_iter=20, eta=1):
        """Projected gradient (NLCKernels.py:74-92): step along -grad, back onto the sphere, clip at zero; the step
        shrinks by 0.8 whenever the update grows; stop once the update is below eps."""
        def project(w):
            return np.maximum(self.normalize(w, u0, fnorm), 0.0)

        u = project(np.ones(self.p))
        last_move = np.inf
        for it in range(n_iter):
            print('Iteration {}, u={}, score={:0.5f}'.format(it, u, last_move))
            alpha = self.svm_step(u)
            nxt = project(u - eta * self.grad(u, alpha))
            move = np.abs(nxt - u).max()
            if move > last_move:
                eta *= 0.8
            u, last_move = nxt, move
            if move < self.eps:
                break
        return u

    def get_K(self, u0=0, fnorm=1, n_iter=50, eta=1):
        """(sum_m u*_m K_m) ** degree over the full kernels, normalised (NLCKernels.py:94-100)."""
        u_star = self.fit(u0, fnorm, n_iter, eta)
        print('Alignment vector : ', u_star)
        print('Normalizing final kernel...')
        if self.kernels is None:  # built from sequences: accumulate, power and normalisation in the Gram epilogues
            from kmg import fused as _fused
            return _fused.combine(self._seqs, self._methods, u_star, degree=self.degree, normalize_inputs=True, normalize=True)
        # combination, power and normalize_K (including its K[0,0]==1 early-out) in one device pass
        return _host.combine(self.kernels, u_star, degree=self.degree, normalize=True)


def cross_validation(k, methods, Cs_NLK, Cs_SVM, degrees, lambdas):
    """Grid search over NLCK's (C, degree, lambda) with an inner 3-fold search of the C-SVM constant -- the harness of
    NLCKernels.py:103-150, called by main.py:72.  Host-side glue only: every Gram it touches comes from
    `utils.get_all_data` (this repo's `kernels.select_method` behind it) and every combination from `NLCK.get_K` above.
    `utils`, pandas and itertools belong to the caller's environment (the reference's utils.py needs cvxopt), so they are
    imported here, not at module load.

    k: data set 1..3; methods: kernel method strings; Cs_NLK / degrees / lambdas: the NLCK grid (`lambda` is get_K's
    `fnorm`); Cs_SVM: candidates for the C-SVM constant.  Returns a DataFrame with one row per grid point and the columns
    'methods', 'C NLCK', 'd', 'lambda', 'Best C CSVM', 'val acc'."""
    from itertools import product

    import pandas as pd
    import utils

    data, data1, data2, data3, kernels, ID = utils.get_all_data(methods)
    p = len(kernels)
    grid = list(product(Cs_NLK, degrees, lambdas))
    blank = np.zeros(len(grid))
    results = pd.DataFrame({'methods': [methods] * len(grid), 'C NLCK': blank, 'd': blank, 'lambda': blank,
                            'Best C CSVM': blank, 'val acc': blank})
    X_train, y_train, X_val, y_val, X_test, kernels, ID = utils.reformat_data((data1, data2, data3)[k - 1], kernels, ID)
    for row, (C, d, lbda) in enumerate(grid):
        print('NLCK C={}, degree={}, lambda={}'.format(C, d, lbda))
        Km = NLCK(X_train, y_train, ID, kernels, C=C, eps=1e-9, degree=d).get_K(fnorm=lbda)
        C_opt, _, _, _, mean_scores_te = utils.cross_validation(
            Ps=Cs_SVM, data=[X_train, y_train, X_val, y_val, X_test], algo='CSVM', kfolds=3, K=Km, ID=ID,
            pickleName='cv_C_SVM_NLCK_C{}_d{}_l{}_p{}_k{}.pkl'.format(C, d, lbda, p, k))
        results.iloc[row, 1:6] = C, d, lbda, C_opt, np.max(mean_scores_te)
    return results
