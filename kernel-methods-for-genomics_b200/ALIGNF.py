"""
ALIGNF.py -- drop-in for the reference's ALIGNF.py with the Gram-side algebra on the GPU.

Same class, constructor and method names as the reference (ALIGNF.py:8-113).  What moves to libkmg.so:
  * sub-block selection K[idx][:, idx] + centring of every kernel (ALIGNF.py:28-29, 36-41),
  * a_i = <Kc_i, y y'>_F (ALIGNF.py:43-48) and M_ij = <Kc_i, Kc_j>_F (ALIGNF.py:50-58)
    -- one call, `kmg_alignf_stats_host`, the centred sub-blocks never leave the device,
  * the final combination sum_i u*_i K_i over the full kernels (ALIGNF.py:91-94), `kmg_combine_host`.
What stays exactly as in the reference: the p-dimensional (p <= ~10) L-BFGS-B problem v'Mv - 2v'a, v >= 0
(ALIGNF.py:60-89) -- it is outside the hot path and runs unchanged on our a and M.

The reference's own ALIGNF.py also works unchanged on top of this package's `kernels.center_K`; this module
additionally keeps p centred n_fit x n_fit matrices and their p^2 pairwise products off the host.
"""
import numpy as np
from scipy.optimize import fmin_l_bfgs_b

from kmg import host as _host


class ALIGNF():
    """
    Implementation of ALIGNF algorithm.
    Reference: "Algorithms for Learning Kernels Based on Centered Alignment", Cortes et al. (2009)
    """
    def __init__(self, X, y, ID, kernels):
        """
        :param X: pd.DataFrame, training features
        :param y: pd.DataFrame, training labels
        :param ID: np.array, Ids (for ordering)
        :param kernels: list of kernels
        """
        self.X = X
        self.y = y.loc[:, 'Bound']
        self.ID = ID
        self.kernels = kernels
        self.Id_X = np.array(X.loc[:, 'Id'])
        self.idx = np.array([np.where(self.ID == self.Id_X[i])[0] for i in range(len(self.Id_X))]).squeeze()  # ALIGNF.py:27
        self.p = len(self.kernels)
        self.Nfeval = 1
        print('Centering kernels...')
        print('Computing vector a...')
        print('Computing matrix M...')
        a, M = _host.alignf_stats(self.kernels, np.atleast_1d(self.idx), np.asarray(self.y, dtype=np.float64))
        self.a = a.T
        self.M = M
        self.u_star = self.get_v()

    @property
    def Y(self):
        return np.outer(self.y, self.y)  # ALIGNF.py:23 (only materialised if a caller asks for it)

    def center(self, kernels):
        """ALIGNF.py:36-41 (kept for API compatibility; the constructor uses the fused path)."""
        print('Centering kernels...')
        return [_host.center(np.ascontiguousarray(K, dtype=np.float64)) for K in kernels]

    def get_a(self):
        return self.a

    def get_M(self):
        return self.M

    def loss(self, v):
        return np.dot(v.T, np.dot(self.M, v)) - 2 * np.dot(v, self.a)  # ALIGNF.py:60-61

    def jac(self, v):
        return 2 * np.dot(self.M, v) - 2 * self.a  # ALIGNF.py:63-64

    def callbackF(self, Xi, Yi=0):
        """ALIGNF.py:66-79."""
        if self.Nfeval == 1:
            self.L = self.loss(Xi)
            print('Iteration {0:2.0f} : loss={1:8.4f}'.format(self.Nfeval, self.L))
        else:
            l_next = self.loss(Xi)
            print('Iteration {0:2.0f} : loss={1:8.4f}, tol={2:8.4f}'.format(self.Nfeval, l_next, abs(self.L - l_next)))
            self.L = l_next
        self.Nfeval += 1

    def get_v(self):
        """ALIGNF.py:81-89 -- unchanged (random init, bounds v >= 0, pgtol 1e-6)."""
        print('Gradient descent...')
        v0 = np.random.randn(self.p)
        bounds = [[0, float(np.inf)]] * self.p
        res = fmin_l_bfgs_b(self.loss, v0, fprime=self.jac, bounds=bounds, pgtol=1e-6, callback=self.callbackF)
        v_star = res[0]
        return v_star / np.linalg.norm(v_star)

    def get_K(self):
        """ALIGNF.py:91-94 -- Km = sum_i u*_i K_i over the full, uncentred kernels."""
        print('Alignment vector : ', self.u_star, '\n-------------------------------------------------------------')
        return _host.combine(self.kernels, self.u_star, degree=1)


def aligned_kernels(methods):
    """ALIGNF.py:97-113 -- needs the reference's `utils` module on sys.path (data loading is out of scope)."""
    import utils
    data, data1, data2, data3, kernels, ID = utils.get_all_data(methods)
    aligned_k = []
    for d in [data1, data2, data3]:
        X, y, _, _, _ = d
        aligned_k.append(ALIGNF(X, y, ID, kernels).get_K())
    return data, data1, data2, data3, aligned_k, ID
