"""
resident.py -- Grams kept resident in HBM between calls, without torch (ctypes + libkmg's own device buffers).

SURVEY.md section 8(f) rows 2-3: the solver-side consumers of a Gram -- NLCK's 50 iterations of
(sum_m u_m K_m)**degree and of the gradient quadratic forms on the 1501 x 1501 fit sub-blocks
(NLCKernels.py:52,62-66), the sub-block gather K[idx][:, idx] (ALIGNF.py:28, NLCKernels.py:36, SVM.py:68-69) --
re-use the same matrices many times.  The `*_host` entry points upload their inputs on every call; the classes here
upload once and drive the `*_dev` entry points on the default stream.
"""
import ctypes as C

import numpy as np

from . import _cabi
from ._cabi import check


def _lib():
    return _cabi.lib()


class DeviceGram:
    """An (rows x cols) fp64 matrix in device memory."""

    def __init__(self, rows, cols=None):
        self.rows, self.cols = int(rows), int(rows if cols is None else cols)
        p = C.c_void_p()
        check(_lib().kmg_dev_malloc(self.rows * self.cols * 8, C.byref(p)))
        self.ptr = p

    @classmethod
    def from_host(cls, K):
        K = np.ascontiguousarray(K, np.float64)
        g = cls(K.shape[0], K.shape[1])
        check(_lib().kmg_dev_upload(g.ptr, K.ctypes.data_as(C.c_void_p), K.nbytes))
        return g

    def to_host(self):
        out = np.empty((self.rows, self.cols), np.float64)
        check(_lib().kmg_dev_download(out.ctypes.data_as(C.c_void_p), self.ptr, out.nbytes))
        return out

    def free(self):
        if getattr(self, "ptr", None) is not None and self.ptr.value:
            _lib().kmg_dev_free(self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass

    # ---- operations (all on the default stream; results stay on the device unless stated)
    def gather(self, idx):
        """K[idx][:, idx] (ALIGNF.py:28, NLCKernels.py:36)."""
        idx = np.ascontiguousarray(idx, np.int64)
        d_idx = C.c_void_p()
        check(_lib().kmg_dev_malloc(idx.nbytes, C.byref(d_idx)))
        try:
            check(_lib().kmg_dev_upload(d_idx, idx.ctypes.data_as(C.c_void_p), idx.nbytes))
            out = DeviceGram(idx.size)
            check(_lib().kmg_gather_dev(self.ptr, self.cols, d_idx, idx.size, out.ptr, idx.size, None))
            _sync()
        finally:
            _lib().kmg_dev_free(d_idx)
        return out

    def matvec(self, v):
        """K v on the device (KLR.IRLS's m = K alpha, KLR.py:37); v and the result are host vectors."""
        v = np.ascontiguousarray(v, np.float64)
        assert v.size == self.cols
        dv, dout = DeviceGram(1, self.cols), DeviceGram(1, self.rows)
        check(_lib().kmg_dev_upload(dv.ptr, v.ctypes.data_as(C.c_void_p), v.nbytes))
        check(_lib().kmg_matvec_dev(self.ptr, self.rows, self.cols, self.cols, dv.ptr, dout.ptr, None))
        return dout.to_host()[0]

    def spd_solve(self, b, c, s=None):
        """x = inv(S K S + c I) b with S = diag(s) (None: identity) by a blocked Cholesky on the device.
        KRR.fit (KRR.py:33): spd_solve(y, lbda * n).  KLR.WKRR (KLR.py:41-57): sqrt(W) * spd_solve(sqrt(W) * z, n * lbda, s=sqrt(W))."""
        n = self.rows
        assert self.cols == n
        b = np.ascontiguousarray(b, np.float64)
        db, dx = DeviceGram(1, n), DeviceGram(1, n)
        check(_lib().kmg_dev_upload(db.ptr, b.ctypes.data_as(C.c_void_p), b.nbytes))
        ds = None
        if s is not None:
            s = np.ascontiguousarray(s, np.float64)
            ds = DeviceGram(1, n)
            check(_lib().kmg_dev_upload(ds.ptr, s.ctypes.data_as(C.c_void_p), s.nbytes))
        work = C.c_void_p()
        check(_lib().kmg_dev_malloc(int(_lib().kmg_spd_solve_workspace_bytes(n)), C.byref(work)))
        try:
            check(_lib().kmg_spd_solve_dev(self.ptr, n, self.cols, None if ds is None else ds.ptr, float(c), db.ptr, dx.ptr, work, None))
        finally:
            _lib().kmg_dev_free(work)
        return dx.to_host()[0]

    def normalize_(self):
        """normalize_K (kernels.py:398-415) in place; returns True on the K[0,0]==1 early-out."""
        k00 = np.empty(1)
        check(_lib().kmg_dev_download(k00.ctypes.data_as(C.c_void_p), self.ptr, 8))
        if k00[0] == 1.0:
            return True
        sd = C.c_void_p()
        check(_lib().kmg_dev_malloc(self.rows * 8, C.byref(sd)))
        try:
            check(_lib().kmg_normalize_dev(self.ptr, self.rows, self.cols, sd, None))
            _sync()
        finally:
            _lib().kmg_dev_free(sd)
        return False


def _sync():
    # a zero-byte download is a no-op; a 8-byte one synchronises the default stream
    tmp = np.empty(1)
    p = C.c_void_p()
    check(_lib().kmg_dev_malloc(8, C.byref(p)))
    try:
        check(_lib().kmg_dev_download(tmp.ctypes.data_as(C.c_void_p), p, 8))
    finally:
        _lib().kmg_dev_free(p)


def combine(grams, u, degree=1, out=None):
    """(sum_m u_m K_m) ** degree on resident Grams (NLCKernels.py:52,97; ALIGNF.py:93). Returns a DeviceGram."""
    p = len(grams)
    rows, cols = grams[0].rows, grams[0].cols
    ptrs = (C.c_void_p * p)(*[g.ptr.value for g in grams])
    lds = np.array([g.cols for g in grams], np.int64)
    u = np.ascontiguousarray(u, np.float64)
    if out is None:
        out = DeviceGram(rows, cols)
    check(_lib().kmg_combine_dev(ptrs, lds.ctypes.data_as(C.c_void_p), u.ctypes.data_as(C.c_void_p), p, int(degree),
                                 rows, cols, out.ptr, cols, None))
    return out


class QuadForms:
    """alpha' (A o B_m) alpha for a list of resident B_m -- NLCK.grad (NLCKernels.py:61-66) without re-uploading anything
    but alpha (nfit doubles) per call."""

    def __init__(self, grams):
        self.grams = list(grams)
        self.n = self.grams[0].rows
        self._alpha = DeviceGram(1, self.n)
        self._part = DeviceGram(1, self.n)
        self._res = DeviceGram(1, len(self.grams))
        self._kt = DeviceGram(self.n, self.n)

    def grad(self, u, alpha, degree):
        alpha = np.ascontiguousarray(alpha, np.float64)
        check(_lib().kmg_dev_upload(self._alpha.ptr, alpha.ctypes.data_as(C.c_void_p), alpha.nbytes))
        combine(self.grams, u, degree - 1, out=self._kt)  # K_t = (sum u K)**(degree-1)
        for m, g in enumerate(self.grams):
            res_m = C.c_void_p(self._res.ptr.value + 8 * m)
            check(_lib().kmg_weighted_dot_dev(self._kt.ptr, self.n, g.ptr, g.cols, self._alpha.ptr, self.n,
                                              self._part.ptr, res_m, None))
        return -float(degree) * self._res.to_host()[0]


def reformat_data(data, grams, ID):
    """utils.reformat_data (utils.py:280-312) for Grams that live on the device: restrict every kernel to the rows of one
    data set -- train, then validation, then test -- and renumber the Ids 0 .. n_sub-1 in that order.  `grams` are
    DeviceGram objects over all sequences (Ids `ID` in kernel order); the sub-blocks K[idx][:, idx] are gathered on the
    device (kmg_gather_dev) and stay there, so a 9 000^2 -> 3 000^2 restriction moves no matrix over PCIe.
    data = (X_train, y_train, X_val, y_val, X_test), DataFrames with an 'Id' column; their 'Id' columns are rewritten in
    place, as the reference does.  Returns (X_train, y_train, X_val, y_val, X_test, sub_grams, new_ID).
    Host arrays are not accepted: for those the reference's own utils.reformat_data applies as it is."""
    from ._dual import rows_of
    X_train, y_train, X_val, y_val, X_test = data
    for g in grams:
        if not hasattr(g, "gather"):
            raise TypeError("resident.reformat_data takes device-resident Grams (DeviceGram); use utils.reformat_data for numpy arrays")
    parts = (X_train, X_val, X_test)
    wanted = np.concatenate([p.loc[:, 'Id'].to_numpy() for p in parts])
    idx = rows_of(ID, wanted)
    sub = [g.gather(idx) for g in grams]
    new_ID = np.arange(wanted.size)
    cuts = np.cumsum([0] + [p.shape[0] for p in parts])
    for p, lo, hi in zip(parts, cuts[:-1], cuts[1:]):
        p.Id = new_ID[lo:hi]
    y_train.Id = new_ID[:y_train.shape[0]]
    y_val.Id = new_ID[cuts[1]:cuts[1] + y_val.shape[0]]
    return X_train, y_train, X_val, y_val, X_test, sub, new_ID
