"""
fused.py -- ALIGNF / NLCK statistics and combinations straight from SEQUENCES (csrc/fused.cu).

The reference builds every kernel as an n x n numpy array first (utils.get_all_data -> kernels.select_method) and hands the
list to ALIGNF / NLCK, which slice, centre, normalise and combine the arrays (ALIGNF.py:28-58,91-94; NLCKernels.py:33-52,
97-99).  The entry points here take the sequences and the method strings instead, so that no Gram ever crosses PCIe on
its way to a statistic or a combination:

    alignf_stats(seqs, methods, idx, y)                    -> a (p,), M (p, p)      kmg_alignf_fused_host
    combine(seqs, methods, u, degree, normalize_inputs, normalize) -> Km (n, n)    kmg_combine_fused_host
    resident_grams(seqs, methods, idx, normalize_inputs)   -> [DeviceGram]          kmg_build_grams_dev

`methods` are the reference's method strings (SP_k6, MM_k5_m1, WD_d10, WDS_d3_s2, LA_e-11_d-1_b0.5_smith0_eig0; parsed with
the same rule as kernels.select_method: the first character of every '_' field is dropped).  `last_report` holds the
PCIe byte counts of the most recent call next to what the array-based path would have moved.
"""
import ctypes as C

import numpy as np

from . import _cabi
from . import host as _host
from ._cabi import check

KIND_SP, KIND_MM, KIND_WD, KIND_WDS, KIND_LA = range(5)


class Method(C.Structure):
    """kmg_method_t (include/kmg.h)."""
    _fields_ = [("kind", C.c_int32), ("k", C.c_int32), ("m", C.c_int32), ("d", C.c_int32), ("S", C.c_int32), ("smith", C.c_int32),
                ("e", C.c_double), ("dd", C.c_double), ("beta", C.c_double)]


def parse_method(method):
    """Method string -> Method, with select_method's field rule (kernels.py:479-502)."""
    f = method.split('_')
    m = Method()
    if method.startswith('WDS'):
        m.kind, m.d, m.S = KIND_WDS, int(f[1][1:]), int(f[2][1:])
    elif method.startswith('SP'):
        m.kind, m.k = KIND_SP, int(f[1][1:])
    elif method.startswith('WD'):
        m.kind, m.d = KIND_WD, int(f[1][1:])
    elif method.startswith('MM'):
        m.kind, m.k, m.m = KIND_MM, int(f[1][1:]), int(f[2][1:])
    elif method.startswith('LA'):
        m.kind, m.e, m.dd, m.beta, m.smith = KIND_LA, float(f[1][1:]), float(f[2][1:]), float(f[3][1:]), int(f[4][len('smith'):])
    else:
        raise NotImplementedError(f"fused path: method {method!r} is outside the hot path (SURVEY.md section 2)")
    return m


def _methods(methods):
    arr = (Method * len(methods))(*[parse_method(s) if isinstance(s, str) else s for s in methods])
    return arr, len(methods)


last_report = {}


def _report(name, n, L, p, nfit, moved):
    """PCIe bytes of the fused call next to the array-based path (kmg_*_host builders + kmg_alignf_stats_host / kmg_combine_host)."""
    global last_report
    if name == "alignf_stats":
        unfused = {"h2d": p * n * L + p * nfit * nfit * 8 + nfit * 16, "d2h": p * n * n * 8 + (p + p * p) * 8}
    else:
        unfused = {"h2d": p * n * L + p * n * n * 8, "d2h": p * n * n * 8 + n * n * 8}
    last_report = {"call": name, "n": n, "p": p, "nfit": nfit, "fused": {"h2d": int(moved[0]), "d2h": int(moved[1])}, "array_based": unfused}
    return last_report


def alignf_stats(seqs, methods, idx, y):
    """ALIGNF's a_i = <Kc_i, y y'>_F and M_ij = <Kc_i, Kc_j>_F (ALIGNF.py:36-58) for the fit rows `idx` of the kernels named
    by `methods` over `seqs` (all n sequences, in kernel order)."""
    buf, fmt = _host.as_seq_buffer(seqs)
    arr, p = _methods(methods)
    idx = np.ascontiguousarray(np.atleast_1d(idx), np.int64)
    y = np.ascontiguousarray(y, np.float64)
    if y.size != idx.size:
        raise ValueError("alignf_stats: one label per fit row")
    a, M = np.zeros(p), np.zeros((p, p))
    moved = (C.c_int64 * 2)()
    n = buf.shape[0]
    check(_cabi.lib().kmg_alignf_fused_host(_host._ptr(buf), n, buf.shape[1] if n else 1, fmt, arr, p, _host._ptr(idx), idx.size,
                                            _host._ptr(y), _host._ptr(a), _host._ptr(M), moved))
    _report("alignf_stats", n, buf.shape[1] if n else 1, p, idx.size, moved)
    return a, M


def combine(seqs, methods, u, degree=1, normalize_inputs=False, normalize=False):
    """(sum_m u_m K_m) ** degree over the kernels named by `methods`, built and accumulated on the device.
    ALIGNF.get_K: defaults (ALIGNF.py:93).  NLCK.get_K: normalize_inputs=True, normalize=True (NLCKernels.py:33,97-99)."""
    buf, fmt = _host.as_seq_buffer(seqs)
    arr, p = _methods(methods)
    u = np.ascontiguousarray(u, np.float64)
    if u.size != p:
        raise ValueError("combine: one weight per method")
    n = buf.shape[0]
    Km = _host._result((n, n))
    moved = (C.c_int64 * 2)()
    check(_cabi.lib().kmg_combine_fused_host(_host._ptr(buf), n, buf.shape[1] if n else 1, fmt, arr, p, _host._ptr(u), int(degree),
                                             1 if normalize_inputs else 0, 1 if normalize else 0, _host._ptr(Km), max(n, 1), moved))
    _report("combine", n, buf.shape[1] if n else 1, p, 0, moved)
    return Km


def resident_grams(seqs, methods, idx=None, normalize_inputs=False):
    """The kernels named by `methods` over seqs[idx] (all of seqs when idx is None) as device-resident matrices
    (kmg.resident.DeviceGram), e.g. NLCK's normalised fit sub-blocks (NLCKernels.py:33,36) without any upload."""
    from . import resident as _res
    buf, fmt = _host.as_seq_buffer(seqs)
    arr, p = _methods(methods)
    n = buf.shape[0]
    if idx is not None:
        idx = np.ascontiguousarray(np.atleast_1d(idx), np.int64)
    nsel = n if idx is None else idx.size
    grams = [_res.DeviceGram(nsel) for _ in range(p)]
    ptrs = (C.c_void_p * p)(*[g.ptr.value for g in grams])
    check(_cabi.lib().kmg_build_grams_dev(_host._ptr(buf), n, buf.shape[1] if n else 1, fmt, arr, p, _host._ptr(idx), nsel,
                                          1 if normalize_inputs else 0, ptrs))
    return grams
