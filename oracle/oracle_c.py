"""
oracle_c.py -- ctypes loader for the plain-C oracle (oracle/kmg_oracle.c).
*** TEST INFRASTRUCTURE ONLY *** (see oracle_np.py / kmg_oracle.c headers).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libkmg_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(HERE, "kmg_oracle.c")
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE])
    return SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(SO)
        _lib.orc_wd_pair.restype = C.c_double
        _lib.orc_la_pair.restype = C.c_double
        _lib.orc_wd_beta.restype = C.c_double
    return _lib


def _u8(a):
    a = np.ascontiguousarray(a, np.uint8)
    return a, a.ctypes.data_as(C.c_void_p)


def num_threads():
    return lib().orc_num_threads()


def set_threads(t):
    lib().orc_set_threads(int(t))


def spectrum_block(rows, cols, ks):
    rows, pr = _u8(rows); cols, pc = _u8(cols)
    ks = np.ascontiguousarray(np.atleast_1d(ks), np.int32)
    out = np.empty((rows.shape[0], cols.shape[0]), np.float64)
    rc = lib().orc_spectrum_block(pr, C.c_int64(rows.shape[0]), pc, C.c_int64(cols.shape[0]), C.c_int(rows.shape[1]),
                                  ks.ctypes.data_as(C.c_void_p), C.c_int(ks.size),
                                  out.ctypes.data_as(C.c_void_p), C.c_int64(out.shape[1]))
    assert rc == 0
    return out


def mismatch_raw_block(rows, cols, k, m):
    from oracle_np import mismatch_table
    rows, pr = _u8(rows); cols, pc = _u8(cols)
    T = np.array(mismatch_table(k, m), np.int64)
    out = np.empty((rows.shape[0], cols.shape[0]), np.int64)
    rc = lib().orc_mismatch_raw_block(pr, C.c_int64(rows.shape[0]), pc, C.c_int64(cols.shape[0]), C.c_int(rows.shape[1]),
                                      C.c_int(k), T.ctypes.data_as(C.c_void_p),
                                      out.ctypes.data_as(C.c_void_p), C.c_int64(out.shape[1]))
    assert rc == 0
    return out


def wd_block(rows, cols, d, row_index0=0, col_index0=0):
    rows, pr = _u8(rows); cols, pc = _u8(cols)
    out = np.empty((rows.shape[0], cols.shape[0]), np.float64)
    rc = lib().orc_wd_block(pr, C.c_int64(rows.shape[0]), C.c_int64(row_index0), pc, C.c_int64(cols.shape[0]),
                            C.c_int64(col_index0), C.c_int(rows.shape[1]), C.c_int(d),
                            out.ctypes.data_as(C.c_void_p), C.c_int64(out.shape[1]))
    assert rc == 0
    return out


def la_block(rows, cols, e, d, beta, smith=0, row_index0=0, col_index0=0):
    rows, pr = _u8(rows); cols, pc = _u8(cols)
    out = np.empty((rows.shape[0], cols.shape[0]), np.float64)
    rc = lib().orc_la_block(pr, C.c_int64(rows.shape[0]), C.c_int64(row_index0), pc, C.c_int64(cols.shape[0]),
                            C.c_int64(col_index0), C.c_int(rows.shape[1]), C.c_double(e), C.c_double(d),
                            C.c_double(beta), C.c_int(smith), out.ctypes.data_as(C.c_void_p), C.c_int64(out.shape[1]))
    assert rc == 0
    return out
