"""
_cabi.py -- ctypes binding of libkmg.so (C-ABI declared in include/kmg.h).

The library is built in-tree (`make -C kernel-methods-for-genomics_b200/csrc`, or
`python -c "import __graft_entry__ as g; g.build()"`) and lives next to this package.  There is no
fallback: if the shared object is missing, or there is no CUDA device at call time, the calls fail
loudly (ImportError / KmgError).
"""
import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(PKG_DIR, "libkmg.so")

KMG_OK = 0
KMG_ERR_CUDA, KMG_ERR_ARG, KMG_ERR_ALPHABET, KMG_ERR_UNSUPPORTED, KMG_ERR_NOMEM = -1, -2, -3, -4, -5
KMG_OUT_S32, KMG_OUT_F64 = 0, 1
KMG_SEQ_CODES, KMG_SEQ_ASCII = 0, 1
KMG_MM_AUTO, KMG_MM_PAIRWISE, KMG_MM_DENSE = 0, 1, 2
KMG_EXCH_SINGLE, KMG_EXCH_STAGED, KMG_EXCH_DIRECT = 0, 1, 2
KMG_EXCH_DEFER_JOIN = 0x100


class KmgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libkmg error {code}: {msg}")
        self.code = code


_i64, _i32, _vp, _dbl = C.c_int64, C.c_int, C.c_void_p, C.c_double

# name -> (restype, argtypes); mirrors include/kmg.h one to one
PROTOTYPES = {
    "kmg_version": (_i32, []),
    "kmg_last_error": (C.c_char_p, []),
    "kmg_device_count": (_i32, []),
    "kmg_set_device": (_i32, [_i32]),
    "kmg_release": (_i32, []),
    "kmg_dev_malloc": (_i32, [_i64, _vp]),
    "kmg_dev_free": (_i32, [_vp]),
    "kmg_dev_upload": (_i32, [_vp, _vp, _i64]),
    "kmg_dev_download": (_i32, [_vp, _vp, _i64]),
    "kmg_host_alloc": (_i32, [_i64, _vp]),
    "kmg_host_free": (_i32, [_vp]),
    "kmg_set_d2h_mode": (_i32, [_i32]),
    "kmg_get_d2h_mode": (_i32, []),
    "kmg_spectrum_host": (_i32, [_vp, _i64, _vp, _i64, _i32, _i32, _vp, _i32, _vp, _i64]),
    "kmg_mismatch_host": (_i32, [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i64]),
    "kmg_spectrum_phi_host": (_i32, [_vp, _i64, _i32, _i32, _vp, _i32, _vp, _i64]),
    "kmg_mismatch_phi_host": (_i32, [_vp, _i64, _i32, _i32, _i32, _i32, _vp, _i64]),
    "kmg_wd_host": (_i32, [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _vp, _i64]),
    "kmg_wds_host": (_i32, [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _i64]),
    "kmg_wds_dev": (_i32, [_vp, _vp, _i64, _i64, _i64, _i64, _i32, _i32, _i32, _vp, _i64, _i32, _vp]),
    "kmg_la_host": (_i32, [_vp, _i64, _vp, _i64, _i32, _i32, _dbl, _dbl, _dbl, _i32, _vp, _i64]),
    "kmg_normalize_host": (_i32, [_vp, _i64, _i64]),
    "kmg_center_host": (_i32, [_vp, _i64, _i64, _vp, _i64]),
    "kmg_combine_host": (_i32, [_vp, _i32, _i64, _vp, _i32, _i32, _vp]),
    "kmg_alignf_stats_host": (_i32, [_vp, _i32, _i64, _vp, _i64, _vp, _vp, _vp]),
    "kmg_nlck_grad_host": (_i32, [_vp, _i32, _i64, _vp, _vp, _i32, _vp]),
    "kmg_alignf_fused_host": (_i32, [_vp, _i64, _i32, _i32, _vp, _i32, _vp, _i64, _vp, _vp, _vp, _vp]),
    "kmg_combine_fused_host": (_i32, [_vp, _i64, _i32, _i32, _vp, _i32, _vp, _i32, _i32, _i32, _vp, _i64, _vp]),
    "kmg_build_grams_dev": (_i32, [_vp, _i64, _i32, _i32, _vp, _i32, _vp, _i64, _i32, _vp]),
    "kmg_pack_dev": (_i32, [_vp, _i32, _i64, _i32, _vp, _vp, _vp]),
    "kmg_spectrum_phi_width": (_i64, [_vp, _i32]),
    "kmg_spectrum_phi_dev": (_i32, [_vp, _i64, _i32, _vp, _i32, _vp, _i64, _vp]),
    "kmg_mismatch_phi_dev": (_i32, [_vp, _i64, _i32, _i32, _i32, _vp, _i64, _vp]),
    "kmg_phi_diag_sqrt_dev": (_i32, [_vp, _i64, _i64, _i64, _vp, _vp]),
    "kmg_gram_i8_dev": (_i32, [_vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _i64, _i32, _i32, _vp, _vp, _i32, _vp]),
    "kmg_gram_sharded_stage_bytes": (_i32, [_i32, _vp, _i32, _i32, _vp]),
    "kmg_gram_sharded_join": (_i32, [_vp]),
    "kmg_gram_sharded_mark": (_i32, [_i32]),
    "kmg_gram_sharded_wait_mark": (_i32, [_i32, _vp]),
    "kmg_gram_sharded_launches": (_i32, [_i32, _vp, _i32, _i32, _vp]),
    "kmg_gram_i8_sharded_dev": (_i32, [_vp, _i64, _i64, _i64, _i32, _i32, _vp, _vp, _i64, _i32, _vp, _i32, _vp, _vp, _vp]),
    "kmg_mma_peak_i8_dev": (_i32, [_i32, _vp, _vp]),
    "kmg_alu_peak_dev": (_i32, [_i32, _i32, _vp, _vp]),
    "kmg_gram_sharded_takes_host": (_i32, [_i32, _vp, _i32, _i32, _i64, _i64]),
    "kmg_ipc_export": (_i32, [_vp, _vp]),
    "kmg_ipc_open": (_i32, [_vp, _vp]),
    "kmg_ipc_close": (_i32, [_vp]),
    "kmg_gram_i8_simt_dev": (_i32, [_vp, _vp, _i64, _i64, _i64, _i64, _vp, _i64, _vp]),
    "kmg_mismatch_dev": (_i32, [_vp, _vp, _i64, _i64, _i64, _i64, _i32, _i32, _i32, _vp, _i64, _i32, _i32, _vp, _vp, _vp]),
    "kmg_mismatch_diag_dev": (_i32, [_vp, _i64, _i32, _i32, _i32, _vp, _vp]),
    "kmg_wd_dev": (_i32, [_vp, _vp, _i64, _i64, _i64, _i64, _i32, _i32, _vp, _i64, _i32, _vp]),
    "kmg_la_dev": (_i32, [_vp, _vp, _i64, _i64, _i64, _i64, _i32, _dbl, _dbl, _dbl, _i32, _vp, _i64, _i32, _vp]),
    "kmg_normalize_dev": (_i32, [_vp, _i64, _i64, _vp, _vp]),
    "kmg_center_workspace_bytes": (_i64, [_i64]),
    "kmg_center_dev": (_i32, [_vp, _i64, _i64, _vp, _i64, _vp, _vp]),
    "kmg_row_sums_dev": (_i32, [_vp, _i64, _i64, _i64, _vp, _vp]),
    "kmg_col_sums_workspace_bytes": (_i64, [_i64, _i64]),
    "kmg_col_sums_dev": (_i32, [_vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    "kmg_center_apply_dev": (_i32, [_vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _i64, _vp]),
    "kmg_gather_dev": (_i32, [_vp, _i64, _vp, _i64, _vp, _i64, _vp]),
    "kmg_combine_dev": (_i32, [_vp, _vp, _vp, _i32, _i32, _i64, _i64, _vp, _i64, _vp]),
    "kmg_matvec_dev": (_i32, [_vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    "kmg_spd_solve_workspace_bytes": (_i64, [_i64]),
    "kmg_spd_solve_dev": (_i32, [_vp, _i64, _i64, _vp, _dbl, _vp, _vp, _vp, _vp]),
    "kmg_spd_solve_host": (_i32, [_vp, _i64, _i64, _vp, _i64, _vp, _dbl, _vp, _vp]),
    "kmg_weighted_dot_dev": (_i32, [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _vp]),
    "kmg_mismatch_table_host": (_i32, [_i32, _i32, _vp]),
}

_lib = None


def lib():
    """Load libkmg.so once and attach the prototypes.  Raises ImportError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build the CUDA extension first "
                "(make -C kernel-methods-for-genomics_b200/csrc). There is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error():
    return lib().kmg_last_error().decode("utf-8", "replace")


def check(rc):
    """Raise KmgError (ValueError for alphabet / argument errors, like the reference's int() failure)."""
    if rc >= 0:
        return rc
    msg = last_error()
    if rc in (KMG_ERR_ALPHABET, KMG_ERR_ARG):
        raise ValueError(f"libkmg: {msg}")
    raise KmgError(rc, msg)
