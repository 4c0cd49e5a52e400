"""
device.py -- device-resident layer over the `*_dev` entry points of libkmg.so.

PyTorch is used for plumbing only (device memory, streams, torch.distributed in dist.py): tensors
are allocated with torch, their raw device pointers and the current CUDA stream are handed to the
C-ABI, and all arithmetic happens in the hand-written sm_100a kernels.  Nothing here falls back to
torch ops for the Gram entries.
"""
import ctypes as C

import numpy as np
import torch

from . import _cabi
from ._cabi import KMG_OUT_F64, KMG_OUT_S32, KMG_SEQ_ASCII, KMG_SEQ_CODES, check


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _dev():
    if not torch.cuda.is_available():
        raise _cabi.KmgError(_cabi.KMG_ERR_CUDA, "no CUDA device available: libkmg has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def pack(seq_bytes, seq_format=KMG_SEQ_CODES):
    """(n, L) uint8 host array or device tensor -> (n, 8) int32 device tensor of bit-planes."""
    dev = _dev()
    if isinstance(seq_bytes, np.ndarray):
        seq_bytes = torch.from_numpy(np.ascontiguousarray(seq_bytes)).to(dev, non_blocking=True)
    seq_bytes = seq_bytes.contiguous()
    n, L = seq_bytes.shape
    planes = torch.empty((n, 8), dtype=torch.int32, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    check(_cabi.lib().kmg_pack_dev(_p(seq_bytes), seq_format, n, L, _p(planes), _p(err), _stream()))
    if int(err.item()) != 0:
        raise ValueError("libkmg: sequence contains a character outside {A,C,G,T}")
    return planes


def phi_width(ks):
    ks = np.ascontiguousarray(np.atleast_1d(ks), np.int32)
    return int(_cabi.lib().kmg_spectrum_phi_width(ks.ctypes.data_as(C.c_void_p), ks.size))


def spectrum_phi(planes, L, ks, out=None):
    """planes -> int8 Phi (n, W) with W = pad128(sum_k 4^k)."""
    ks = np.ascontiguousarray(np.atleast_1d(ks), np.int32)
    n = planes.shape[0]
    W = phi_width(ks)
    if out is None:
        out = torch.empty((n, W), dtype=torch.int8, device=planes.device)
    check(_cabi.lib().kmg_spectrum_phi_dev(_p(planes), n, L, ks.ctypes.data_as(C.c_void_p), ks.size, _p(out), W, _stream()))
    return out


def mismatch_phi(planes, L, k, m):
    n = planes.shape[0]
    W = (4 ** k + 127) // 128 * 128
    out = torch.empty((n, W), dtype=torch.int8, device=planes.device)
    check(_cabi.lib().kmg_mismatch_phi_dev(_p(planes), n, L, k, m, _p(out), W, _stream()))
    return out


def phi_diag_sqrt(phi):
    n, W = phi.shape
    sd = torch.empty(n, dtype=torch.float64, device=phi.device)
    check(_cabi.lib().kmg_phi_diag_sqrt_dev(_p(phi), n, W, phi.stride(0), _p(sd), _stream()))
    return sd


def _out(rows, cols, dtype, device, out):
    if out is not None:
        return out
    return torch.empty((rows, cols), dtype=torch.float64 if dtype == KMG_OUT_F64 else torch.int32, device=device)


def gram_i8(phi_rows, phi_cols, row_index0=0, col_index0=0, out_dtype=KMG_OUT_F64, symmetric=False,
            sd_rows=None, sd_cols=None, m_sub=0, out=None):
    """Block of K = Phi_rows Phi_cols^T on the tensor cores."""
    rows, W = phi_rows.shape
    cols = phi_cols.shape[0]
    out = _out(rows, cols, out_dtype, phi_rows.device, out)
    check(_cabi.lib().kmg_gram_i8_dev(_p(phi_rows), _p(phi_cols), rows, cols, W, phi_rows.stride(0), row_index0, col_index0,
                                      _p(out), out.stride(0), out_dtype, 1 if symmetric else 0, _p(sd_rows), _p(sd_cols),
                                      m_sub, _stream()))
    return out


def sharded_stage_bytes(part_row0, part, out_dtype=KMG_OUT_F64):
    bounds = np.ascontiguousarray(part_row0, np.int64)
    out = C.c_int64(0)
    check(_cabi.lib().kmg_gram_sharded_stage_bytes(bounds.size - 1, bounds.ctypes.data_as(C.c_void_p), int(part), out_dtype, C.byref(out)))
    return out.value


def sharded_launches(part_row0, part, exchange):
    """GEMM launches one gram_i8_sharded call enqueues for this part."""
    bounds = np.ascontiguousarray(part_row0, np.int64)
    mode = {"single": _cabi.KMG_EXCH_SINGLE, "staged": _cabi.KMG_EXCH_STAGED, "direct": _cabi.KMG_EXCH_DIRECT}[exchange]
    out = C.c_int(0)
    check(_cabi.lib().kmg_gram_sharded_launches(bounds.size - 1, bounds.ctypes.data_as(C.c_void_p), int(part), mode, C.byref(out)))
    return out.value


def sharded_join():
    """Make the current stream wait for the peer copies of the staged exchanges enqueued with defer_join=True."""
    check(_cabi.lib().kmg_gram_sharded_join(_stream()))


def sharded_mark(slot):
    """Remember the peer copies enqueued so far (slot 0..3)."""
    check(_cabi.lib().kmg_gram_sharded_mark(int(slot)))


def sharded_wait_mark(slot):
    """The current stream waits for the copies remembered in `slot` (and for nothing enqueued after them)."""
    check(_cabi.lib().kmg_gram_sharded_wait_mark(int(slot), _stream()))


def gram_i8_sharded(phi, part_row0, part, part_ptrs, ldo, out_dtype=KMG_OUT_F64, sd=None, stage=None, exchange=None, defer_join=False):
    """Part `part`'s launch of the sharded symmetric Gram of all rows of `phi` (kmg_gram_i8_sharded_dev).
    part_row0: len(parts)+1 boundaries; part_ptrs: device address (int) of every part's block-row buffer (row stride ldo
    elements) -- this device's own allocations or peer memory opened through CUDA IPC.  exchange: "direct" (one launch per
    peer block, the epilogue's TMA stores write the owner's buffer), "staged" (stage = device address of
    sharded_stage_bytes() bytes of local staging: one launch + one peer copy per peer block) or "single" (one launch,
    thread-issued stores into the peers' buffers); default: "staged" when `stage` is given, else "single".
    Returns the entries computed."""
    if exchange is None:
        exchange = "staged" if stage is not None else "single"
    mode = {"single": _cabi.KMG_EXCH_SINGLE, "staged": _cabi.KMG_EXCH_STAGED, "direct": _cabi.KMG_EXCH_DIRECT}[exchange]
    if defer_join and exchange == "staged":
        mode |= _cabi.KMG_EXCH_DEFER_JOIN
    n, W = phi.shape
    g = len(part_ptrs)
    bounds = np.ascontiguousarray(part_row0, np.int64)
    assert bounds.size == g + 1
    ptrs = (C.c_void_p * g)(*[int(p) for p in part_ptrs])
    computed = C.c_int64(0)
    check(_cabi.lib().kmg_gram_i8_sharded_dev(_p(phi), n, W, phi.stride(0), g, int(part), bounds.ctypes.data_as(C.c_void_p),
                                              ptrs, int(ldo), out_dtype, _p(sd), mode,
                                              None if stage is None else C.c_void_p(int(stage)), C.byref(computed), _stream()))
    return computed.value


def mma_peak_i8(iters=20000, repeats=3):
    """Measured int8 tensor-core peak in TOP/s (kmg_mma_peak_i8_dev timed with CUDA events on the current stream)."""
    _dev()
    ops = C.c_int64(0)
    best = 0.0
    for _ in range(repeats + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(_cabi.lib().kmg_mma_peak_i8_dev(int(iters), C.byref(ops), _stream()))
        e1.record()
        torch.cuda.synchronize()
        best = max(best, ops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


ALU_PEAK_KINDS = {"lop3": 0, "shf": 1, "popc": 2, "dfma": 3, "dadd": 4, "dmul": 5}


def alu_peak(kind, iters=4000, repeats=3):
    """Measured issue peak of one CUDA-core instruction in thread-level instructions per second (kmg_alu_peak_dev timed
    with CUDA events on the current stream): "lop3" / "shf" (INT32 ALU pipe), "popc" (XU), "dfma" / "dadd" / "dmul" (FP64)."""
    _dev()
    ops = C.c_int64(0)
    best = 0.0
    for _ in range(repeats + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(_cabi.lib().kmg_alu_peak_dev(ALU_PEAK_KINDS[kind], int(iters), C.byref(ops), _stream()))
        e1.record()
        torch.cuda.synchronize()
        best = max(best, ops.value / (e0.elapsed_time(e1) * 1e-3))
    return best


def gram_i8_simt(phi_rows, phi_cols):
    rows, W = phi_rows.shape
    cols = phi_cols.shape[0]
    out = torch.empty((rows, cols), dtype=torch.int32, device=phi_rows.device)
    check(_cabi.lib().kmg_gram_i8_simt_dev(_p(phi_rows), _p(phi_cols), rows, cols, W, phi_rows.stride(0), _p(out), out.stride(0), _stream()))
    return out


def mismatch_block(planes_rows, planes_cols, L, k, m, row_index0=0, col_index0=0, out_dtype=KMG_OUT_F64, symmetric=False,
                   sd_rows=None, sd_cols=None, out=None):
    rows, cols = planes_rows.shape[0], planes_cols.shape[0]
    out = _out(rows, cols, out_dtype, planes_rows.device, out)
    check(_cabi.lib().kmg_mismatch_dev(_p(planes_rows), _p(planes_cols), rows, cols, row_index0, col_index0, L, k, m,
                                       _p(out), out.stride(0), out_dtype, 1 if symmetric else 0, _p(sd_rows), _p(sd_cols), _stream()))
    return out


def mismatch_diag_sqrt(planes, L, k, m):
    n = planes.shape[0]
    sd = torch.empty(n, dtype=torch.float64, device=planes.device)
    check(_cabi.lib().kmg_mismatch_diag_dev(_p(planes), n, L, k, m, _p(sd), _stream()))
    return sd


def wd_block(planes_rows, planes_cols, L, d, row_index0=0, col_index0=0, symmetric=False, out=None):
    rows, cols = planes_rows.shape[0], planes_cols.shape[0]
    out = _out(rows, cols, KMG_OUT_F64, planes_rows.device, out)
    check(_cabi.lib().kmg_wd_dev(_p(planes_rows), _p(planes_cols), rows, cols, row_index0, col_index0, L, d,
                                 _p(out), out.stride(0), 1 if symmetric else 0, _stream()))
    return out


def wds_block(planes_rows, planes_cols, L, d, S, row_index0=0, col_index0=0, symmetric=False, out=None):
    rows, cols = planes_rows.shape[0], planes_cols.shape[0]
    out = _out(rows, cols, KMG_OUT_F64, planes_rows.device, out)
    check(_cabi.lib().kmg_wds_dev(_p(planes_rows), _p(planes_cols), rows, cols, row_index0, col_index0, L, d, S,
                                  _p(out), out.stride(0), 1 if symmetric else 0, _stream()))
    return out


def la_block(planes_rows, planes_cols, L, e, d, beta, smith=0, row_index0=0, col_index0=0, symmetric=False, out=None):
    rows, cols = planes_rows.shape[0], planes_cols.shape[0]
    out = _out(rows, cols, KMG_OUT_F64, planes_rows.device, out)
    check(_cabi.lib().kmg_la_dev(_p(planes_rows), _p(planes_cols), rows, cols, row_index0, col_index0, L,
                                 float(e), float(d), float(beta), int(smith), _p(out), out.stride(0), 1 if symmetric else 0, _stream()))
    return out


def normalize_(K):
    """normalize_K (kernels.py:398-415) in place on a device-resident Gram, including the reference's early-out: a matrix
    whose K[0,0] is exactly 1 is returned untouched (kernels.py:404-405)."""
    n = K.shape[0]
    if n == 0 or float(K[0, 0].item()) == 1.0:
        return K
    sd = torch.empty(n, dtype=torch.float64, device=K.device)
    check(_cabi.lib().kmg_normalize_dev(_p(K), n, K.stride(0), _p(sd), _stream()))
    return K


def center(K):
    n = K.shape[0]
    ws = torch.empty(int(_cabi.lib().kmg_center_workspace_bytes(n)), dtype=torch.uint8, device=K.device)
    out = torch.empty_like(K)
    check(_cabi.lib().kmg_center_dev(_p(K), n, K.stride(0), _p(out), out.stride(0), _p(ws), _stream()))
    return out


def row_sums(K):
    rows, cols = K.shape
    rs = torch.empty(rows, dtype=torch.float64, device=K.device)
    check(_cabi.lib().kmg_row_sums_dev(_p(K), rows, cols, K.stride(0), _p(rs), _stream()))
    return rs


def col_sums(K):
    rows, cols = K.shape
    cs = torch.empty(cols, dtype=torch.float64, device=K.device)
    ws = torch.empty(max(int(_cabi.lib().kmg_col_sums_workspace_bytes(rows, cols)), 8), dtype=torch.uint8, device=K.device)
    check(_cabi.lib().kmg_col_sums_dev(_p(K), rows, cols, K.stride(0), _p(cs), _p(ws), _stream()))
    return cs


def center_apply(K, rs, cs, g, n_total):
    rows, cols = K.shape
    out = torch.empty_like(K)
    check(_cabi.lib().kmg_center_apply_dev(_p(K), rows, cols, n_total, K.stride(0), _p(rs), _p(cs), _p(g), _p(out), out.stride(0), _stream()))
    return out


def combine(Ks, u, degree=1):
    p = len(Ks)
    rows, cols = Ks[0].shape
    ptrs = (C.c_void_p * p)(*[k.data_ptr() for k in Ks])
    lds = np.array([k.stride(0) for k in Ks], np.int64)
    u = np.ascontiguousarray(u, np.float64)
    out = torch.empty((rows, cols), dtype=torch.float64, device=Ks[0].device)
    check(_cabi.lib().kmg_combine_dev(ptrs, lds.ctypes.data_as(C.c_void_p), u.ctypes.data_as(C.c_void_p), p, int(degree),
                                      rows, cols, _p(out), out.stride(0), _stream()))
    return out


def weighted_dot(A, B=None, w=None):
    n = A.shape[0]
    part = torch.empty(n, dtype=torch.float64, device=A.device)
    res = torch.empty(1, dtype=torch.float64, device=A.device)
    check(_cabi.lib().kmg_weighted_dot_dev(_p(A), A.stride(0), _p(B), 0 if B is None else B.stride(0), _p(w), n, _p(part), _p(res), _stream()))
    return res
