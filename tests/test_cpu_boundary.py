"""CPU: the C-ABI library loads and exports every symbol include/kmg.h declares; the host-side shim has the
reference's surface; without a GPU every compute entry point fails loudly (no CPU fallback)."""
import ctypes as C
import inspect
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "kmg.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kmg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from kmg import _cabi
    lib = _cabi.lib()
    names = _declared_symbols()
    assert len(names) >= 38
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/kmg.h but not exported by libkmg.so"
        assert name in _cabi.PROTOTYPES, f"{name} has no ctypes prototype in kmg/_cabi.py"
    assert set(_cabi.PROTOTYPES) == set(names)
    assert lib.kmg_version() >= 100
    # only libkmg's own symbols are bound: nothing from oracle/ is reachable from the product package
    import kmg
    src = "".join(open(os.path.join(os.path.dirname(kmg.__file__), f)).read() for f in os.listdir(os.path.dirname(kmg.__file__)) if f.endswith(".py"))
    assert "oracle" not in src.replace("oracle functions in the CPU tests", "")


def test_reference_surface():
    """Same names and signatures as the reference's kernels.py."""
    import kernels as km
    sig = {n: list(inspect.signature(getattr(km, n)).parameters) for n in (
        "select_method", "get_spectrum_K", "get_WD_K", "get_mismatch_K", "get_LA_K", "center_K", "normalize_K", "beta", "format",
        "letter_to_num", "get_phi_u", "get_phi_km", "get_WD_d", "affine_align", "Smith_Waterman")}
    assert sig["select_method"] == ["X", "method"]
    assert sig["get_spectrum_K"] == ["X", "k"] and sig["get_WD_K"] == ["X", "d"] and sig["get_mismatch_K"] == ["X", "k", "m"]
    assert sig["get_LA_K"] == ["X", "e", "d", "beta", "smith", "eig"]
    assert inspect.signature(km.get_LA_K).parameters["beta"].default == 0.5
    assert sig["get_WD_d"] == ["x", "y", "d", "L"] and sig["affine_align"] == ["x", "y", "e", "d", "beta"]
    assert km.S.tolist() == [[4, 0, 0, 0], [0, 9, -3, -1], [0, -3, 6, 2], [0, -1, -2, 5]]
    assert km.beta(10, 3) == 2 * (10 - 3 + 1) / 10 / 11
    assert km.letter_to_num("ACGT") == "1234" and km.format("TTA").tolist() == [4, 4, 1]


def test_mismatch_table_host_utility():
    import oracle_np as onp
    from kmg import host
    for k, m in ((10, 1), (10, 2), (4, 2), (5, 0), (16, 3), (101, 1)):
        assert host.mismatch_table(k, m).tolist() == onp.mismatch_table(k, m), (k, m)


def test_no_cpu_fallback():
    """On a box without a GPU the product path must fail, not compute."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from kmg import KmgError, host
    codes = np.zeros((4, 101), np.uint8)
    for call in (lambda: host.spectrum_gram(codes, 3), lambda: host.wd_gram(codes, 4), lambda: host.mismatch_gram(codes, 10, 1),
                 lambda: host.la_gram(codes, -11, -1, 0.5), lambda: host.center(np.eye(3)), lambda: host.combine([np.eye(3)], [1.0])):
        with pytest.raises(KmgError) as e:
            call()
        assert e.value.code == -1 and "no CPU fallback" in str(e.value)
    K = np.eye(3) * 2
    with pytest.raises(KmgError):
        host.normalize_inplace(K)


def test_argument_errors_are_value_errors():
    from kmg import host
    with pytest.raises(ValueError):
        host.as_seq_buffer(["ACGT", "ACG"])
    with pytest.raises(ValueError):
        host.normalize_inplace(np.zeros((3, 4)))
    buf, fmt = host.as_seq_buffer(["ACGT", "TTTT"])
    assert buf.shape == (2, 4) and fmt == 1
    buf, fmt = host.as_seq_buffer(np.zeros((2, 5), np.uint8))
    assert fmt == 0


def test_result_blocks_are_recycled():
    """kmg_host_alloc / kmg_host_free (the memory behind every large result array): 2 MB aligned, kept alive by numpy
    views, handed out again after the last view dies, unmapped by kmg_release, foreign pointers rejected."""
    import gc
    from kmg import _cabi, host
    a = host._result((1024, 1024))
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.flags["WRITEABLE"] and a.ctypes.data % (1 << 21) == 0
    a[:] = 3.0
    addr = a.ctypes.data
    view = a[5:7]
    del a
    gc.collect()
    b = host._result((1024, 1024))                  # the first block is still referenced by `view`
    assert b.ctypes.data != addr and view.sum() == 2 * 1024 * 3.0
    del view
    gc.collect()
    c = host._result((1000, 1024))                  # slightly smaller request: same recycled block
    assert c.ctypes.data == addr
    small = host._result((4, 4))
    assert small.base is None and not small.any()   # small results stay plain zeroed numpy arrays
    with pytest.raises(ValueError):
        _cabi.check(_cabi.lib().kmg_host_free(C.c_void_p(b.ctypes.data + 64)))
    del b, c
    gc.collect()
    assert _cabi.lib().kmg_release() == 0


def test_select_method_parses_the_reference_dsl(monkeypatch, capsys):
    """kernels.select_method (kernels.py:461-505): every method string of the reference's mini-language reaches the
    right builder with the right arguments (builders stubbed: no GPU)."""
    import kernels as km
    calls = []
    for name in ("get_spectrum_K", "get_WD_K", "get_mismatch_K", "get_LA_K", "get_WDShifts_K", "get_string_K", "get_gappy_K"):
        monkeypatch.setattr(km, name, (lambda n: (lambda X, *a: calls.append((n, a)) or n))(name))
    for method in ("SP_k6", "WD_d10", "WDS_d3_s2", "MM_k10_m1", "LA_e11_d1_b0.5_smith0_eig1", "LA_e11_d1_b0.5_smith1_eig0",
                   "SS_l0.5_k3", "GP_k5_g1"):
        km.select_method("X", method)
    assert calls == [("get_spectrum_K", (6,)), ("get_WD_K", (10,)), ("get_WDShifts_K", (3, 2)), ("get_mismatch_K", (10, 1)),
                     ("get_LA_K", (11.0, 1.0, 0.5, 0, 1)), ("get_LA_K", (11.0, 1.0, 0.5, 1, 0)), ("get_string_K", (0.5, 3)),
                     ("get_gappy_K", (5, 1))]
    assert capsys.readouterr().out == "['WD', 'd10']\n"      # the reference echoes the split string for WD only
    with pytest.raises(NotImplementedError):
        km.select_method("X", "XX_k1")
    assert km.letter_to_num("ACGTN") == "1234N" and km.format("GATTACA").tolist() == [3, 1, 4, 4, 1, 2, 1]
    with pytest.raises(ValueError):
        km.format("ACGN")


def test_fused_method_parser_follows_the_reference_dsl():
    """kmg.fused.parse_method reads the reference's method strings with select_method's rule (kernels.py:479-502: the
    first character of every '_' field is dropped; 'smith' / 'eig' words for LA); kernels outside the hot path raise."""
    from kmg import fused
    m = fused.parse_method("SP_k6")
    assert (m.kind, m.k) == (fused.KIND_SP, 6)
    m = fused.parse_method("MM_k10_m1")
    assert (m.kind, m.k, m.m) == (fused.KIND_MM, 10, 1)
    m = fused.parse_method("WD_d10")
    assert (m.kind, m.d) == (fused.KIND_WD, 10)
    m = fused.parse_method("WDS_d3_s2")
    assert (m.kind, m.d, m.S) == (fused.KIND_WDS, 3, 2)
    m = fused.parse_method("LA_e11_d1_b0.5_smith1_eig0")
    assert (m.kind, m.e, m.dd, m.beta, m.smith) == (fused.KIND_LA, 11.0, 1.0, 0.5, 1)
    for bad in ("SS_l1_k3", "GP_k3_g1", "XX_1"):
        with pytest.raises(NotImplementedError):
            fused.parse_method(bad)


def test_solver_dropins_import_without_a_gpu():
    """KRR.py / KLR.py / ALIGNF.py / NLCKernels.py mirror the reference's classes and import on a CPU-only box; the first
    device call is what fails (no CPU fallback)."""
    import numpy as np
    import ALIGNF
    import KLR
    import KRR
    import NLCKernels
    from kmg import _cabi
    assert hasattr(NLCKernels, "cross_validation") and hasattr(ALIGNF, "aligned_kernels")
    assert hasattr(ALIGNF.ALIGNF, "from_sequences") and hasattr(NLCKernels.NLCK, "from_sequences")
    m = KRR.KRR(np.eye(4), np.arange(4), lbda=0.5)
    assert (m.lbda, m.eps) == (0.5, 1e-5)
    k = KLR.KLR(np.eye(4), np.arange(4))
    assert abs(k.sigmoid(0.0) - 0.5) < 1e-16
    import pandas as pd
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(_cabi.KmgError):
            m.fit(pd.DataFrame({"Id": [0, 1]}), pd.DataFrame({"Id": [0, 1], "Bound": [1.0, -1.0]}))


def test_header_is_plain_c_and_a_c_program_links_the_library(tmp_path):
    """The boundary is a C ABI: include/kmg.h compiles as strict C99 and as C++11, and a C program -- no Python, no torch --
    links libkmg.so and calls host-side entry points (a compute entry point without a GPU reports an error, it does not
    fall back to the CPU)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("needs gcc")
    inc = os.path.join(ROOT, "include")
    libdir = os.path.join(ROOT, "kernel-methods-for-genomics_b200")
    src = tmp_path / "t.c"
    src.write_text(r'''
#include <stdio.h>
#include <stdint.h>
#include "kmg.h"
int main(void) {
    int64_t T[11];
    int ks[2] = {3, 6};
    if (kmg_mismatch_table_host(10, 1, T) != 0) return 2;
    printf("T0=%lld T1=%lld width=%d\n", (long long)T[0], (long long)T[1], (int)kmg_spectrum_phi_width(ks, 2));
    {
        unsigned char seqs[2 * 8] = {0,1,2,3,0,1,2,3, 3,2,1,0,3,2,1,0};
        double K[4];
        int k1[1] = {2};
        int rc = kmg_spectrum_host(seqs, 2, NULL, 0, 8, KMG_SEQ_CODES, k1, 1, K, 2);
        printf("rc=%d err=%s\n", rc, kmg_last_error());
    }
    return 0;
}
''')
    for compiler, std, name in (("gcc", "-std=c99", "t.c"), ("g++", "-std=c++11", "t.cpp")):
        if name == "t.cpp":
            (tmp_path / name).write_text(src.read_text())
        subprocess.check_call([compiler, std, "-Wall", "-Wextra", "-pedantic", "-Werror", f"-I{inc}", "-fsyntax-only", str(tmp_path / name)])
    exe = tmp_path / "t"
    subprocess.check_call(["gcc", "-std=c99", f"-I{inc}", str(src), "-o", str(exe), f"-L{libdir}", "-lkmg", f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    first, second = out.stdout.strip().splitlines()
    assert first == "T0=31 T1=4 width=4224", first          # (10,1): 1 + 3*10 neighbours; distance 1: the two k-mers + 2 letters; pad128(4^3 + 4^6)
    import torch
    if not torch.cuda.is_available():
        assert second.startswith("rc=-") and "err=" in second and len(second) > len("rc=-1 err="), second
