"""GPU: the fused ALIGNF / NLCK entry points (csrc/fused.cu, kmg/fused.py) -- sequences in, statistics / combination out,
no Gram on the host -- against the reference's golden vectors (tests/golden/ref_vectors.npz: ALIGNF a / M / get_K and NLCK
grad / get_K produced by the unmodified ALIGNF.py / NLCKernels.py) and against the array-based path on larger inputs."""
import contextlib
import io

import numpy as np
import pytest

import oracle_np as onp

pytestmark = pytest.mark.gpu
METHODS = ["SP_k3", "WD_d5", "MM_k3_m1"]  # the kernels oracle/gen_golden.py handed to the reference's ALIGNF / NLCK


@pytest.fixture(scope="module")
def fused():
    torch = pytest.importorskip("torch")
    assert torch.cuda.is_available(), "these tests need a GPU"
    from kmg import fused as f
    return f


def test_alignf_golden_through_the_fused_call(fused, golden, dna):
    """a and M at 1e-12 (normwise), get_K bit for bit -- from SEQUENCES, through kmg_alignf_fused_host / kmg_combine_fused_host."""
    codes, _ = dna
    c = codes[:96]
    a, M = fused.alignf_stats(c, METHODS, golden["alignf_fit_rows"], golden["alignf_y"])
    assert np.abs(a - golden["alignf_a"]).max() <= 1e-12 * np.abs(golden["alignf_a"]).max()
    assert np.abs(M - golden["alignf_M"]).max() <= 1e-12 * np.abs(golden["alignf_M"]).max()
    assert np.array_equal(M, M.T)
    rep = fused.last_report
    assert rep["fused"]["d2h"] == (3 + 9) * 8 and rep["fused"]["h2d"] == 96 * 101 + 64 * 16
    assert rep["array_based"]["d2h"] > 1000 * rep["fused"]["d2h"]
    Km = fused.combine(c, METHODS, golden["alignf_u"])
    assert np.array_equal(Km, golden["alignf_Km"])


def test_nlck_golden_through_the_fused_calls(fused, golden, dna):
    """NLCK.get_K for degree 1..3 (bit-exact for degree <= 2, 1e-12 for the pow() of degree 3) and NLCK.grad on fit
    sub-blocks that were built, normalised and kept on the device."""
    from kmg import resident as res
    codes, _ = dna
    c = codes[:96]
    u = golden["nlck_u"]
    for deg in (1, 2, 3):
        Km = fused.combine(c, METHODS, u, degree=deg, normalize_inputs=True, normalize=True)
        want = golden[f"nlck_Km_deg{deg}"]
        if deg <= 2:
            assert np.array_equal(Km, want), deg
        else:
            assert np.abs(Km - want).max() <= 1e-12 * np.abs(want).max()
    grams = fused.resident_grams(c, METHODS, golden["alignf_fit_rows"], normalize_inputs=True)
    full0 = golden["nlck_K0_normalized"]
    fit = golden["alignf_fit_rows"]
    assert np.array_equal(grams[0].to_host(), full0[fit][:, fit])      # normalised SP_k3 sub-block == slice of the reference's
    q = res.QuadForms(grams)
    for deg in (1, 2, 3):
        g = q.grad(u, golden["nlck_alpha"], deg)
        want = golden[f"nlck_grad_deg{deg}"]
        assert np.abs(g - want).max() <= 1e-12 * np.abs(want).max(), deg


def test_fused_equals_array_based_on_every_kernel_family(fused, dna):
    """700 real sequences, every kernel family (dense and pairwise mismatch, WD, WDS, LA): statistics within 1e-12 of the
    array-based path, combinations bit for bit -- the epilogue accumulate does what numpy's sum over stacked kernels does."""
    from kmg import host as kh
    codes, labels = dna
    c = np.ascontiguousarray(codes[1000:1700])
    methods = ["SP_k6", "MM_k5_m1", "MM_k10_m1", "WD_d10", "WDS_d3_s2", "LA_e-11_d-1_b0.5_smith0_eig0", "SP_k1"]
    Ks = [kh.spectrum_gram(c, 6), kh.mismatch_gram(c, 5, 1), kh.mismatch_gram(c, 10, 1), kh.wd_gram(c, 10), kh.wds_gram(c, 3, 2),
          kh.la_gram(c, -11, -1, 0.5, 0), kh.spectrum_gram(c, 1)]
    rng = np.random.default_rng(3)
    idx = np.sort(rng.choice(700, 451, replace=False))
    y = labels[1000:1700][idx].astype(np.float64)
    a0, M0 = kh.alignf_stats(Ks, idx, y)
    a1, M1 = fused.alignf_stats(c, methods, idx, y)
    assert np.abs(a1 - a0).max() <= 1e-12 * np.abs(a0).max()
    assert np.abs(M1 - M0).max() <= 1e-12 * np.abs(M0).max()
    u = rng.random(len(methods))
    assert np.array_equal(fused.combine(c, methods, u), kh.combine(Ks, u))
    # NLCK: normalise every kernel, combine, square, normalise
    Kn = [onp.normalize_K(K.copy()) for K in Ks]
    want = kh.combine(Kn, u, degree=2, normalize=True)
    got = fused.combine(c, methods, u, degree=2, normalize_inputs=True, normalize=True)
    assert np.array_equal(got, want)
    grams = fused.resident_grams(c, methods[:4], idx, normalize_inputs=True)
    for g, K in zip(grams, Kn[:4]):
        assert np.array_equal(g.to_host(), K[idx][:, idx])


def test_classes_from_sequences(fused, golden, dna):
    """ALIGNF.from_sequences / NLCK.from_sequences: the drop-in classes without any host-side kernel."""
    import ALIGNF as A
    import NLCKernels as N
    codes, _ = dna
    c = codes[:96]
    fit, y = golden["alignf_fit_rows"], golden["alignf_y"]
    with contextlib.redirect_stdout(io.StringIO()):
        np.random.seed(11)
        al = A.ALIGNF.from_sequences(c, METHODS, fit, y)
        Km = al.get_K()
    assert np.abs(np.asarray(al.a) - golden["alignf_a"]).max() <= 1e-12 * np.abs(golden["alignf_a"]).max()
    assert np.abs(al.u_star - golden["alignf_u"]).max() <= 1e-6      # same start (seed 11), same QP
    assert np.abs(Km - golden["alignf_Km"]).max() <= 1e-5 * np.abs(golden["alignf_Km"]).max()
    with contextlib.redirect_stdout(io.StringIO()):
        nl = N.NLCK.from_sequences(c, METHODS, fit, y, degree=2)
        g = nl.grad(golden["nlck_u"], golden["nlck_alpha"])
        nl.fit = lambda *a, **k: golden["nlck_u"]
        Kn = nl.get_K()
    assert np.abs(g - golden["nlck_grad_deg2"]).max() <= 1e-12 * np.abs(golden["nlck_grad_deg2"]).max()
    assert np.array_equal(Kn, golden["nlck_Km_deg2"])
