"""GPU: the K_fit algebra of the reference's closed-form solvers on the device (csrc/solve.cu; drop-ins KRR.py / KLR.py)
against golden vectors produced by the UNMODIFIED reference KRR.py / KLR.py (oracle/gen_golden_solvers.py ->
tests/golden/ref_solvers.npz).  Tolerances: the reference inverts the matrix (LAPACK LU) and multiplies, the device
factors it (Cholesky) and substitutes; for a system with condition number kappa the two agree to ~kappa * 1e-16, so the
tests use 1e-9 relative for the coefficients (kappa <= ~1e5 here) and a residual bound for the solve itself."""
import contextlib
import io
import os

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gs():
    torch = pytest.importorskip("torch")
    assert torch.cuda.is_available(), "these tests need a GPU"
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_solvers.npz"))


def _frames(gs, n_all):
    fit, y = gs["fit_rows"], gs["y"]
    return pd.DataFrame({"Id": fit}), pd.DataFrame({"Id": fit, "Bound": y}), pd.DataFrame({"Id": np.arange(n_all)})


def test_spd_solve_residual():
    """(S K S + c I) x = b at the reference's fit size (1501), with and without the diagonal scaling: relative residual
    <= 1e-12, and a non-positive-definite matrix is reported, not silently factored."""
    from kmg import host as kh
    from kmg import resident as res
    rng = np.random.default_rng(5)
    n = 1501
    F = rng.standard_normal((n, 300))
    K = F @ F.T / 300.0
    b = rng.standard_normal(n)
    for s in (None, np.sqrt(rng.random(n) * 0.25 + 1e-3)):
        c = 0.1 * n
        x = kh.spd_solve(K, b, c, s=s)
        A = K if s is None else s[:, None] * K * s[None, :]
        r = A @ x + c * x - b
        assert np.linalg.norm(r) <= 1e-12 * np.linalg.norm(b)
        idx = np.sort(rng.choice(n, 700, replace=False))
        x2 = kh.spd_solve(K, b[idx], 3.0, idx=idx)
        assert np.linalg.norm(K[idx][:, idx] @ x2 + 3.0 * x2 - b[idx]) <= 1e-12 * np.linalg.norm(b[idx])
    g = res.DeviceGram.from_host(K)
    assert np.linalg.norm(g.matvec(b) - K @ b) <= 1e-13 * np.linalg.norm(K @ b)
    with pytest.raises(ValueError):
        kh.spd_solve(-K, b, 1e-3)


def test_krr_dropin_matches_the_reference(gs, golden):
    import KRR as ours
    for name, K in (("nlck2", golden["nlck_Km_deg2"]), ("wd5", golden["alignf_K1"])):
        X, Y, Xall = _frames(gs, K.shape[0])
        for lbda in (0.1, 1e-3):
            m = ours.KRR(K.copy(), np.arange(K.shape[0]), lbda=lbda)
            m.fit(X, Y)
            want = gs[f"krr_{name}_l{lbda}_a"]
            assert np.array_equal(m.idx_sv, gs[f"krr_{name}_l{lbda}_sv"])
            assert np.abs(m.a - want).max() <= 1e-9 * np.abs(want).max(), (name, lbda)
            assert abs(m.b - float(gs[f"krr_{name}_l{lbda}_b"])) <= 1e-9 * max(1.0, abs(float(gs[f"krr_{name}_l{lbda}_b"])))
            assert np.array_equal(m.predict(Xall), gs[f"krr_{name}_l{lbda}_pred"])
            assert m.score(m.predict(X), Y) >= 0.5


def test_klr_dropin_matches_the_reference(gs, golden):
    import KLR as ours
    for name, K in (("nlck2", golden["nlck_Km_deg2"]), ("wd5", golden["alignf_K1"])):
        X, Y, Xall = _frames(gs, K.shape[0])
        fit, y = gs["fit_rows"], gs["y"]
        m = ours.KLR(K.copy(), np.arange(K.shape[0]), lbda=0.1)
        with np.errstate(over="ignore"):
            m.n = fit.size
            W, z = m.IRLS(np.ascontiguousarray(K[fit][:, fit]), y, np.linspace(-0.01, 0.01, fit.size))
            got = m.WKRR(np.ascontiguousarray(K[fit][:, fit]), W, z)
            want = gs[f"klr_{name}_wkrr"]
            assert np.abs(got - want).max() <= 1e-9 * np.abs(want).max(), name
            if name == "wd5":
                continue  # the raw weighted-degree Gram (entries ~97) saturates the sigmoids: the full fit is degenerate in the reference too
            m.fit(X, Y)
        want = gs[f"klr_{name}_a"]
        assert np.array_equal(m.idx_sv, gs[f"klr_{name}_sv"])
        assert np.abs(m.a - want).max() <= 1e-7 * np.abs(want).max(), name   # up to 50 IRLS iterations feed back
        assert np.array_equal(m.predict(Xall), gs[f"klr_{name}_pred"])


def test_resident_pipeline_from_sequences(gs, golden, dna):
    """Sequences -> combined normalised Gram on the device -> sub-block gather -> ridge solve, nothing but vectors on
    the host: the KRR coefficients equal the drop-in's on the host-built Gram."""
    from kmg import fused
    from kmg import resident as res
    import KRR as ours
    codes, _ = dna
    methods = ["SP_k3", "WD_d5", "MM_k3_m1"]
    Km = fused.combine(codes[:96], methods, golden["nlck_u"], degree=2, normalize_inputs=True, normalize=True)
    fit, y = gs["fit_rows"], gs["y"]
    X, Y, _ = _frames(gs, 96)
    m = ours.KRR(Km, np.arange(96), lbda=0.1)
    m.fit(X, Y)
    g = res.DeviceGram.from_host(Km)
    sub = g.gather(fit)                       # reformat_data's K[idx][:, idx] (utils.py:301-304) without leaving the device
    a = sub.spd_solve(y, 0.1 * fit.size)
    keep = np.abs(a) > 1e-5
    assert np.abs(a[keep] - m.a).max() <= 1e-12 * np.abs(m.a).max()
