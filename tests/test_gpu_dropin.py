"""GPU parity through the reference-facing boundary: the drop-in `kernels` module (same names, arguments
and return values as the reference's kernels.py) and the ALIGNF / NLCK Gram-side algebra, against the
golden vectors recorded from the unmodified reference."""
import numpy as np
import pandas as pd
import pytest

import oracle_np as onp

pytestmark = pytest.mark.gpu

NORMWISE = 1e-12  # centre / combine: max|err| <= 1e-12 * max|K|  


@pytest.fixture(scope="module")
def km():
    torch = pytest.importorskip("torch")
    assert torch.cuda.is_available(), "these tests need a GPU"
    import kernels
    return kernels


@pytest.fixture(scope="module")
def X0(dna):
    codes, _ = dna
    seqs = onp.decode(codes[:2000])
    return pd.DataFrame({"Id": np.arange(2000), "seq": seqs})


def test_select_method_dsl(km, golden, X0, capsys):
    assert np.array_equal(km.select_method(X0.iloc[:16], "SP_k3"), golden["sel_SP_k3_n16"])
    assert np.array_equal(km.select_method(X0.iloc[:16], "WD_d5"), golden["sel_WD_d5_n16"])
    assert "['WD', 'd5']" in capsys.readouterr().out  # the reference prints the split method (kernels.py:484)
    assert np.array_equal(km.select_method(X0.iloc[:16], "MM_k3_m1"), golden["sel_MM_k3_m1_n16"])
    with pytest.raises(NotImplementedError):
        km.select_method(X0.iloc[:4], "XX_k3")
    K = km.select_method(X0.iloc[:6], "LA_e-11_d-1_b0.5_smith0_eig1")
    assert K.shape == (6, 6) and np.array_equal(K, K.T) and (K > 0).all()


def test_builders_match_reference(km, golden, X0):
    for name in golden.files:
        if name.startswith("sp_k"):
            k, n = int(name.split("_")[1][1:]), int(name.split("_")[2][1:])
            K = km.get_spectrum_K(X0.iloc[:n], k)
        elif name.startswith("wd_d") and "pair" not in name:
            d, n = int(name.split("_")[1][1:]), int(name.split("_")[2][1:])
            K = km.get_WD_K(X0.iloc[:n], d)
        elif name.startswith("mm_k"):
            k, m, n = int(name.split("_")[1][1:]), int(name.split("_")[2][1:]), int(name.split("_")[3][1:])
            K = km.get_mismatch_K(X0.iloc[:n], k, m)
        else:
            continue
        assert isinstance(K, np.ndarray) and K.dtype == np.float64 and K.flags.c_contiguous
        assert np.array_equal(K, golden[name]), name


def test_duplicate_index_is_read_positionally(km, X0):
    """utils.py:151 concatenates frames, so the index has duplicates; the reference iterates X.loc[:, 'seq']."""
    X = pd.concat((X0.iloc[:5], X0.iloc[:5]), axis=0)
    K = km.get_spectrum_K(X, 3)
    assert K.shape == (10, 10) and np.array_equal(K[:5, :5], K[5:, 5:]) and np.array_equal(K[:5, :5], K[:5, 5:])


def test_helpers(km, golden, X0):
    x, y = X0.seq[0], X0.seq[1]
    assert km.beta(10, 3) == onp.wd_beta(10, 3)
    assert km.get_WD_d(x, x, 4, 101) == float(golden["wd_d4_pair00"])
    assert km.get_WD_d(x, y, 4, 101) == golden["wd_d4_n40"][0, 1]
    from itertools import product
    betas = ["".join(c) for c in product("ACGT", repeat=3)]
    phi = km.get_phi_u(x, 3, betas)
    assert phi.dtype == np.float64 and np.array_equal(phi, onp.spectrum_phi(onp.encode([x]), 3)[0])
    assert phi @ phi == golden["sp_k3_n128"][0, 0]
    fb = np.array([km.format("".join(c)) for c in product("ACGT", repeat=3)])
    phikm = km.get_phi_km(km.format(x), 3, 1, fb)
    assert np.array_equal(phikm, onp.mismatch_phi(onp.encode([x]), 3, 1)[0])
    assert np.array_equal(km.format("ACGT"), [1, 2, 3, 4]) and km.letter_to_num("GATTACA") == "3144121"
    assert km.S[2, 3] == 2 and km.S[3, 2] == -2


def test_normalize_center(km, golden, capsys):
    K = golden["norm_in_sp3_n96"].copy()
    out = km.normalize_K(K)
    assert out is K and np.array_equal(K, golden["norm_out_sp3_n96"])
    K1 = golden["norm_in_sp3_n96"].copy(); K1[0, 0] = 1.0
    capsys.readouterr()
    out = km.normalize_K(K1)
    assert "Kernel already normalized" in capsys.readouterr().out
    assert out is K1 and np.array_equal(K1, golden["norm_out_early_n96"])
    for tag, Kin in (("sp3", golden["norm_in_sp3_n96"]), ("wd5", golden["center_in_wd5_n96"])):
        Kc = km.center_K(Kin)
        assert Kc is not Kin and np.abs(Kc - golden[f"center_out_{tag}_n96"]).max() <= NORMWISE * np.abs(Kin).max()
    # larger, against the numpy restatement of the reference's multi_dot
    rng = np.random.Generator(np.random.PCG64(5))
    A = rng.standard_normal((700, 40)); Kb = A @ A.T
    assert np.abs(km.center_K(Kb) - onp.center_K(Kb)).max() <= NORMWISE * np.abs(Kb).max()
    Kn = Kb.copy(); km.normalize_K(Kn)
    assert np.array_equal(Kn, onp.normalize_K(Kb.copy()))


def test_la_reference_compat(km, golden, X0, monkeypatch):
    """The reference's LA kernel is identically zero; the compat switch reproduces it."""
    monkeypatch.setattr(km, "LA_REFERENCE_COMPAT", True)
    assert km.affine_align(X0.seq[0], X0.seq[1], 11, 1, 0.5) == float(golden["la_affine_pair01"]) == 0.0
    assert km.Smith_Waterman(X0.seq[0], X0.seq[1]) == float(golden["la_smith_pair01"]) == 0.0
    assert np.array_equal(km.get_LA_K(X0.iloc[:4], 11, 1, 0.5, 0, 0), golden["la_eig0_n4"])
    with pytest.raises(Exception):
        km.get_LA_K(X0.iloc[:8])  # eig=1: ARPACK fails on the zero matrix, as in the reference
    monkeypatch.setattr(km, "LA_REFERENCE_COMPAT", False)
    v = km.affine_align(X0.seq[0], X0.seq[1], -11, -1, 0.5)
    assert abs(v - 397.196) < 1e-3


def test_alignf_nlck_algebra(golden):
    from kmg import host
    Ks = [golden[f"alignf_K{i}"] for i in range(3)]
    idx = golden["alignf_fit_rows"]
    a, M = host.alignf_stats(Ks, idx, golden["alignf_y"])
    scale = np.sqrt(np.outer(np.diag(golden["alignf_M"]), np.diag(golden["alignf_M"])))
    assert np.abs(M - golden["alignf_M"]).max() <= 1e-12 * scale.max()
    assert np.all(np.abs(M - golden["alignf_M"]) <= 1e-10 * scale)
    assert np.allclose(a, golden["alignf_a"], rtol=1e-10, atol=1e-12 * np.abs(golden["alignf_a"]).max())
    Km = host.combine(Ks, golden["alignf_u"])
    assert np.array_equal(Km, golden["alignf_Km"])                      # ALIGNF.get_K: bit-exact
    Kn = []
    for k in Ks:
        k = k.copy(); host.normalize_inplace(k); Kn.append(k)
    assert np.array_equal(Kn[0], golden["nlck_K0_normalized"])
    u, alpha = golden["nlck_u"], golden["nlck_alpha"]
    fit = [np.ascontiguousarray(k[idx][:, idx]) for k in Kn]
    for deg in (1, 2, 3):
        g = host.nlck_grad(fit, u, alpha, deg)
        ref = golden[f"nlck_grad_deg{deg}"]
        assert np.all(np.abs(g - ref) <= 1e-12 * np.abs(ref).max()), deg
        Km = host.combine(Kn, u, deg, normalize=True)
        assert np.abs(Km - golden[f"nlck_Km_deg{deg}"]).max() <= 1e-12, deg
        if deg <= 2:
            assert np.array_equal(Km, golden[f"nlck_Km_deg{deg}"]), deg     # deg 1, 2: bit-exact (x*x == np.square)


def test_resident_grams(golden):
    """kmg/resident.py: Grams uploaded once and re-used -- same bits as the per-call host entry points."""
    from kmg import host, resident
    Ks = [golden[f"alignf_K{i}"] for i in range(3)]
    idx = golden["alignf_fit_rows"]
    dev = [resident.DeviceGram.from_host(k) for k in Ks]
    assert np.array_equal(dev[1].to_host(), Ks[1])
    assert dev[0].normalize_() is False
    Kn = [k.copy() for k in Ks]
    host.normalize_inplace(Kn[0])
    assert np.array_equal(dev[0].to_host(), Kn[0]) and np.array_equal(Kn[0], golden["nlck_K0_normalized"])
    assert dev[0].normalize_() is True                                   # K[0,0] == 1 early-out (kernels.py:403)
    for g in dev[1:]:
        g.normalize_()
    for k in Kn[1:]:
        host.normalize_inplace(k)
    fit_dev = [g.gather(idx) for g in dev]
    fit = [np.ascontiguousarray(k[idx][:, idx]) for k in Kn]
    for a, b in zip(fit_dev, fit):
        assert np.array_equal(a.to_host(), b)
    u, alpha = golden["nlck_u"], golden["nlck_alpha"]
    q = resident.QuadForms(fit_dev)
    for deg in (1, 2, 3):
        for _ in range(2):                                               # second call re-uses every buffer
            g = q.grad(u, alpha, deg)
            assert np.array_equal(g, host.nlck_grad(fit, u, alpha, deg)), deg
        ref = golden[f"nlck_grad_deg{deg}"]
        assert np.all(np.abs(g - ref) <= 1e-12 * np.abs(ref).max()), deg
        assert np.array_equal(resident.combine(fit_dev, u, deg).to_host(), host.combine(fit, u, deg)), deg
    for g in dev + fit_dev:
        g.free()
        g.free()                                                         # idempotent


def test_wds_dropin(km, golden, X0):
    assert np.array_equal(km.select_method(X0.iloc[:12], "WDS_d3_s2"), golden["wds_d3_s2_n12"])
    assert np.array_equal(km.get_WDShifts_K(X0.iloc[:10], 5, 1), golden["wds_d5_s1_n10"])
    assert km.get_WDShifts_d(X0.seq[0], X0.seq[1], 3, 2, 101) == golden["wds_d3_s2_n12"][0, 1]
    assert km.delta(2) == 1 / 2 / 3


def test_alignf_nlck_classes(golden, X0, dna, capsys):
    """The drop-in ALIGNF / NLCK classes (same constructor and methods as the reference's) against the values the
    reference's own classes produced on the same kernels (oracle/gen_golden.py)."""
    import ALIGNF as A
    import NLCKernels as N
    _, labels = dna
    n_all = 96
    Xa = X0.iloc[:n_all].copy()
    ya = pd.DataFrame({"Id": np.arange(n_all), "Bound": labels[:n_all].astype(float)})
    rows = golden["alignf_fit_rows"]
    Ks = [golden[f"alignf_K{i}"].copy() for i in range(3)]
    np.random.seed(11)
    al = A.ALIGNF(Xa.iloc[rows], ya.iloc[rows], np.arange(n_all), Ks)
    scale = np.sqrt(np.outer(np.diag(golden["alignf_M"]), np.diag(golden["alignf_M"])))
    assert np.all(np.abs(al.M - golden["alignf_M"]) <= 1e-10 * scale)
    assert np.allclose(al.a, golden["alignf_a"], rtol=1e-10, atol=1e-12 * np.abs(golden["alignf_a"]).max())
    assert np.allclose(al.u_star, golden["alignf_u"], atol=1e-8)
    Km = al.get_K()
    assert np.abs(Km - golden["alignf_Km"]).max() <= 1e-8 * np.abs(golden["alignf_Km"]).max()
    u, alpha = golden["nlck_u"], golden["nlck_alpha"]
    for deg in (1, 2, 3):
        nl = N.NLCK(Xa.iloc[rows], ya.iloc[rows], np.arange(n_all), [golden[f"alignf_K{i}"].copy() for i in range(3)], degree=deg)
        ref = golden[f"nlck_grad_deg{deg}"]
        assert np.all(np.abs(nl.grad(u, alpha) - ref) <= 1e-12 * np.abs(ref).max())
        nl.fit = lambda *a, **k: u   # skip the cvxopt QP exactly as gen_golden.py did
        assert np.abs(nl.get_K() - golden[f"nlck_Km_deg{deg}"]).max() <= 1e-12
    capsys.readouterr()
