"""CPU: pins oracle/oracle_np.py against golden vectors produced by the UNMODIFIED reference
(oracle/gen_golden.py ran /root/reference/kernels.py, ALIGNF.py, NLCKernels.py in the build container)."""
import hashlib
import re

import numpy as np
import pytest

import oracle_np as onp


def _sha(K):
    return hashlib.sha256(np.ascontiguousarray(K, np.float64).tobytes()).hexdigest()


def _cases(golden, prefix):
    return sorted(k for k in golden.files if k.startswith(prefix))


def test_spectrum_bit_exact(golden, dna):
    codes, _ = dna
    names = _cases(golden, "sp_k")
    assert len(names) == 7
    for name in names:
        k, n = map(int, re.match(r"sp_k(\d+)_n(\d+)", name).groups())
        K = onp.spectrum_gram(codes[:n], k)
        assert K.dtype == np.float64 and np.array_equal(K, golden[name]), name


def test_wd_bit_exact(golden, dna):
    codes, _ = dna
    names = [k for k in _cases(golden, "wd_d") if "pair" not in k]
    assert len(names) == 5
    for name in names:
        d, n = map(int, re.match(r"wd_d(\d+)_n(\d+)", name).groups())
        K = onp.wd_gram(codes[:n], d)
        assert np.array_equal(K, golden[name]), name
    # the loop value for an identical pair differs from the closed-form diagonal in the last bit
    assert float(golden["wd_d4_pair00"]) == 99.00000000000001
    assert golden["wd_d4_n40"][0, 0] == 99.0


def test_mismatch_bit_exact(golden, dna):
    codes, _ = dna
    names = _cases(golden, "mm_k")
    assert len(names) == 7
    for name in names:
        k, m, n = map(int, re.match(r"mm_k(\d+)_m(\d+)_n(\d+)", name).groups())
        K = onp.mismatch_gram(codes[:n], k, m)
        assert np.array_equal(K, golden[name]), name
        if k <= 5:  # the dense feature map of kernels.py:161-175 gives the same raw integers
            raw = onp.mismatch_gram_raw(codes[:n], k, m)
            assert np.array_equal(raw, onp.mismatch_gram_raw_dense(codes[:n], k, m)), name


def test_mismatch_table_values():
    assert onp.mismatch_table(10, 1)[:4] == [31, 4, 2, 0]
    assert onp.mismatch_table(10, 2)[:6] == [436, 112, 64, 18, 6, 0]
    assert onp.mismatch_table(4, 2) == [67, 40, 28, 18, 6]
    assert onp.mismatch_table(5, 0) == [1, 0, 0, 0, 0, 0]


def test_normalize_center(golden):
    K = golden["norm_in_sp3_n96"].copy()
    out = onp.normalize_K(K)
    assert out is K and np.array_equal(K, golden["norm_out_sp3_n96"])
    K1 = golden["norm_in_sp3_n96"].copy(); K1[0, 0] = 1.0
    assert np.array_equal(onp.normalize_K(K1.copy()), golden["norm_out_early_n96"])
    assert np.array_equal(golden["norm_out_early_n96"], K1)  # early-out left it untouched
    for tag in ("sp3", "wd5"):
        Kin = golden["norm_in_sp3_n96"] if tag == "sp3" else golden["center_in_wd5_n96"]
        ref = golden[f"center_out_{tag}_n96"]
        got = onp.center_K(Kin)
        assert np.abs(got - ref).max() <= 1e-12 * np.abs(Kin).max()


def test_la_reference_is_identically_zero(golden, dna):
    """the reference's LA kernel returns 0.0 for every pair."""
    for k in ("la_affine_pair01", "la_affine_pair01_neg", "la_smith_pair01"):
        assert float(golden[k]) == 0.0
    assert not golden["la_eig0_n4"].any() and not golden["la_smith_eig0_n3"].any()
    codes, _ = dna
    assert np.array_equal(onp.la_reference_compat(codes[:4]), golden["la_eig0_n4"])


def test_la_intended_against_mpmath(dna):
    """The intended recursion is unpinned by the reference; pin it against a 50-digit evaluation."""
    mp = pytest.importorskip("mpmath")
    codes, _ = dna
    x, y = codes[0][:23], codes[1][:19]
    for (e, d, beta) in ((11, 1, 0.5), (-11, -1, 0.5), (11, 1, 0.1)):
        mp.mp.dps = 50
        nx, ny = len(x), len(y)
        Z = lambda: [[mp.mpf(0)] * (ny + 1) for _ in range(nx + 1)]
        M, X, Y, X2, Y2 = Z(), Z(), Z(), Z(), Z()
        b = mp.mpf(beta)
        for i in range(1, nx + 1):
            for j in range(1, ny + 1):
                s = int(onp.S_LA[x[i - 1], y[j - 1]])
                M[i][j] = mp.e ** (b * s) * (1 + X[i - 1][j - 1] + Y[i - 1][j - 1] + M[i - 1][j - 1])
                X[i][j] = mp.e ** (b * d) * M[i - 1][j] + mp.e ** (b * e) * X[i - 1][j]
                Y[i][j] = mp.e ** (b * d) * (M[i][j - 1] + X[i][j - 1]) + mp.e ** (b * e) * Y[i][j - 1]
                X2[i][j] = M[i - 1][j] + X2[i - 1][j]
                Y2[i][j] = M[i][j - 1] + X2[i][j - 1] + Y2[i][j - 1]
        want = float(mp.log(1 + X2[nx][ny] + Y2[nx][ny] + M[nx][ny]) / b)
        got = onp.la_affine_intended(x, y, e, d, beta)
        assert abs(got - want) <= 1e-12 * abs(want), (e, d, beta, got, want)


def test_la_intended_survey_values(dna):
    """SURVEY.md A.5 probed values for (Xtr0[0], Xtr0[1])."""
    codes, _ = dna
    v = onp.la_affine_intended(codes[0], codes[1], -11, -1, 0.5)
    assert abs(v - 397.196) < 1e-3
    v = onp.la_affine_intended(codes[0], codes[1], 11, 1, 0.1)
    assert abs(v - 2463.6) < 0.1


def test_alignf_nlck_algebra(golden):
    Ks = [golden[f"alignf_K{i}"] for i in range(3)]
    idx = golden["alignf_fit_rows"]
    a, M = onp.alignf_stats(Ks, idx, golden["alignf_y"])
    assert np.allclose(a, golden["alignf_a"], rtol=1e-12, atol=0)
    assert np.allclose(M, golden["alignf_M"], rtol=1e-12, atol=0)
    Km = onp.combine(Ks, golden["alignf_u"])
    assert np.array_equal(Km, golden["alignf_Km"])
    # NLCK normalises its kernels in place first (NLCKernels.py:33,43-48)
    Kn = [onp.normalize_K(k.copy()) for k in Ks]
    assert np.array_equal(Kn[0], golden["nlck_K0_normalized"])
    u, alpha = golden["nlck_u"], golden["nlck_alpha"]
    fit = [k[idx][:, idx] for k in Kn]
    for deg in (1, 2, 3):
        g = onp.nlck_grad(fit, u, alpha, deg)
        assert np.allclose(g, golden[f"nlck_grad_deg{deg}"], rtol=1e-12, atol=0)
        Km = onp.normalize_K(onp.combine(Kn, u, deg))
        assert np.abs(Km - golden[f"nlck_Km_deg{deg}"]).max() <= 1e-12


def test_sha256_known_answers(golden, dna):
    """hashes of the reference's outputs on real rows."""
    codes, _ = dna
    kat = dict(zip(golden["kat_names"].tolist(), golden["kat_sha256"].tolist()))
    assert _sha(onp.spectrum_gram(codes[:256], 3)) == kat["sp_k3_Xtr0_256"]
    assert _sha(onp.wd_gram(codes[:256], 5)) == kat["wd_d5_Xtr0_256"]
    assert _sha(onp.wd_gram(codes[:256], 10)) == kat["wd_d10_Xtr0_256"]
    assert _sha(onp.mismatch_gram(codes[:64], 4, 1)) == kat["mm_k4_m1_Xtr0_64"]
    # BASELINE config 1: Xtr0 (2000) then Xte0 (1000); 360 s in the reference, ~2 s here
    c1 = np.concatenate((codes[:2000], codes[6000:7000]))
    K = onp.spectrum_gram(c1, 6)
    assert _sha(K) == kat["sp_k6_Xtr0_Xte0_3000"]
    assert K.sum() == 33300484.0 and np.trace(K) == 314044.0


def test_wds_bit_exact(golden, dna):
    """weighted degree with shifts (first 'next' row); delta_2 = 1/6 makes the order matter."""
    codes, _ = dna
    assert np.array_equal(onp.wds_gram(codes[:12], 3, 2), golden["wds_d3_s2_n12"])
    assert np.array_equal(onp.wds_gram(codes[:10], 5, 1), golden["wds_d5_s1_n10"])
