"""GPU parity: (k,m)-mismatch, weighted-degree and local-alignment kernels against the oracle and the
reference's golden vectors.  Mismatch raw integers and WD fp64 values are bit-exact; local alignment
(intended recursion, parity unpinned by the reference) is within 1e-12 relative."""
import hashlib
import re

import numpy as np
import pytest

import oracle_c as oc
import oracle_np as onp

pytestmark = pytest.mark.gpu

LA_RTOL = 1e-12  # BASELINE.json north_star: "within a stated relative tolerance (1e-12 in fp64)"


@pytest.fixture(scope="module")
def kd():
    torch = pytest.importorskip("torch")
    assert torch.cuda.is_available(), "these tests need a GPU"
    from kmg import device
    return device


@pytest.fixture(scope="module")
def kh():
    from kmg import host
    return host


def _sha(K):
    return hashlib.sha256(np.ascontiguousarray(K, np.float64).tobytes()).hexdigest()


# ------------------------------------------------------------------------------------------ WD
def test_wd_golden(kh, golden, dna):
    codes, _ = dna
    for name in [k for k in golden.files if k.startswith("wd_d") and "pair" not in k]:
        d, n = map(int, re.match(r"wd_d(\d+)_n(\d+)", name).groups())
        assert np.array_equal(kh.wd_gram(codes[:n], d), golden[name]), name
    kat = dict(zip(golden["kat_names"].tolist(), golden["kat_sha256"].tolist()))
    assert _sha(kh.wd_gram(codes[:256], 10)) == kat["wd_d10_Xtr0_256"]
    assert _sha(kh.wd_gram(codes[:256], 5)) == kat["wd_d5_Xtr0_256"]
    # identical pair off the diagonal gets the loop value, not the closed form (kernels.py:96 vs :74-81)
    assert kh.wd_gram(codes[:1], 4, cols=codes[:1])[0, 0] == float(golden["wd_d4_pair00"]) == 99.00000000000001


def test_wd_blocks_and_edges(kd, kh, dna):
    codes, _ = dna
    c = codes[:777]
    want = oc.wd_block(c, c, 10)
    planes = kd.pack(c, 0)
    full = kd.wd_block(planes, planes, 101, 10, symmetric=True).cpu().numpy()
    assert np.array_equal(full, want)
    nosym = kd.wd_block(planes, planes, 101, 10).cpu().numpy()
    assert np.array_equal(nosym, want)
    blk = kd.wd_block(planes[100:333], planes, 101, 10, row_index0=100).cpu().numpy()
    assert np.array_equal(blk, want[100:333])
    # duplicates inside the data set (11 in Xtr0): off-diagonal identical pairs take the loop value
    dup = np.concatenate((c[:5], c[:5]))
    assert np.array_equal(kh.wd_gram(dup, 7), oc.wd_block(dup, dup, 7))
    for L, d in ((2, 1), (33, 40), (64, 3), (128, 20)):
        cc = onp.synthetic_codes(70, L, seed=L)
        assert np.array_equal(kh.wd_gram(cc, d), oc.wd_block(cc, cc, d)), (L, d)
    # long runs (near-duplicates) around the 96-bit boundary where the kernel drops to three words per vector
    rng = np.random.default_rng(7)
    for L, d in ((95, 10), (96, 10), (97, 12), (100, 30), (101, 10), (101, 60), (101, 101), (128, 127), (128, 33)):
        base = onp.synthetic_codes(6, L, seed=100 + L)
        cc = np.repeat(base, 8, axis=0)
        hit = rng.random(cc.shape) < 0.03
        cc = np.where(hit, (cc + rng.integers(1, 4, cc.shape)) % 4, cc).astype(np.uint8)
        assert np.array_equal(kh.wd_gram(cc, d), oc.wd_block(cc, cc, d)), (L, d)
        assert np.array_equal(kh.wd_gram(cc[:7], d, cols=cc[5:]), oc.wd_block(cc[:7], cc[5:], d, 0, 1 << 40)), (L, d)
    assert kh.wd_gram(codes[:0], 3).shape == (0, 0)
    assert np.array_equal(kh.wd_gram(codes[:9], 5, cols=codes[20:51]), oc.wd_block(codes[:9], codes[20:51], 5, 0, 1 << 40))


# ------------------------------------------------------------------------------------ mismatch
@pytest.mark.parametrize("algo", [1, 2])
def test_mismatch_golden(kh, golden, dna, algo):
    codes, _ = dna
    for name in [k for k in golden.files if k.startswith("mm_k")]:
        k, m, n = map(int, re.match(r"mm_k(\d+)_m(\d+)_n(\d+)", name).groups())
        K = kh.mismatch_gram(codes[:n], k, m, algo=algo)
        assert np.array_equal(K, golden[name]), (name, algo)
    kat = dict(zip(golden["kat_names"].tolist(), golden["kat_sha256"].tolist()))
    assert _sha(kh.mismatch_gram(codes[:64], 4, 1, algo=algo)) == kat["mm_k4_m1_Xtr0_64"]


def test_mismatch_pairwise_raw_vs_oracle(kd, kh, dna):
    codes, _ = dna
    c = codes[:150]
    planes = kd.pack(c, 0)
    for (k, m) in ((10, 1), (10, 2), (1, 0), (2, 1), (7, 3), (16, 1), (17, 2), (20, 0), (13, 3), (101, 1)):
        want = oc.mismatch_raw_block(c, c, k, m)
        got = kd.mismatch_block(planes, planes, 101, k, m, out_dtype=0).cpu().numpy()
        assert np.array_equal(got.astype(np.int64), want), (k, m)
        got = kd.mismatch_block(planes, planes, 101, k, m, out_dtype=1, symmetric=True).cpu().numpy()
        assert np.array_equal(got, want.astype(np.float64)), (k, m, "sym")
    # normalised, block rows, cross
    sd = kd.mismatch_diag_sqrt(planes, 101, 10, 1)
    raw = oc.mismatch_raw_block(c, c, 10, 1).astype(np.float64)
    assert np.array_equal(sd.cpu().numpy(), np.sqrt(np.diag(raw)))
    want = onp.normalize_K(raw.copy())
    assert np.array_equal(kh.mismatch_gram(c, 10, 1), want)
    blk = kd.mismatch_block(planes[33:77], planes, 101, 10, 1, row_index0=33, sd_rows=sd[33:77], sd_cols=sd).cpu().numpy()
    assert np.array_equal(blk, want[33:77])
    assert np.array_equal(kh.mismatch_gram(c[:10], 10, 1, cols=c[50:90], normalize=False), raw[:10, 50:90])
    # other lengths
    for L, k, m in ((12, 5, 1), (64, 9, 2), (128, 11, 1), (128, 8, 1), (1, 1, 0), (33, 1, 0), (97, 2, 1), (100, 4, 1), (100, 5, 2),
                    (128, 128, 1), (96, 1, 1), (127, 32, 3)):
        cc = onp.synthetic_codes(40, L, seed=L + k)
        got = kh.mismatch_gram(cc, k, m, normalize=False, algo=1)
        assert np.array_equal(got, oc.mismatch_raw_block(cc, cc, k, m).astype(np.float64)), (L, k, m)


def test_mismatch_dense_vs_pairwise(kd, kh, dna):
    """dense feature map + tcgen05 GEMM (the reference's own structure) == pairwise identity, k <= 8."""
    codes, _ = dna
    c = codes[2000:2300]
    for (k, m) in ((5, 1), (6, 1), (6, 2), (8, 1), (4, 3), (3, 0)):
        a = kh.mismatch_gram(c, k, m, normalize=False, algo=1)
        b = kh.mismatch_gram(c, k, m, normalize=False, algo=2)
        assert np.array_equal(a, b), (k, m)
        a = kh.mismatch_gram(c, k, m, algo=1)
        b = kh.mismatch_gram(c, k, m, algo=2)
        assert np.array_equal(a, b), (k, m, "normalised")
    phi = kh.mismatch_phi(c[:20], 4, 2)
    assert np.array_equal(phi.astype(np.int64), onp.mismatch_phi(c[:20], 4, 2))
    # MM(k,0) normalised == normalised spectrum
    sp = onp.normalize_K(onp.spectrum_gram(c, 5))
    assert np.array_equal(kh.mismatch_gram(c, 5, 0), sp)


def test_mismatch_config2_subset(kh, dna):
    """BASELINE config 2 shape: (k,m) = (10,1) over the challenge sequences (a 512-row sample here; the C oracle
    needs ~20 s for it).  Normalised Gram bit-exact against the oracle."""
    codes, _ = dna
    idx = np.arange(0, 9000, 9000 // 512)[:512]
    c = codes[idx]
    want = onp.normalize_K(oc.mismatch_raw_block(c, c, 10, 1).astype(np.float64))
    got = kh.mismatch_gram(c, 10, 1)
    assert np.array_equal(got, want)
    assert np.all(np.diag(got) == 1.0) and np.array_equal(got, got.T)


# --------------------------------------------------------------------------- local alignment
def test_la_affine_vs_oracle(kh, dna):
    codes, _ = dna
    c = codes[:24]
    for (e, d, beta) in ((11, 1, 0.5), (-11, -1, 0.5), (11, 1, 0.1), (-5.5, -0.7, 1.3)):
        want = oc.la_block(c, c, e, d, beta, 0)
        got = kh.la_gram(c, e, d, beta, 0)
        assert np.array_equal(got, got.T)
        rel = np.abs(got - want).max() / np.abs(want).max()
        assert rel <= LA_RTOL, (e, d, beta, rel)
        assert np.all(np.abs(got - want) <= LA_RTOL * np.abs(want)), (e, d, beta)
    v = kh.la_gram(codes[:1], -11, -1, 0.5, 0, cols=codes[1:2])[0, 0]
    assert abs(v - 397.196) < 1e-3  # SURVEY.md A.5 probe


def test_la_smith_and_shapes(kh, dna):
    codes, _ = dna
    c = codes[100:116]
    for (e, d, beta) in ((11, 1, 0.5), (-11, -1, 0.5)):
        want = oc.la_block(c, c, e, d, beta, 1)
        got = kh.la_gram(c, e, d, beta, 1)
        assert np.all(np.abs(got - want) <= LA_RTOL * np.abs(want) + 1e-300), (e, d, beta)
    for L in (1, 5, 32, 33, 100, 128):
        cc = onp.synthetic_codes(6, L, seed=L)
        want = oc.la_block(cc, cc, -11, -1, 0.5, 0)
        got = kh.la_gram(cc, -11, -1, 0.5, 0)
        assert np.all(np.abs(got - want) <= LA_RTOL * np.abs(want) + 1e-300), L
    # cross-Gram: x is always the row sequence
    want = oc.la_block(c[:3], c[5:9], -11, -1, 0.5, 0, 0, 1 << 40)
    got = kh.la_gram(c[:3], -11, -1, 0.5, 0, cols=c[5:9])
    assert np.all(np.abs(got - want) <= LA_RTOL * np.abs(want))


# ------------------------------------------------------------------ weighted degree with shifts
def test_wds_golden_and_oracle(kd, kh, golden, dna):
    codes, _ = dna
    assert np.array_equal(kh.wds_gram(codes[:12], 3, 2), golden["wds_d3_s2_n12"])
    assert np.array_equal(kh.wds_gram(codes[:10], 5, 1), golden["wds_d5_s1_n10"])
    c = codes[500:540]
    for d, S in ((4, 0), (6, 3), (10, 2), (3, 7), (12, 5)):
        want = onp.wds_gram(c, d, S)
        assert np.array_equal(kh.wds_gram(c, d, S), want), (d, S)
    planes = kd.pack(c, 0)
    want = onp.wds_gram(c, 5, 2)
    assert np.array_equal(kd.wds_block(planes[8:24], planes, 101, 5, 2, row_index0=8).cpu().numpy(), want[8:24])
    assert np.array_equal(kd.wds_block(planes, planes, 101, 5, 2, symmetric=True).cpu().numpy(), want)
    cc = onp.synthetic_codes(20, 37, seed=4)
    assert np.array_equal(kh.wds_gram(cc, 6, 4), onp.wds_gram(cc, 6, 4))


# ------------------------------------------------------------------ local alignment: third, literal route
LA_PARAM_SETS = ((11, 1, 0.5), (-11, -1, 0.5), (11, 1, 0.1), (-5.5, -0.7, 1.3))


@pytest.mark.parametrize("smith", [0, 1])
def test_la_256_pairs_literal_logsumexp_rtol_1e12(kh, dna, smith):
    """Parity-unpinned kernel (the reference returns 0.0): 256 sampled pairs of real 101-bp sequences, all
    four parameter sets, against a LITERAL log-sum-exp evaluation of the intended recursion (tests/la_literal.py: chains
    of numpy.logaddexp, no code shared with the kernel or the oracle).  Tolerance: 1e-12 relative, per entry."""
    from la_literal import la_pairs_logspace
    codes, _ = dna
    rng = np.random.default_rng(7)
    ri, ci = rng.choice(9000, 16, replace=False), rng.choice(9000, 16, replace=False)
    rows, cols = codes[ri], codes[ci]
    xs, ys = np.repeat(rows, 16, axis=0), np.tile(cols, (16, 1))       # pair p = (row p // 16, col p % 16)
    for (e, d, beta) in LA_PARAM_SETS:
        want = la_pairs_logspace(xs, ys, e, d, beta, smith).reshape(16, 16)
        got = kh.la_gram(rows, e, d, beta, smith, cols=cols)           # cross-Gram: x is always the row sequence
        assert np.all(np.isfinite(got)) and np.all(got > 0)
        rel = np.abs(got - want) / np.abs(want)
        assert rel.max() <= 1e-12, (smith, e, d, beta, rel.max())


def test_config5_la_block_at_scale(kd):
    """BASELINE config 5 shape: local alignment (affine, reference defaults e=11 d=1 beta=0.5 taken literally) on the
    n = 20 000 synthetic problem (seed 5): a 512 x 20 000 block-row.  Size-independent properties -- the block that holds
    the diagonal is symmetric about it and equals the symmetric-mode build, K[A,B] == K[B,A]^T across blocks (x is always
    the sequence of smaller index, kernels.py:289-291), every value finite and positive -- and sampled tiles against the C
    oracle at 1e-12 relative."""
    import torch
    n, r0, R = 20_000, 7_000, 512
    c = onp.synthetic_codes(n, 101, seed=5)
    planes = kd.pack(c, 0)
    blk = kd.la_block(planes[r0:r0 + R], planes, 101, 11, 1, 0.5, 0, row_index0=r0)
    assert bool(torch.isfinite(blk).all()) and float(blk.min()) > 0.0
    dg = blk[:, r0:r0 + R]
    assert torch.equal(dg, dg.t())
    sym = kd.la_block(planes[r0:r0 + R], planes[r0:r0 + R], 101, 11, 1, 0.5, 0, row_index0=r0, col_index0=r0, symmetric=True)
    assert torch.equal(sym, dg)
    other = kd.la_block(planes[1000:1128], planes[r0:r0 + R], 101, 11, 1, 0.5, 0, row_index0=1000, col_index0=r0)
    assert torch.equal(other.t(), blk[:, 1000:1128])
    other = kd.la_block(planes[19_000:19_128], planes[r0:r0 + R], 101, 11, 1, 0.5, 0, row_index0=19_000, col_index0=r0)
    assert torch.equal(other.t(), blk[:, 19_000:19_128])
    rng = np.random.default_rng(11)
    for _ in range(6):
        a, b = int(rng.integers(0, R - 8)), int(rng.integers(0, n - 32))
        want = oc.la_block(c[r0 + a:r0 + a + 8], c[b:b + 32], 11, 1, 0.5, 0, r0 + a, b)
        got = blk[a:a + 8, b:b + 32].cpu().numpy()
        assert np.all(np.abs(got - want) <= LA_RTOL * np.abs(want)), (a, b)


def test_config4_wd_full_blockrow_at_scale(kd):
    """BASELINE config 4 at the benchmark's unit: WD d=10, one 12 500 x 100 000 block-row of the synthetic 100k problem
    (seed 4).  Properties: closed-form diagonal exactly where the indices coincide, 0 <= K <= diag, K[A,B] == K[B,A]^T
    across blocks, the block holding the diagonal equals the symmetric-mode build, the entry histogram is a set of
    multiples of 1/110 within rounding (beta_k c_k with d=10), and 10 random tiles == the C oracle bit for bit."""
    import torch
    n, r0, R = 100_000, 37_500, 12_500
    c = onp.synthetic_codes(n, 101, seed=4)
    planes = kd.pack(c, 0)
    blk = kd.wd_block(planes[r0:r0 + R], planes, 101, 10, row_index0=r0)
    diag = blk[torch.arange(R), torch.arange(r0, r0 + R)]
    dv = 100 + (1 - 10) / 3
    assert torch.all(diag == dv)
    assert float(blk.min()) >= 0.0 and float(blk.max()) <= dv
    dg = blk[:4096, r0:r0 + 4096]
    assert torch.equal(dg, dg.t())
    assert torch.equal(kd.wd_block(planes[r0:r0 + 4096], planes[r0:r0 + 4096], 101, 10, row_index0=r0, col_index0=r0, symmetric=True), dg)
    for b0 in (0, 90_000):
        other = kd.wd_block(planes[b0:b0 + 2048], planes[r0:r0 + R], 101, 10, row_index0=b0, col_index0=r0)
        assert torch.equal(other.t(), blk[:, b0:b0 + 2048])
    # 110 K is within a few ulps of an integer (beta_k = 2(11-k)/110): a cheap whole-block sanity check of the weights
    frac = (blk[:2048] * 110.0)
    assert float((frac - frac.round()).abs().max()) < 1e-9
    rng = np.random.default_rng(13)
    for _ in range(10):
        a, b = int(rng.integers(0, R - 16)), int(rng.integers(0, n - 64))
        want = oc.wd_block(c[r0 + a:r0 + a + 16], c[b:b + 64], 10, r0 + a, b)
        assert np.array_equal(blk[a:a + 16, b:b + 64].cpu().numpy(), want), (a, b)
