"""CPU: the plain-C oracle (oracle/kmg_oracle.c) against the golden vectors and the numpy oracle."""
import re

import numpy as np

import oracle_c as oc
import oracle_np as onp


def test_c_spectrum(golden, dna):
    codes, _ = dna
    for name in [k for k in golden.files if k.startswith("sp_k")]:
        k, n = map(int, re.match(r"sp_k(\d+)_n(\d+)", name).groups())
        assert np.array_equal(oc.spectrum_block(codes[:n], codes[:n], [k]), golden[name]), name
    c = codes[100:164]
    assert np.array_equal(oc.spectrum_block(c, c, [1, 2, 3, 4, 5, 6, 7]), onp.spectrum_gram(c, range(1, 8)))


def test_c_wd(golden, dna):
    codes, _ = dna
    for name in [k for k in golden.files if k.startswith("wd_d") and "pair" not in k]:
        d, n = map(int, re.match(r"wd_d(\d+)_n(\d+)", name).groups())
        assert np.array_equal(oc.wd_block(codes[:n], codes[:n], d), golden[name]), name
    # rectangular block with global indices: diagonal closed form only where indices coincide
    blk = oc.wd_block(codes[8:24], codes[:40], 4, row_index0=8, col_index0=0)
    assert np.array_equal(blk, golden["wd_d4_n40"][8:24, :40])


def test_c_mismatch(golden, dna):
    codes, _ = dna
    for name in [k for k in golden.files if k.startswith("mm_k")]:
        k, m, n = map(int, re.match(r"mm_k(\d+)_m(\d+)_n(\d+)", name).groups())
        raw = oc.mismatch_raw_block(codes[:n], codes[:n], k, m)
        assert np.array_equal(onp.normalize_K(raw.astype(np.float64)), golden[name]), name
    c = codes[:24]
    assert np.array_equal(oc.mismatch_raw_block(c, c, 10, 1), onp.mismatch_gram_raw(c, 10, 1))


def test_c_la_matches_numpy(dna):
    codes, _ = dna
    c = codes[:3, :37]
    for smith in (0, 1):
        for (e, d, b) in ((11, 1, 0.5), (-11, -1, 0.5)):
            want = onp.la_gram_intended(c, e, d, b, smith)
            got = oc.la_block(c, c, e, d, b, smith)
            assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
