"""GPU: size-independent properties of the Gram builders at sizes the CPU oracle cannot check entry by entry
(BASELINE configs 2-4 shapes), plus randomised small cases.  All integer / fp64-ordered kernels: exact equality."""
import numpy as np
import pytest

import oracle_c as oc
import oracle_np as onp

pytestmark = pytest.mark.gpu


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


def test_permutation_equivariance_and_symmetry(kh):
    """K(Px) = P K(x) P^T for every kernel; symmetric builds equal their own transpose."""
    rng = np.random.Generator(np.random.PCG64(17))
    c = onp.synthetic_codes(257, 101, seed=17)
    perm = rng.permutation(c.shape[0])
    builders = {
        "sp6": lambda x: kh.spectrum_gram(x, 6),
        "sp1-7": lambda x: kh.spectrum_gram(x, list(range(1, 8))),
        "mm(10,1)": lambda x: kh.mismatch_gram(x, 10, 1),
        "mm(5,2) dense": lambda x: kh.mismatch_gram(x, 5, 2, algo=2),
        "wd10": lambda x: kh.wd_gram(x, 10),
        "wds(5,2)": lambda x: kh.wds_gram(x, 5, 2),
    }
    for name, f in builders.items():
        K = f(c)
        assert np.array_equal(K, K.T), name
        assert np.array_equal(f(c[perm]), K[np.ix_(perm, perm)]), name
    K = kh.la_gram(c[:64], -11, -1, 0.5, 0)
    Kp = kh.la_gram(c[perm[:64]], -11, -1, 0.5, 0)
    # LA evaluates (x = lower index, y = higher index): a permutation may swap the roles, and the recursion is not
    # symmetric in (x, y), so only the diagonal and the symmetry of each build are invariant
    assert np.array_equal(K, K.T) and np.array_equal(Kp, Kp.T)


def test_randomised_small_cases_against_oracle(kh):
    rng = np.random.Generator(np.random.PCG64(23))
    for trial in range(12):
        L = int(rng.integers(12, 129))
        n = int(rng.integers(1, 70))
        c = onp.synthetic_codes(n, L, seed=100 + trial)
        k = int(rng.integers(1, min(L, 12) + 1))
        m = int(rng.integers(0, 4))
        d = int(rng.integers(1, min(L - 1, 14) + 1))
        assert np.array_equal(kh.spectrum_gram(c, min(k, 8)), oc.spectrum_block(c, c, [min(k, 8)])), (trial, "sp", L, n, k)
        raw = oc.mismatch_raw_block(c, c, k, m).astype(np.float64)
        assert np.array_equal(kh.mismatch_gram(c, k, m, normalize=False, algo=1), raw), (trial, "mm", L, n, k, m)
        if L - k + 1 <= 127 and k <= 8:
            assert np.array_equal(kh.mismatch_gram(c, k, m, normalize=False, algo=2), raw), (trial, "mm dense", L, n, k, m)
        assert np.array_equal(kh.wd_gram(c, d), oc.wd_block(c, c, d)), (trial, "wd", L, n, d)


def test_config2_mismatch_full_9000_properties(kd, dna):
    """BASELINE config 2: (k,m) = (10,1) over all 9000 challenge sequences, normalised.  Properties: exact symmetry,
    unit diagonal, duplicate sequences (8508 unique of 9000) give entries == 1 up to the rounding of d*d, row blocks equal the
    symmetric build, and a sampled tile equals the oracle bit for bit."""
    import torch
    codes, _ = dna
    planes = kd.pack(codes, 0)
    sd = kd.mismatch_diag_sqrt(planes, 101, 10, 1)
    K = kd.mismatch_block(planes, planes, 101, 10, 1, symmetric=True, sd_rows=sd, sd_cols=sd)
    assert torch.equal(K, K.t())
    assert torch.all(torch.diagonal(K) == 1.0)
    assert float(K.max()) <= 1.0 + 1e-15 and float(K.min()) >= 0.0
    blk = kd.mismatch_block(planes[4096:4352], planes, 101, 10, 1, row_index0=4096, sd_rows=sd[4096:4352], sd_cols=sd)
    assert torch.equal(blk, K[4096:4352])
    rows, cols = slice(7000, 7024), slice(100, 148)
    want = oc.mismatch_raw_block(codes[rows], codes[cols], 10, 1).astype(np.float64)
    sdh = sd.cpu().numpy()
    want = want / (sdh[rows][:, None] * sdh[cols][None, :])
    assert np.array_equal(K[rows, cols].cpu().numpy(), want)


def test_config2_mismatch_full_9000_sha256(kh, dna):
    """BASELINE config 2 in full through the host C-ABI: get_mismatch_K(X, 10, 1) on all 9000 challenge sequences against
    the SHA-256 known answer computed once by the plain-C oracle (oracle/gen_config2_kat.py -> tests/golden/
    config2_mm10_kat.json; the reference cannot run this configuration).  Raw integers and the normalised fp64 matrix,
    bit for bit."""
    import hashlib
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config2_mm10_kat.json")
    if not os.path.exists(path):
        pytest.skip("known answer not generated (python oracle/gen_config2_kat.py)")
    kat = json.load(open(path))
    codes, _ = dna
    K = kh.mismatch_gram(codes, 10, 1)
    assert hashlib.sha256(np.ascontiguousarray(K).tobytes()).hexdigest() == kat["sha256"]
    assert float(np.trace(K)) == kat["trace"] and K[0, 1] == kat["K_0_1"] and K[4503, 8999] == kat["K_4503_8999"]
    raw = kh.mismatch_gram(codes, 10, 1, normalize=False)
    assert hashlib.sha256(np.ascontiguousarray(raw.astype(np.int64)).tobytes()).hexdigest() == kat["raw_sha256"]
    assert int(raw[0, 0]) == kat["raw_0_0"] and int(np.trace(raw)) == kat["raw_trace"]


def test_config4_wd_blockrow_properties(kd):
    """BASELINE config 4 shape: WD d=10, one 2048 x 100000 block-row of the synthetic 100k problem.  Properties:
    diagonal closed form exactly where indices coincide, bounds 0 <= K <= L-1+(1-d)/3 + eps, sampled tiles == oracle."""
    import torch
    n, r0, R = 100_000, 50_000, 2048
    c = onp.synthetic_codes(n, 101, seed=4)
    planes = kd.pack(c, 0)
    blk = kd.wd_block(planes[r0:r0 + R], planes, 101, 10, row_index0=r0)
    diag = blk[torch.arange(R), torch.arange(r0, r0 + R)]
    assert torch.all(diag == (100 + (1 - 10) / 3))
    assert float(blk.min()) >= 0.0
    for (a, b) in ((0, 0), (1000, 99_000), (2047, r0)):
        want = oc.wd_block(c[r0 + a:r0 + a + 1], c[b:b + 96], 10, r0 + a, b)
        assert np.array_equal(blk[a:a + 1, b:b + 96].cpu().numpy(), want)
    sym = kd.wd_block(planes[:4096], planes[:4096], 101, 10, symmetric=True)
    full = kd.wd_block(planes[:4096], planes[:4096], 101, 10)
    assert torch.equal(sym, full)


def test_spectrum_linearity_over_k(kd):
    """sum_k K_k built as ONE concatenated-feature GEMM == the sum of the single-k Grams (exact integers)."""
    import torch
    c = onp.synthetic_codes(3000, 101, seed=8)
    planes = kd.pack(c, 0)
    total = torch.zeros((3000, 3000), dtype=torch.int64, device="cuda")
    for k in range(1, 8):
        phi = kd.spectrum_phi(planes, 101, [k])
        total += kd.gram_i8(phi, phi, out_dtype=0).to(torch.int64)
    phi = kd.spectrum_phi(planes, 101, list(range(1, 8)))
    assert torch.equal(kd.gram_i8(phi, phi, out_dtype=0, symmetric=True).to(torch.int64), total)


def test_streamed_block_rows_equal_single_launch():
    """The host API streams row blocks through two device buffers when n x n fp64 exceeds the device budget
    (n > ~88 000 on a 180 GB B200).  KMG_DEVICE_BUDGET_BYTES forces that path at n = 3000 in a fresh process."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {os.path.join(root, 'kernel-methods-for-genomics_b200')!r}); sys.path.insert(0, {os.path.join(root, 'oracle')!r})\n"
        "from kmg import host as kh\nimport oracle_np as onp, oracle_c as oc\n"
        "c = onp.synthetic_codes(3000, 101, seed=12)\n"
        "K = kh.spectrum_gram(c, [1, 2, 3, 4, 5, 6, 7])\n"
        "assert np.array_equal(K, K.T)\n"
        "assert np.array_equal(K[1000:1016, 2900:3000], oc.spectrum_block(c[1000:1016], c[2900:3000], list(range(1, 8))))\n"
        "W = kh.wd_gram(c, 10)\n"
        "assert np.array_equal(W[2990:3000], oc.wd_block(c[2990:3000], c, 10, 2990, 0))\n"
        "M = kh.mismatch_gram(c[:1500], 10, 1)\n"
        "assert np.array_equal(M, M.T) and np.all(np.diag(M) == 1.0)\n"
        "np.save(sys.argv[1], np.array([K.sum(), W.sum(), M.sum()]))\n")
    outs = []
    for budget in (None, "48000000"):
        env = dict(os.environ)
        if budget:
            env["KMG_DEVICE_BUDGET_BYTES"] = budget
        path = os.path.join("/tmp", f"kmg_stream_{budget}.npy")
        subprocess.run([sys.executable, "-c", code, path], check=True, env=env, timeout=600)
        outs.append(np.load(path))
    assert np.array_equal(outs[0], outs[1])


def test_concurrent_host_calls_from_threads(kh):
    """ctypes drops the GIL inside libkmg: host entry points called from several Python threads at once share the
    pinned ring, the overflow flags, the buffer caches and the wave counters, and must return the sequential results."""
    import threading
    ks = list(range(1, 8))
    data = [onp.synthetic_codes(1500 + 200 * i, 101, seed=40 + i) for i in range(4)]
    jobs = [lambda d=d: kh.spectrum_gram(d, ks) for d in data[:2]]
    jobs += [lambda d=data[2]: kh.wd_gram(d[:800], 10), lambda d=data[3]: kh.mismatch_gram(d[:600], 10, 1),
             lambda d=data[0]: kh.spectrum_gram(d[:1200], ks, cols=onp.synthetic_codes(56000, 101, seed=50))]
    want = [j() for j in jobs]
    got = [None] * len(jobs)

    def run(i):
        got[i] = jobs[i]()
    for _ in range(2):
        threads = [threading.Thread(target=run, args=(i,)) for i in range(len(jobs))]
        for t in threads:
            t.start()
        for t in threads:
            t.join(timeout=300)
        for i in range(len(jobs)):
            assert got[i] is not None and np.array_equal(got[i], want[i]), i
