"""GPU parity: packing, spectrum feature map and the tcgen05 int8 Gram GEMM against the oracle and the
reference's golden vectors.  Integer work: bit-exact."""
import hashlib
import re

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


def _planes_ref(codes):
    n, L = codes.shape
    pl = np.zeros((n, 8), np.uint32)
    for p in range(L):
        pl[:, p // 32] |= ((codes[:, p] & 1).astype(np.uint32) << np.uint32(p % 32))
        pl[:, 4 + p // 32] |= (((codes[:, p] >> 1) & 1).astype(np.uint32) << np.uint32(p % 32))
    return pl


def test_pack_codes_and_ascii(kd, dna):
    codes, _ = dna
    c = codes[:300]
    want = _planes_ref(c)
    got = kd.pack(c, 0).cpu().numpy().view(np.uint32)
    assert np.array_equal(got, want)
    ascii_ = np.frombuffer("".join(onp.decode(c)).encode(), np.uint8).reshape(c.shape)
    got = kd.pack(ascii_, 1).cpu().numpy().view(np.uint32)
    assert np.array_equal(got, want)
    bad = ascii_.copy(); bad[7, 50] = ord("N")
    with pytest.raises(ValueError):
        kd.pack(bad, 1)
    # ragged lengths: L = 1, 33, 128
    for L in (1, 33, 128):
        cc = onp.synthetic_codes(17, L, seed=L)
        assert np.array_equal(kd.pack(cc, 0).cpu().numpy().view(np.uint32), _planes_ref(cc))


def test_spectrum_phi_matches_oracle(kd, dna):
    codes, _ = dna
    c = codes[1000:1100]
    planes = kd.pack(c, 0)
    for ks in ([1], [3], [6], [1, 2, 3, 4, 5, 6, 7], [7], [8]):
        phi = kd.spectrum_phi(planes, 101, ks).cpu().numpy()
        want = np.concatenate([onp.spectrum_phi(c, k) for k in ks], axis=1)
        D = want.shape[1]
        assert phi.shape[1] % 128 == 0 and phi.shape[1] >= D
        assert np.array_equal(phi[:, :D].astype(np.int64), want), ks
        assert not phi[:, D:].any()


@pytest.mark.parametrize("m_sub", [1, 2, 3])
def test_gram_tcgen05_vs_oracle_small(kd, dna, m_sub):
    """Known answers: real rows, every k, both tile shapes, s32 and f64 outputs, ragged n."""
    codes, _ = dna
    for ks, n in (([3], 128), ([1], 64), ([6], 40), ([6], 257), ([1, 2, 3, 4, 5, 6, 7], 300), ([7], 130), ([2, 5], 513)):
        c = codes[:n]
        planes = kd.pack(c, 0)
        phi = kd.spectrum_phi(planes, 101, ks)
        want = oc.spectrum_block(c, c, ks)
        got = kd.gram_i8(phi, phi, out_dtype=1, m_sub=m_sub).cpu().numpy()
        assert got.dtype == np.float64 and np.array_equal(got, want), (ks, n, "f64")
        got = kd.gram_i8(phi, phi, out_dtype=0, m_sub=m_sub).cpu().numpy()
        assert np.array_equal(got.astype(np.float64), want), (ks, n, "s32")
        got = kd.gram_i8(phi, phi, out_dtype=1, symmetric=True, m_sub=m_sub).cpu().numpy()
        assert np.array_equal(got, want), (ks, n, "symmetric")


def test_gram_golden_vectors(kd, golden, dna):
    codes, _ = dna
    for name in [k for k in golden.files if k.startswith("sp_k")]:
        k, n = map(int, re.match(r"sp_k(\d+)_n(\d+)", name).groups())
        planes = kd.pack(codes[:n], 0)
        phi = kd.spectrum_phi(planes, 101, [k])
        got = kd.gram_i8(phi, phi, symmetric=True).cpu().numpy()
        assert np.array_equal(got, golden[name]), name


@pytest.mark.parametrize("m_sub", [1, 2, 3])
def test_gram_tcgen05_vs_simt_medium(kd, m_sub):
    """n = 3000 synthetic, D = 21 844: the tcgen05 kernel against an independent dp4a evaluation on device,
    plus block-row / cross-block launches against slices of the same matrix."""
    import torch
    c = onp.synthetic_codes(3000, 101, seed=3)
    planes = kd.pack(c, 0)
    phi = kd.spectrum_phi(planes, 101, list(range(1, 8)))
    ref = kd.gram_i8_simt(phi, phi)
    got = kd.gram_i8(phi, phi, out_dtype=0, m_sub=m_sub)
    assert torch.equal(got, ref)
    sym = kd.gram_i8(phi, phi, out_dtype=0, symmetric=True, m_sub=m_sub)
    assert torch.equal(sym, ref)
    blk = kd.gram_i8(phi[700:1500], phi, row_index0=700, out_dtype=1, m_sub=m_sub)
    assert torch.equal(blk, ref[700:1500].double())
    cross = kd.gram_i8(phi[100:357], phi[2000:2999], out_dtype=0, m_sub=m_sub)
    assert torch.equal(cross, ref[100:357, 2000:2999])
    # CPU oracle on a sampled tile
    want = oc.spectrum_block(c[1234:1250], c[77:141], list(range(1, 8)))
    assert np.array_equal(ref[1234:1250, 77:141].cpu().numpy().astype(np.float64), want)


def test_gram_normalised_epilogue(kd, dna):
    codes, _ = dna
    c = codes[:200]
    planes = kd.pack(c, 0)
    phi = kd.spectrum_phi(planes, 101, [4])
    sd = kd.phi_diag_sqrt(phi)
    raw = onp.spectrum_gram(c, 4)
    assert np.array_equal(sd.cpu().numpy(), np.sqrt(np.diag(raw)))
    want = onp.normalize_K(raw.copy())
    got = kd.gram_i8(phi, phi, symmetric=True, sd_rows=sd, sd_cols=sd).cpu().numpy()
    assert np.array_equal(got, want)
    got = kd.gram_i8(phi[50:120], phi, row_index0=50, sd_rows=sd[50:120], sd_cols=sd).cpu().numpy()
    assert np.array_equal(got, want[50:120])


def test_host_api_spectrum(kh, golden, dna):
    codes, _ = dna
    # BASELINE config 1: Xtr0 (2000) then Xte0 (1000), k=6 -> SHA-256 recorded from the reference
    c1 = np.concatenate((codes[:2000], codes[6000:7000]))
    K = kh.spectrum_gram(c1, 6)
    kat = dict(zip(golden["kat_names"].tolist(), golden["kat_sha256"].tolist()))
    assert hashlib.sha256(np.ascontiguousarray(K).tobytes()).hexdigest() == kat["sp_k6_Xtr0_Xte0_3000"]
    # cross-Gram = the [train, test] sub-block
    Kx = kh.spectrum_gram(codes[:2000], 6, cols=codes[6000:7000])
    assert np.array_equal(Kx, K[:2000, 2000:])
    # ASCII path, sum over k, empty input, k > L, k = 9 (pairwise route)
    seqs = onp.decode(codes[:50])
    assert np.array_equal(kh.spectrum_gram(seqs, [1, 2, 3]), onp.spectrum_gram(codes[:50], [1, 2, 3]))
    assert kh.spectrum_gram(codes[:0], 3).shape == (0, 0)
    assert np.array_equal(kh.spectrum_gram(codes[:20], 9), onp.spectrum_gram(codes[:20], 9))
    assert not kh.spectrum_gram(codes[:5, :4], 5).any()
    with pytest.raises(ValueError):
        kh.spectrum_gram(["ACGT", "ACGN"], 2)
    with pytest.raises(ValueError):
        kh.spectrum_gram(["ACGT", "ACG"], 2)


def test_host_link_transports(kh, dna, monkeypatch):
    """The unnormalised counts cross PCIe as u16 when every entry of the block fits, as s32 otherwise, or as fp64
    (KMG_D2H_F64): same bits in the caller's array either way.  Homopolymers at k=1..7 reach 67 256 > 65 535."""
    codes, _ = dna
    ks = [1, 2, 3, 4, 5, 6, 7]
    c = codes[:700].copy()
    want = onp.spectrum_gram(c, ks)
    assert want.max() <= 65535
    assert np.array_equal(kh.spectrum_gram(c, ks), want)                 # u16 link (large enough for a recycled block)
    assert np.array_equal(kh.spectrum_gram(c[:300], ks, cols=c), want[:300])
    c[3] = 0
    c[11] = 0
    want = onp.spectrum_gram(c, ks)
    assert want[3, 11] == 67256
    assert np.array_equal(kh.spectrum_gram(c, ks), want)                 # overflow detected on the device: s32 link
    monkeypatch.setenv("KMG_D2H_S32", "1")
    assert np.array_equal(kh.spectrum_gram(c, ks), want)
    monkeypatch.setenv("KMG_D2H_F64", "1")
    assert np.array_equal(kh.spectrum_gram(c, ks), want)
    monkeypatch.setenv("KMG_DEVICE_BUDGET_BYTES", str(2 * 8 * 256 * 700))  # streamed block-rows, u16 and s32 blocks mixed
    monkeypatch.delenv("KMG_D2H_F64")
    monkeypatch.delenv("KMG_D2H_S32")
    assert np.array_equal(kh.spectrum_gram(c, ks), want)
    assert np.array_equal(kh.spectrum_gram(c[:300], ks, cols=c), want[:300])


def test_host_chunked_cross_gram(kh, kd):
    """A cross-Gram of >= 1024 rows and >= 64e6 entries is built in row chunks that drain to the host while the GPU
    builds the next ones; one chunk holds an entry > 65535 and crosses as s32, the others as u16.  Same bits as the
    device-resident GEMM."""
    ks = [1, 2, 3, 4, 5, 6, 7]
    c = onp.synthetic_codes(60000, 101, seed=11)
    rows = c[:1100].copy()
    rows[700] = 0
    c[5] = 0
    got = kh.spectrum_gram(rows, ks, cols=c)
    phi_c = kd.spectrum_phi(kd.pack(c, 0), 101, ks)
    phi_r = kd.spectrum_phi(kd.pack(rows, 0), 101, ks)
    want = kd.gram_i8(phi_r, phi_c, out_dtype=1).cpu().numpy()
    assert got[700, 5] == 67256 and got.shape == (1100, 60000)
    assert np.array_equal(got, want)
    assert np.array_equal(got[:16, :64], onp.spectrum_gram(np.concatenate((rows[:16], c[:64])), ks)[:16, 16:])


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_sharded_symmetric_gram_single_device(kd, world):
    """kmg_gram_i8_sharded_dev with every part's buffer on this one GPU: each part's launch computes about half of its
    block-row and mirror-stores the rest into the other parts' buffers; together they must equal the plain Gram."""
    import torch
    from kmg import dist as kdist
    n = 3000 if world < 8 else 2300
    c = onp.synthetic_codes(n, 101, seed=3)
    planes = kd.pack(c, 0)
    for ks, dt in (([3], 1), (list(range(1, 8)), 1), ([1, 2, 3, 4, 5, 6, 7], 0)):
        phi = kd.spectrum_phi(planes, 101, ks)
        ref = kd.gram_i8(phi, phi, out_dtype=dt)
        bounds = kdist.sym_bounds(n, world)
        bufs = [torch.full((bounds[p + 1] - bounds[p], n), -7, dtype=ref.dtype, device="cuda") for p in range(world)]
        ptrs = [b.data_ptr() for b in bufs]
        computed = [kd.gram_i8_sharded(phi, bounds, p, ptrs, n, out_dtype=dt) for p in range(world)]
        torch.cuda.synchronize()
        assert torch.equal(torch.cat(bufs), ref), (world, ks, dt)
        if world > 1:
            assert sum(computed) < 0.62 * n * n
            assert all(computed[p] < 0.75 * (bounds[p + 1] - bounds[p]) * n for p in range(world)), computed
        # staged variant: one launch per peer block into local staging + one pitched copy per block
        for b in bufs:
            b.fill_(-7)
        stages = [torch.empty(max(kd.sharded_stage_bytes(bounds, p, dt), 8), dtype=torch.uint8, device="cuda") for p in range(world)]
        computed2 = [kd.gram_i8_sharded(phi, bounds, p, ptrs, n, out_dtype=dt, stage=stages[p].data_ptr()) for p in range(world)]
        torch.cuda.synchronize()
        assert torch.equal(torch.cat(bufs), ref), (world, ks, dt, "staged")
        assert computed2 == computed
        # direct variant: one launch per peer block, mirror stores issued by the TMA engine into the owner's buffer
        for b in bufs:
            b.fill_(-7)
        computed3 = [kd.gram_i8_sharded(phi, bounds, p, ptrs, n, out_dtype=dt, exchange="direct") for p in range(world)]
        torch.cuda.synchronize()
        assert torch.equal(torch.cat(bufs), ref), (world, ks, dt, "direct")
        assert computed3 == computed
    # fused cosine normalisation through the mirror stores (kernels.py:398-415)
    phi = kd.spectrum_phi(planes, 101, [4])
    sd = kd.phi_diag_sqrt(phi)
    ref = kd.gram_i8(phi, phi, symmetric=True, sd_rows=sd, sd_cols=sd)
    bounds = kdist.sym_bounds(n, world)
    bufs = [torch.zeros((bounds[p + 1] - bounds[p], n), dtype=torch.float64, device="cuda") for p in range(world)]
    for mode in ("single", "staged", "direct"):
        for b in bufs:
            b.zero_()
        for p in range(world):
            st = torch.empty(max(kd.sharded_stage_bytes(bounds, p, 1), 8), dtype=torch.uint8, device="cuda") if mode == "staged" else None
            kd.gram_i8_sharded(phi, bounds, p, [b.data_ptr() for b in bufs], n, sd=sd, stage=None if st is None else st.data_ptr(), exchange=mode)
            torch.cuda.synchronize()
        assert torch.equal(torch.cat(bufs), ref), mode


def test_full_size_properties(kd):
    """BASELINE-sized feature width (k=1..7) at n = 20 000: size-independent properties on device --
    symmetry, diagonal = sum of squares of Phi, row sums = Phi (Phi^T 1), and agreement of the two tile shapes."""
    import torch
    n = 20000
    c = onp.synthetic_codes(n, 101, seed=3)
    planes = kd.pack(c, 0)
    phi = kd.spectrum_phi(planes, 101, list(range(1, 8)))
    K = kd.gram_i8(phi, phi, out_dtype=0, symmetric=True, m_sub=2)
    assert torch.equal(K, K.t())
    sd = kd.phi_diag_sqrt(phi)
    assert torch.equal(torch.diagonal(K).double(), (sd * sd).round())
    colsum = phi.to(torch.float64).sum(0)                       # Phi^T 1 (exact in fp64)
    rows = torch.arange(0, n, 97, device=phi.device)
    want = phi[rows].to(torch.float64) @ colsum                 # exact: all integers < 2^53
    assert torch.equal(K[rows].to(torch.float64).sum(1), want)
    K1 = kd.gram_i8(phi[:4096], phi, out_dtype=0, m_sub=1)
    assert torch.equal(K1, K[:4096])


def test_mma_peak_microbenchmark(kd):
    """bench.py's roofline denominator: the back-to-back tcgen05.mma.kind::i8 rate must be a plausible B200 figure and
    at least what the full GEMM (loads + epilogue included) reaches."""
    import torch
    peak = kd.mma_peak_i8(iters=50000, repeats=2)
    assert 3000.0 < peak < 6000.0, peak
    c = onp.synthetic_codes(8192, 101, seed=3)
    phi = kd.spectrum_phi(kd.pack(c, 0), 101, list(range(1, 8)))
    out = torch.empty((8192, 8192), dtype=torch.int32, device="cuda")
    best = 0.0
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); kd.gram_i8(phi, phi, out_dtype=0, out=out); e1.record(); torch.cuda.synchronize()
        best = max(best, 2.0 * 8192 * 8192 * phi.shape[1] / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    assert best < peak * 1.02, (best, peak)
