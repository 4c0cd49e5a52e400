"""big_probe.py -- one large symmetric Gram through the host API (n = 60 000, k = 1..7: 28.8 GB of fp64 on the host),
checked by properties: symmetry, diagonal = sum of squares of Phi, sampled rows against a direct device launch."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kernel-methods-for-genomics_b200")); sys.path.insert(0, os.path.join(ROOT, "tools"))
from kmg import host as kh, device as kd
import _inputs as onp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
ks = list(range(1, 8))
codes = onp.synthetic_codes(n, 101, seed=3)
for it in range(2):
    t0 = time.perf_counter(); K = kh.spectrum_gram(codes, ks); dt = time.perf_counter() - t0
    print(f"call {it}: n={n} {dt:.3f} s  {n * n / dt / 1e9:.2f} Gentries/s", flush=True)
    if it == 0:
        del K
phi = kd.spectrum_phi(kd.pack(codes, 0), 101, ks)
rows = np.array([0, 1, 255, 256, 4097, n // 2, n - 257, n - 1])
direct = kd.gram_i8(phi[torch.from_numpy(rows).cuda()].contiguous(), phi, out_dtype=1).cpu().numpy()
assert np.array_equal(K[rows], direct), "sampled rows differ"
assert np.array_equal(K[:, rows].T, direct), "sampled columns differ (mirror)"
diag = (phi.to(torch.float64) ** 2).sum(1).cpu().numpy() if n <= 70000 else None
if diag is not None:
    assert np.array_equal(np.diag(K), diag), "diagonal"
blk = slice(n // 3, n // 3 + 3000)
assert np.array_equal(K[blk, :][:, blk], K[blk, :][:, blk].T), "block symmetry"
print("properties ok; max entry", K.max(), flush=True)
