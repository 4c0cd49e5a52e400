"""small_probe.py -- latency of the reference-sized calls (BASELINE configs[0]: n = 3000, k = 6) through the host API."""
import os, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kernel-methods-for-genomics_b200")); sys.path.insert(0, os.path.join(ROOT, "tools"))
from kmg import host as kh
import _inputs as onp
codes = onp.synthetic_codes(9000, 101, seed=3)
for name, fn in (("spectrum k=6 n=3000", lambda: kh.spectrum_gram(codes[:3000], 6)),
                 ("spectrum k=1..7 n=3000", lambda: kh.spectrum_gram(codes[:3000], list(range(1, 8)))),
                 ("wd d=10 n=3000", lambda: kh.wd_gram(codes[:3000], 10)),
                 ("mismatch (6,1) n=3000 [dense]", lambda: kh.mismatch_gram(codes[:3000], 6, 1)),
                 ("mismatch (10,1) n=9000", lambda: kh.mismatch_gram(codes, 10, 1)),
                 ("la affine n=1000", lambda: kh.la_gram(codes[:1000], 11, 1, 0.5))):
    fn()
    ts = []
    for _ in range(9):
        t0 = time.perf_counter(); K = fn(); ts.append(time.perf_counter() - t0); del K
    print(f"{name}: {min(ts) * 1e3:.2f} ms (median {sorted(ts)[2] * 1e3:.2f})  all: " + " ".join(f"{t * 1e3:.1f}" for t in ts), flush=True)
if os.environ.get("KMG_TRACE"):
    kh.spectrum_gram(codes[:3000], 6)
