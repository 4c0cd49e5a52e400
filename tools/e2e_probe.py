"""e2e_probe.py -- where the time goes in one host-API call (run with KMG_TRACE=1)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kernel-methods-for-genomics_b200")); sys.path.insert(0, os.path.join(ROOT, "tools"))
from kmg import host as kh
import _inputs as onp
n = 200000
codes = onp.synthetic_codes(n, 101, seed=3)
rows = np.ascontiguousarray(codes[:2048])
kh.spectrum_gram(rows[:256], list(range(1, 8)), cols=codes)
for i in range(3):
    t0 = time.perf_counter(); K = kh.spectrum_gram(rows, list(range(1, 8)), cols=codes); dt = time.perf_counter() - t0
    print(f"call {i}: {dt*1e3:.1f} ms  {2048*n/dt/1e9:.3f} Gentries/s  {2048*n*8/dt/1e9:.2f} GB/s", flush=True)
    del K
t0 = time.perf_counter(); K = kh.spectrum_gram(codes[:20000], 6); dt = time.perf_counter() - t0
print(f"sym k=6 n=20000: {dt*1e3:.1f} ms {4e8/dt/1e9:.3f} Gentries/s", flush=True)
