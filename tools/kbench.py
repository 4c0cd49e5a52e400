"""kbench.py -- per-kernel device timings (CUDA events on the launching stream) for tuning.
Not the contract benchmark (that is bench.py); prints one line per configuration."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "kernel-methods-for-genomics_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)

from kmg import device as kd  # noqa: E402
import _inputs as onp  # noqa: E402


def timeit(fn, warmup=2, iters=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="gemm,phi,wd,mm,la")
    ap.add_argument("--n", type=int, default=16384)
    ap.add_argument("--cols", type=int, default=0)
    ap.add_argument("--quick", action="store_true", help="gemm: only k=1..7, CTA-pair kernel, f64")
    args = ap.parse_args()
    what = args.what.split(",")
    n = args.n
    codes = onp.synthetic_codes(max(n, args.cols), 101, seed=3)
    planes = kd.pack(codes, 0)
    cols = args.cols or n
    if "phi" in what:
        for ks in ([6], list(range(1, 8))):
            W = kd.phi_width(ks)
            out = torch.empty((n, W), dtype=torch.int8, device="cuda")
            med, best = timeit(lambda: kd.spectrum_phi(planes[:n], 101, ks, out=out))
            print(f"phi ks={ks} n={n}: {med:.3f} ms  write {n * W / med / 1e6:.1f} GB/s")
    if "gemm" in what:
        for ks in ([6], list(range(1, 8))):
            if args.quick and len(ks) == 1:
                continue
            phi = kd.spectrum_phi(planes, 101, ks)
            W = phi.shape[1]
            for m_sub in (1, 2, 3):
                for dt in (0, 1):
                    if args.quick and (m_sub, dt) != (3, 1):
                        continue
                    out = torch.empty((n, cols), dtype=torch.float64 if dt else torch.int32, device="cuda")
                    med, best = timeit(lambda: kd.gram_i8(phi[:n], phi[:cols], out_dtype=dt, m_sub=m_sub, out=out))
                    ops = 2.0 * n * cols * W
                    print(f"gemm ks={ks[0]}..{ks[-1]} W={W} {n}x{cols} m_sub={m_sub} out={'f64' if dt else 's32'}: {med:.3f} ms "
                          f"(best {best:.3f})  {ops / med / 1e9:.1f} TOPS  {n * cols / med / 1e6:.2f} Gentries/s  "
                          f"write {n * cols * (8 if dt else 4) / med / 1e6:.0f} GB/s")
            if n == cols:
                out = torch.empty((n, n), dtype=torch.float64, device="cuda")
                med, best = timeit(lambda: kd.gram_i8(phi[:n], phi[:n], out_dtype=1, symmetric=True, out=out))
                print(f"gemm ks={ks[0]}..{ks[-1]} W={W} {n}x{n} symmetric f64: {med:.3f} ms  {n * n / med / 1e6:.2f} Gentries/s delivered")
    if "ew" in what:
        # stored-Gram passes (HBM bound): bytes = reads + writes of the pass
        m = min(n, 16384)
        Ks = [torch.rand((m, m), dtype=torch.float64, device="cuda") + 1.0 for _ in range(3)]
        u = [0.2, 0.3, 0.5]
        for name, fn, nbytes in (
                ("center (row sums + col sums + apply)", lambda: kd.center(Ks[0]), 8.0 * m * m * 4),
                ("combine p=3 degree 2", lambda: kd.combine(Ks, u, 2), 8.0 * m * m * 4),
                ("combine p=1 degree 1", lambda: kd.combine(Ks[:1], u[:1], 1), 8.0 * m * m * 2),
                ("normalize (in place)", lambda: kd.normalize_(Ks[1]), 8.0 * m * m * 1.5),
                ("row_sums", lambda: kd.row_sums(Ks[2]), 8.0 * m * m),
                ("weighted_dot <A,B>", lambda: kd.weighted_dot(Ks[0], Ks[2]), 8.0 * m * m * 2)):
            med, best = timeit(fn)
            print(f"ew {name} {m}x{m}: {med:.3f} ms  {nbytes / med / 1e6:.0f} GB/s")
    if "cublas" in what:
        # library reference for context only (cuBLASLt int8 IMMA through torch._int_mm, s32 output): not a product path
        for W in (4096, 21888):
            a = torch.randint(0, 4, (n, W), dtype=torch.int8, device="cuda")
            b = torch.randint(0, 4, (W, cols), dtype=torch.int8, device="cuda")
            try:
                med, best = timeit(lambda: torch._int_mm(a, b))
                print(f"cublasLt int8 (torch._int_mm) {n}x{cols}x{W} s32: {med:.3f} ms  {2.0 * n * cols * W / med / 1e9:.1f} TOPS")
            except Exception as exc:  # noqa: BLE001
                print("cublasLt int8 unavailable:", exc)
            del a, b
    if "wd" in what:
        nn = min(n, 32768)
        out = torch.empty((nn, nn), dtype=torch.float64, device="cuda")
        for d in (5, 10):
            med, best = timeit(lambda: kd.wd_block(planes[:nn], planes[:nn], 101, d, out=out))
            print(f"wd d={d} {nn}x{nn} full: {med:.3f} ms  {nn * nn / med / 1e6:.2f} Gentries/s  write {nn * nn * 8 / med / 1e6:.0f} GB/s")
            med, best = timeit(lambda: kd.wd_block(planes[:nn], planes[:nn], 101, d, symmetric=True, out=out))
            print(f"wd d={d} {nn}x{nn} symmetric: {med:.3f} ms  {nn * nn / med / 1e6:.2f} Gentries/s delivered")
    if "mm" in what:
        nn = min(n, 4096)
        out = torch.empty((nn, nn), dtype=torch.float64, device="cuda")
        for (k, m) in ((10, 1), (10, 2), (6, 1), (20, 1)):
            med, best = timeit(lambda: kd.mismatch_block(planes[:nn], planes[:nn], 101, k, m, out=out), warmup=1, iters=3)
            W = 101 - k + 1
            print(f"mm pairwise ({k},{m}) {nn}x{nn} full: {med:.3f} ms  {nn * nn / med / 1e3:.1f} Mentries/s  {nn * nn * W * W / med / 1e9:.2f} T window-pairs/s")
    if "la" in what:
        nn = min(n, 1024)
        out = torch.empty((nn, nn), dtype=torch.float64, device="cuda")
        for smith in (0, 1):
            med, best = timeit(lambda: kd.la_block(planes[:nn], planes[:nn], 101, 11, 1, 0.5, smith, out=out), warmup=1, iters=3)
            print(f"la smith={smith} {nn}x{nn} full: {med:.3f} ms  {nn * nn / med / 1e3:.1f} Mentries/s  {nn * nn * 10201 / med / 1e9:.3f} T cells/s")


if __name__ == "__main__":
    main()
