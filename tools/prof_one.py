"""prof_one.py -- run one kernel configuration a few times (target for ncu)."""
import argparse, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "kernel-methods-for-genomics_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
from kmg import device as kd
import _inputs as onp
ap = argparse.ArgumentParser()
ap.add_argument("--kind", default="gemm")
ap.add_argument("--rows", type=int, default=16384)
ap.add_argument("--cols", type=int, default=16384)
ap.add_argument("--ks", default="6")
ap.add_argument("--m_sub", type=int, default=3)
ap.add_argument("--dt", type=int, default=1)
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
n = max(a.rows, a.cols)
planes = kd.pack(onp.synthetic_codes(n, 101, seed=3), 0)
if a.kind == "gemm":
    ks = [int(x) for x in a.ks.split(",")]
    phi = kd.spectrum_phi(planes, 101, ks)
    out = torch.empty((a.rows, a.cols), dtype=torch.float64 if a.dt else torch.int32, device="cuda")
    fn = lambda: kd.gram_i8(phi[:a.rows], phi[:a.cols], out_dtype=a.dt, m_sub=a.m_sub, out=out)
elif a.kind == "n1job":
    # the N = 1 job of bench.py: 25 000^2 symmetric (in-place TMA mirror stores) + 25 000 x 175 000 plain
    R, nn = 25000, 200000
    planes = kd.pack(onp.synthetic_codes(nn, 101, seed=3), 0)
    phi = kd.spectrum_phi(planes, 101, [1, 2, 3, 4, 5, 6, 7])
    out = torch.empty((R, nn), dtype=torch.float64, device="cuda")
    def fn():
        kd.gram_i8(phi[:R], phi[:R], out_dtype=1, symmetric=True, out=out[:, :R])
        kd.gram_i8(phi[:R], phi[R:], row_index0=0, col_index0=R, out_dtype=1, out=out[:, R:])
elif a.kind == "wd":
    out = torch.empty((a.rows, a.cols), dtype=torch.float64, device="cuda")
    fn = lambda: kd.wd_block(planes[:a.rows], planes[:a.cols], 101, 10, out=out)
elif a.kind == "mm":
    out = torch.empty((a.rows, a.cols), dtype=torch.float64, device="cuda")
    fn = lambda: kd.mismatch_block(planes[:a.rows], planes[:a.cols], 101, 10, 1, out=out)
elif a.kind == "la":
    out = torch.empty((a.rows, a.cols), dtype=torch.float64, device="cuda")
    fn = lambda: kd.la_block(planes[:a.rows], planes[:a.cols], 101, 11, 1, 0.5, 0, out=out)
elif a.kind == "all":
    # one launch of every non-GEMM kernel on the path (ncu target for profiles/r1_*_ncu_summary.txt)
    o1 = torch.empty((8192, 16384), dtype=torch.float64, device="cuda")
    o2 = torch.empty((4096, 4096), dtype=torch.float64, device="cuda")
    o3 = torch.empty((512, 4096), dtype=torch.float64, device="cuda")
    sq = torch.rand((8192, 8192), dtype=torch.float64, device="cuda") + 1.0
    def fn():
        kd.wd_block(planes[:8192], planes[:16384], 101, 10, out=o1)
        kd.wds_block(planes[:512], planes[:4096], 101, 3, 2, out=o3)
        kd.mismatch_block(planes[:4096], planes[:4096], 101, 10, 1, out=o2)
        kd.mismatch_block(planes[:4096], planes[:4096], 101, 6, 2, out=o2)
        kd.la_block(planes[:512], planes[:4096], 101, 11, 1, 0.5, 0, out=o3)
        kd.la_block(planes[:512], planes[:4096], 101, 11, 1, 0.5, 1, out=o3)
        phi = kd.mismatch_phi(planes[:16384], 101, 6, 1)
        kd.spectrum_phi(planes[:16384], 101, [1, 2, 3, 4, 5, 6, 7])
        kd.normalize_(sq)
        kd.center(sq)
        kd.combine([sq, sq, sq], [0.2, 0.3, 0.5], 2)
ts = []
for _ in range(a.iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(a.kind, a.rows, a.cols, "ms:", ["%.3f" % t for t in ts])
