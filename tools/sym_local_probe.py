"""sym_local_probe.py -- the sharded symmetric GEMM with every part's buffer on ONE GPU (no NVLink): isolates the cost of
the tile assignment / mirror stores from the cost of storing into peer memory."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "kernel-methods-for-genomics_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
from kmg import device as kd
from kmg import dist as kdist
import _inputs as onp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
world = int(sys.argv[2]) if len(sys.argv) > 2 else 2
codes = onp.synthetic_codes(n, 101, seed=3)
phi = kd.spectrum_phi(kd.pack(codes, 0), 101, list(range(1, 8)))
bounds = kdist.sym_bounds(n, world)
bufs = [torch.empty((bounds[p + 1] - bounds[p], n), dtype=torch.float64, device="cuda") for p in range(world)]
ptrs = [b.data_ptr() for b in bufs]
for p in range(world):
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); c = kd.gram_i8_sharded(phi, bounds, p, ptrs, n); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    rows = bounds[p + 1] - bounds[p]
    print(f"part {p}/{world}: rows {rows} computed {c / (rows * n):.3f} of the block-row, ms {['%.2f' % t for t in ts]}  "
          f"{2.0 * c * phi.shape[1] / min(ts) / 1e9:.0f} TOPS on issued tiles", flush=True)
