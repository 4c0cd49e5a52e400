"""sym_local_probe.py -- ONE part's launch sequence of the sharded symmetric GEMM with every buffer on ONE GPU (no
NVLink): isolates the cost of the tile assignment / doubled epilogue stores / staging from the cost of the link.
All peers' block-rows alias one dummy buffer (their contents are irrelevant here), so n = 200 000, world = 8 fits.
   python tools/sym_local_probe.py 200000 8 direct,staged,single"""
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
modes = (sys.argv[3] if len(sys.argv) > 3 else "direct,staged,single").split(",")
part = int(sys.argv[4]) if len(sys.argv) > 4 else 0
codes = onp.synthetic_codes(n, 101, seed=3)
phi = kd.spectrum_phi(kd.pack(codes, 0), 101, list(range(1, 8)))
bounds = kdist.sym_bounds(n, world)
rows_max = max(bounds[p + 1] - bounds[p] for p in range(world))
own = torch.empty((bounds[part + 1] - bounds[part], n), dtype=torch.float64, device="cuda")
dummy = torch.empty((rows_max, n), dtype=torch.float64, device="cuda")
ptrs = [own.data_ptr() if p == part else dummy.data_ptr() for p in range(world)]
for mode in modes:
    stage = None
    if mode == "staged":
        stage = torch.empty(max(kd.sharded_stage_bytes(bounds, part, 1), 8), dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        c = kd.gram_i8_sharded(phi, bounds, part, ptrs, n, stage=None if stage is None else stage.data_ptr(), exchange=mode)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    rows = bounds[part + 1] - bounds[part]
    print(f"part {part}/{world} n={n} exchange={mode} TMA_MIRROR={os.environ.get('KMG_GEMM_TMA_MIRROR', '1')}: rows {rows} computed {c / (rows * n):.3f} of the block-row, "
          f"ms {['%.2f' % t for t in ts]}  {2.0 * c * 21844 / min(ts[1:]) / 1e9:.0f} TOPS on issued tiles", flush=True)
    del stage
# the N = 1 job of bench.py: leading square symmetric (in-place mirror) + plain remainder
R = bounds[1] - bounds[0]
for tma in ("1", "0"):
    ts = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        kd.gram_i8(phi[:R], phi[:R], symmetric=True, out=own[:, :R])
        if R < n:
            kd.gram_i8(phi[:R], phi[R:], col_index0=R, out=own[:, R:])
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"N=1 job ({R} x {n}: symmetric leading square + plain remainder): ms {['%.2f' % t for t in ts]}", flush=True)
    break
ts = []
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); kd.gram_i8(phi[:R], phi, out=own); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(f"plain block-row {R} x {n}: ms {['%.2f' % t for t in ts]}", flush=True)
