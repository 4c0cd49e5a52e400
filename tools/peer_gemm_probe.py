"""peer_gemm_probe.py -- ONE process, two GPUs: part 0's launches of the sharded symmetric GEMM on cuda:0 with part 1's
block-row buffer living on cuda:1 (peer access), so that Nsight Compute can look at the kernel whose epilogue stores
into peer memory (ncu cannot follow a torchrun job).   python tools/peer_gemm_probe.py [direct|single|staged] [n]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "kernel-methods-for-genomics_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
from kmg import device as kd
from kmg import dist as kdist
import _inputs as onp
mode = sys.argv[1] if len(sys.argv) > 1 else "direct"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
torch.cuda.set_device(0)
bounds = kdist.sym_bounds(n, 2)
own = torch.empty((bounds[1], n), dtype=torch.float64, device="cuda:0")
peer_local = torch.empty((n - bounds[1], n), dtype=torch.float64, device="cuda:0")
peer = torch.empty((n - bounds[1], n), dtype=torch.float64, device="cuda:1")
peer[:1].copy_(own[:1])  # makes torch enable peer access 0 <-> 1
torch.cuda.synchronize()
codes = onp.synthetic_codes(n, 101, seed=3)
phi = kd.spectrum_phi(kd.pack(codes, 0), 101, list(range(1, 8)))
for name, buf in (("LOCAL second buffer", peer_local), ("PEER second buffer (cuda:1)", peer)):
    stage = None
    if mode == "staged":
        stage = torch.empty(max(kd.sharded_stage_bytes(bounds, 0, 1), 8), dtype=torch.uint8, device="cuda:0")
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        c = kd.gram_i8_sharded(phi, bounds, 0, [own.data_ptr(), buf.data_ptr()], n, stage=None if stage is None else stage.data_ptr(), exchange=mode)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"{mode}: part 0 of 2, n={n}, {name}: ms {['%.2f' % t for t in ts]}", flush=True)
