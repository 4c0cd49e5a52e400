"""e2e_modes.py -- the end-to-end call (numpy in -> numpy fp64 out, kmg_spectrum_host) under the three delivery modes
of the host link, at N ranks sharing one host (run under torchrun).  Prints aggregate entries/s per mode.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 tools/e2e_modes.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "kernel-methods-for-genomics_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
from kmg import host as kh  # noqa: E402
import _inputs as onp  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, rows = 200_000, int(sys.argv[1]) if len(sys.argv) > 1 else 2048
codes = onp.synthetic_codes(n, 101, seed=3)
r0 = rank * 25_000
rows_h = np.ascontiguousarray(codes[r0:r0 + rows])
ks = list(range(1, 8))
ref = None
res = {}
for mode in ("widen", "dma", "mapped", "widen"):
    kh.set_d2h_mode(mode)
    K = kh.spectrum_gram(rows_h, ks, cols=codes)  # first call in the mode: pinning / cold pages
    if ref is None:
        ref = K[:64, ::997].copy()
    ok = bool(np.array_equal(K[:64, ::997], ref))
    ts = []
    for _ in range(4):
        del K
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        K = kh.spectrum_gram(rows_h, ks, cols=codes)
        ts.append(time.perf_counter() - t0)
    del K
    t = torch.tensor([float(np.mean(ts[1:]))], dtype=torch.float64, device="cuda")
    okt = torch.tensor([1 if ok else 0], device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"N={world} mode={mode:7s} {rows} x {n} per rank: {float(t.item()) * 1e3:8.1f} ms/call  aggregate {rows * float(n) * world / float(t.item()):.3e} entries/s "
              f"= {rows * float(n) * world * 8 / float(t.item()) / 1e9:.1f} GB/s of fp64 into host DRAM  parity {'ok' if int(okt.item()) else 'MISMATCH'}", flush=True)
if rank == 0:
    print(f"host: {os.cpu_count()} logical CPUs visible, affinity {len(os.sched_getaffinity(0))}", flush=True)
if dist is not None:
    dist.destroy_process_group()
