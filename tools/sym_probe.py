"""sym_probe.py -- sharded symmetric spectrum Gram over NVLink peer memory vs plain block-rows (run under torchrun).
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/sym_probe.py --size 100000"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "kernel-methods-for-genomics_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
from kmg import device as kd  # noqa: E402
from kmg import dist as kdist  # noqa: E402
import _inputs as onp  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", dest="n", type=int, default=100000)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--dtype", default="f64")
ap.add_argument("--exchange", default="direct")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = a.n
codes = onp.synthetic_codes(n, 101, seed=3)
planes = kd.pack(codes, 0)
phi = kd.spectrum_phi(planes, 101, list(range(1, 8)))
dt = torch.float64 if a.dtype == "f64" else torch.int32
shards = kdist.SymmetricShards(n, dtype=dt, exchange=a.exchange)


def timed(fn):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


ts = [timed(lambda: shards.build_spectrum(phi)) for _ in range(a.iters + 1)]
shards.finish()
# spot check against a directly computed slab of this rank's rows
r0 = shards.r0 + 256
ref = kd.gram_i8(phi[r0:r0 + 512], phi, row_index0=r0, out_dtype=1 if dt == torch.float64 else 0)
ok = bool(torch.equal(ref, shards.block[256:768]))
okt = torch.tensor([1 if ok else 0], device="cuda")
if world > 1:
    dist.all_reduce(okt, op=dist.ReduceOp.MIN)
rows = shards.r1 - shards.r0
plain = [timed(lambda: kd.gram_i8(phi[shards.r0:shards.r1], phi, row_index0=shards.r0, out_dtype=1 if dt == torch.float64 else 0,
                                  out=shards.block)) for _ in range(a.iters + 1)]
if rank == 0:
    best, bestp = min(ts[1:]), min(plain[1:])
    print(f"world={world} n={n} {a.dtype} exchange={a.exchange}: sharded-symmetric {['%.2f' % t for t in ts]} ms -> {float(n) * n / best / 1e6:.1f} Gentries/s delivered | "
          f"plain block-rows {['%.2f' % t for t in plain]} ms -> {float(n) * n / bestp / 1e6:.1f} Gentries/s | parity {'ok' if int(okt.item()) else 'MISMATCH'}",
          flush=True)
shards.close()
if world > 1:
    dist.destroy_process_group()
