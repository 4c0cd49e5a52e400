"""
bench_flow_dryrun.py -- *** TEST INFRASTRUCTURE, not a product path ***

Runs bench.py's main() on a box WITHOUT a GPU to check its control flow and its JSON contract: the device layer
(kmg.device, kmg.dist.SymmetricShards, the torch.cuda calls bench.py makes) is replaced by stand-ins that fill the
"device" buffers from the CPU oracle, torch.distributed runs over gloo.  Nothing here is timed for real and nothing here
is reachable from the package: the product has no CPU fallback (tests/test_cpu_bench_contract.py::
test_product_arm_needs_a_gpu).  What this catches: a Python error in a branch of bench.py (N = 1 symmetric job, N > 1
shared square with one or two buffer sets, plain remainder, parity sampling, roofline / line assembly) before a GPU
minute is spent on it.  Used by tests/test_cpu_bench_contract.py, alone and under torch.distributed.run.

  DRY_N=2048 DRY_ROWS=512 python tests/bench_flow_dryrun.py --steps 2 --no-extras --no-e2e --no-cpu [--single-buffer]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
for _p in (os.path.join(ROOT, "kernel-methods-for-genomics_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, _p)

import bench  # noqa: E402
import oracle_c as oc  # noqa: E402

bench.KS = [1, 2, 3]  # small feature space: the oracle fills whole block-rows in milliseconds
bench.D_ALG = sum(4 ** k for k in bench.KS)
CODES = {}


def _strip_device(fn):
    def wrapped(*a, **kw):
        kw.pop("device", None)
        return fn(*a, **kw)
    return wrapped


class _Event:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max((other.t - self.t) * 1e3, 1e-3)


def install():
    from kmg import device as kd
    from kmg import dist as kdist
    import torch.distributed as dist

    torch.empty, torch.zeros, torch.tensor = _strip_device(torch.empty), _strip_device(torch.zeros), _strip_device(torch.tensor)
    torch.cuda.is_available = lambda: True
    torch.cuda.set_device = lambda *_a: None
    torch.cuda.synchronize = lambda *_a: None
    torch.cuda.empty_cache = lambda: None
    torch.cuda.Event = _Event
    real_init = dist.init_process_group
    dist.init_process_group = lambda backend=None, **_kw: real_init("gloo")

    L = bench.L

    def pack(codes, _fmt=0):
        CODES["all"] = np.ascontiguousarray(codes)
        return torch.arange(codes.shape[0], dtype=torch.int64).reshape(-1, 1)  # "planes": the sequence numbers

    def spectrum_phi(planes, _L, _ks, out=None):
        idx = planes[:, 0].numpy()
        out[:, :L] = torch.from_numpy(CODES["all"][idx].astype(np.int8))
        out[:, L] = 1  # built
        return out

    def rows_of(phi):
        assert bool((phi[:, L] == 1).all()), "a launch reads Phi rows that were not built in this step"
        return np.ascontiguousarray(phi[:, :L].numpy().astype(np.uint8))

    def gram_i8(a, b, row_index0=0, col_index0=0, out_dtype=1, symmetric=False, sd_rows=None, sd_cols=None, m_sub=0, out=None):
        out.copy_(torch.from_numpy(oc.spectrum_block(rows_of(a), rows_of(b), bench.KS)))
        return out

    kd.pack, kd.spectrum_phi, kd.gram_i8 = pack, spectrum_phi, gram_i8
    kd.mma_peak_i8 = lambda **_kw: 4590.0
    marks = {}
    kd.sharded_mark = lambda slot: marks.__setitem__(slot, True)

    def wait_mark(slot):
        marks.pop(slot, None)  # waiting on a slot that was never marked is allowed (first use of a buffer set)
    kd.sharded_wait_mark = wait_mark

    class Shards:
        """Stand-in for kmg.dist.SymmetricShards: same attributes; the block-row is filled from the oracle, reading only
        the Phi rows needed_row_ranges() names (plus nothing else), so a step that built too few rows fails."""
        needed_row_ranges = kdist.SymmetricShards.needed_row_ranges

        def __init__(self, n, ldo=None, exchange=None, **_kw):
            self.exchange = exchange or "staged"
            self.world, self.rank = dist.get_world_size(), dist.get_rank()
            self.n, self.ldo = n, ldo or n
            self.bounds = kdist.sym_bounds(n, self.world)
            self.r0, self.r1 = self.bounds[self.rank], self.bounds[self.rank + 1]
            self.block = torch.zeros((self.r1 - self.r0, self.ldo), dtype=torch.float64)
            self.launches = kd.sharded_launches(self.bounds, self.rank, self.exchange)  # host-side planning call of libkmg
            self.joined, self.closed = True, False

        def build_spectrum(self, phi, sd=None, defer_join=False):
            assert not self.closed and phi.shape[0] == self.n
            for lo, hi in self.needed_row_ranges():
                rows_of(phi[lo:hi])
            own = CODES["all"][self.r0:self.r1]
            self.block[:, :self.n] = torch.from_numpy(oc.spectrum_block(own, CODES["all"][:self.n], bench.KS))
            self.joined = not defer_join
            return 0.5 * (self.r1 - self.r0) * self.n

        def join(self):
            self.joined = True

        def finish(self):
            dist.barrier()

        def close(self):
            assert not self.closed
            self.closed, self.block = True, None
            dist.barrier()

    kdist.SymmetricShards = Shards


if __name__ == "__main__":
    # sizes through the environment: torch.distributed.run's own parser trips over bench.py's (debug) --n option
    sys.argv += ["--n", os.environ.get("DRY_N", "2048"), "--rows", os.environ.get("DRY_ROWS", "512")]
    install()
    bench.main()
