#!/usr/bin/env python
"""
bench.py -- Gram entries/sec for the B200-native Gram construction (contract: see the task statement).

Headline workload (BASELINE.json configs[2]): sum of spectrum kernels k=1..7 over n = 200 000 synthetic
uniform-random 101-bp sequences (PCG64 seed 3), Gram block-rows sharded over the GPUs.  The n x n fp64 Gram
(320 GB) does not fit one GPU, so the unit of work is ONE BLOCK-ROW of 25 000 x 200 000 entries per GPU
(weak scaling: at --gpus 8 the ranks together build the complete 200k x 200k Gram; at --gpus N < 8 they
build its first N block-rows).  Every rank holds all packed sequences (6.4 MB) and builds its own int8
feature matrix Phi (4.4 GB) locally: no collective on the data path.

A "step" = spectrum_phi (packed sequences -> Phi, all 200k rows) + the tcgen05 int8 Gram GEMM of the
rank's block-row with the fp64 epilogue, inputs (packed sequences) resident in HBM.  Phi (4.4 GB) and the
40 GB output are far larger than the 126 MB L2, so no L2 flush is needed between iterations.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
  python bench.py --impl reference ...                          # the CPU restatement on the host cores

One JSON line on stdout (rank 0).  Extra keys beyond the contract: "roofline", "cpu_baseline", "kernels"
(device-resident numbers for the other BASELINE configs at N=1), "clocks".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "kernel-methods-for-genomics_b200"), os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

N_SEQ = 200_000
ROWS_PER_GPU = 25_000
L = 101
KS = list(range(1, 8))
D_ALG = sum(4 ** k for k in KS)  # 21 844 algorithmic feature width (padding to 21 888 is not counted)
SEED = 3
NOMINAL_INT8_TOPS = 4500.0
NOMINAL_HBM_GBS = 7700.0


def synthetic_codes(n, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.integers(0, 4, size=(n, L), dtype=np.uint8)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_burst": float(p["bf16_tflops"]),
                "bf16_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the plain-C restatement of the reference's algorithm (oracle/kmg_oracle.c,
# dense Phi + dot products as kernels.py:12-47 does) on all host threads, on a bounded sample of the workload
# ----------------------------------------------------------------------------------------------------
def cpu_sample_shape(budget_s, oc):
    """rows x cols of the block-row the CPU can do in ~budget_s seconds (calibrated on a tiny sample)."""
    codes = synthetic_codes(1024, SEED)
    t0 = time.perf_counter()
    oc.spectrum_block(codes[:16], codes[:256], KS)
    rate = 16 * 256 / max(time.perf_counter() - t0, 1e-6)  # entries/s incl. feature build (pessimistic)
    cols = 4096
    rows = int(max(16, min(8192, budget_s * rate * 2 / cols)))
    return rows, cols


def run_cpu(rows, cols, steps, warmup, oc):
    codes = synthetic_codes(max(rows, cols), SEED)
    r, c = codes[:rows], codes[:cols]
    for _ in range(warmup):
        oc.spectrum_block(r, c, KS)
    t0 = time.perf_counter()
    for _ in range(steps):
        oc.spectrum_block(r, c, KS)
    dt = (time.perf_counter() - t0) / steps
    return rows * cols / dt, dt


def reference_arm(args, rank):
    if rank != 0:
        return
    import oracle_c as oc
    oc.build()
    threads = oc.num_threads()
    rows, cols = cpu_sample_shape(2.0, oc)
    value, dt = run_cpu(rows, cols, args.steps, args.warmup, oc)
    sample = f"{rows} x {cols} entries of the block-row per step (k=1..7 dense Phi + dot products, oracle/kmg_oracle.c)"
    line = {
        "impl": "reference", "metric": "gram_entries_per_sec", "value": value, "unit": "entries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "s8->s32->f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "entries/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "entries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference is single-threaded pure Python (2.5e4 entries/s on config 1, BASELINE.md) and cannot travel to the GPU box; "
                "this arm times its algorithm restated in C on all host threads",
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus, exchange="n/a (1 GPU)", exchange_checked=None):
    return {"workload": "BASELINE configs[2]: sum of spectrum kernels k=1..7, n=200000 synthetic 101-bp sequences (PCG64 seed 3), "
                        "one 25000 x 200000 fp64 Gram block-row per GPU",
            "n": N_SEQ, "rows_per_gpu": ROWS_PER_GPU, "block_rows_built": n_gpus, "L": L, "ks": KS, "feature_width": D_ALG,
            "output": "fp64 (exact integers)", "sharding": f"block-row x{n_gpus} (boundaries on multiples of 256 rows when N > 1), no collective on the data path",
            "exchange": exchange, "exchange_checked": exchange_checked,
            "l2": "inputs (Phi 4.4 GB) and output (40 GB) exceed the 126 MB L2; no flush needed"}


# ----------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip the per-kernel numbers for the other BASELINE configs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--n", type=int, default=N_SEQ, help="(debug) number of sequences")
    ap.add_argument("--rows", type=int, default=ROWS_PER_GPU, help="(debug) rows per GPU")
    ap.add_argument("--no-sym", action="store_true", help="N > 1: plain block-rows, no symmetric sharing between the GPUs")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, rank)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from kmg import device as kd
    from kmg import host as kh
    from kmg._cabi import lib
    lib()

    n, R = args.n, args.rows
    peaks = load_peaks()
    codes = synthetic_codes(n, SEED)
    planes = kd.pack(codes, 0)  # every rank holds all packed sequences
    W = kd.phi_width(KS)
    phi = torch.empty((n, W), dtype=torch.int8, device="cuda")
    # N > 1: the job is rows [0, N*R) of the Gram.  Its leading N*R x N*R square is symmetric, so the ranks share it: each
    # computes about half of its part of the square and ships the transposed blocks to their owners over NVLink
    # (kmg/dist.py SymmetricShards: CUDA IPC buffers, one pitched peer copy per block on the copy stream); the columns
    # [N*R, n) of every block-row are a plain cross-Gram launch.  One all-reduce of a token per step is the barrier that
    # orders step k+1 after every rank's copies of step k.
    sym, sym_note, R_tot = None, "n/a (1 GPU)", min(n, world * R)
    if world > 1 and not args.no_sym:
        from kmg import dist as kdist
        try:
            sym = kdist.SymmetricShards(R_tot, ldo=n)
            ok, sym_note = 1, "symmetric square shared over NVLink peer memory"
        except Exception as exc:  # noqa: BLE001 - e.g. CUDA IPC not permitted in this container
            ok, sym_note = 0, f"plain block-rows ({type(exc).__name__}: {exc})"
        flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            if sym is not None:
                sym.close()
            sym, sym_note = None, sym_note if ok == 0 else "plain block-rows (another rank could not map peer memory)"
    elif world > 1:
        sym_note = "plain block-rows (--no-sym)"
    if sym is not None:
        row0, R, out = sym.r0, sym.r1 - sym.r0, sym.block
        g = world
        gemm_launches = sum(1 for d in range(1, g) if 2 * d < g) + (1 if g % 2 == 0 else 0) + 1 + (1 if R_tot < n else 0)
        token = torch.zeros(1, dtype=torch.int32, device="cuda")
    else:
        row0 = (rank * R) % max(n - R + 1, 1)
        out = torch.empty((R, n), dtype=torch.float64, device="cuda")
        gemm_launches = 1
    issued = [0]

    def build():
        if sym is None:
            kd.gram_i8(phi[row0:row0 + R], phi, row_index0=row0, col_index0=0, out_dtype=1, m_sub=0, out=out)
            issued[0] = R * n
            return
        issued[0] = sym.build_spectrum(phi[:R_tot])
        if R_tot < n:
            kd.gram_i8(phi[row0:row0 + R], phi[R_tot:], row_index0=row0, col_index0=R_tot, out_dtype=1, m_sub=0, out=out[:, R_tot:])
            issued[0] += R * (n - R_tot)

    def step():
        kd.spectrum_phi(planes, L, KS, out=phi)
        build()
        if sym is not None:
            dist.all_reduce(token)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 2)]
    ev[0].record()
    gemm_ms = []
    for i in range(args.steps):
        kd.spectrum_phi(planes, L, KS, out=phi)
        ev[2 * i + 1].record()
        build()
        ev[2 * i + 2].record()
        if sym is not None:
            dist.all_reduce(token)
    ev[-1].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = ev[0].elapsed_time(ev[-1])
    gemm_ms = [ev[2 * i + 1].elapsed_time(ev[2 * i + 2]) for i in range(args.steps)]
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    entries_per_step = float(R_tot) * n if sym is not None else float(R) * n * world  # Gram entries DELIVERED by all ranks
    value = entries_per_step / (ms_per_step * 1e-3)
    sym_checked = None
    if sym is not None:
        # parity of the shared build: 512 of this rank's rows against a direct launch of the same kernel
        sym.finish()
        lo = min(256, max(R - 512, 0))
        direct = kd.gram_i8(phi[row0 + lo:row0 + lo + 512], phi, row_index0=row0 + lo, col_index0=0, out_dtype=1, m_sub=0)
        okt = torch.tensor([1 if torch.equal(direct, out[lo:lo + 512]) else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        sym_checked = bool(int(okt.item()))
        del direct
        if not sym_checked:
            raise SystemExit("bench.py: the shared symmetric build disagrees with a direct launch")

    # ---- roofline of the dominant kernel (gram_i8_2cta_kernel, the CTA-pair tcgen05 GEMM): tensor bound
    gemm_avg_ms = float(np.mean(gemm_ms))
    # 2*D ops per entry ISSUED to the tensor cores: at N = 1 one launch = one block-row, every entry issued;
    # in the shared symmetric build a rank issues about half of the entries it ends up holding
    alg_ops = 2.0 * D_ALG * float(issued[0])
    achieved_tops = alg_ops / (gemm_avg_ms * 1e-3) / 1e12
    # Denominator: the int8 tensor-core rate measured in this run, right after the timed steps (same clocks and power
    # state), by issuing the kernel's own MMA instruction back to back with operands resident in shared memory
    # (kmg_mma_peak_i8_dev, ~55 ms per sample).  MEASURED_PEAKS.json has a cuBLAS bf16 rate but no int8 entry; twice its
    # sustained bf16 figure is reported beside it.
    peak_tops = kd.mma_peak_i8(iters=200000, repeats=2)
    peak_2x_bf16 = 2.0 * peaks["bf16_sustained"]
    roofline = {
        "kernel": "gram_i8_2cta_kernel (tcgen05.mma.cta_group::2.kind::i8, 256x256 pair tile)", "bound": "tensor", "achieved": achieved_tops, "peak": peak_tops,
        "unit": "TOP/s (int8)", "frac": achieved_tops / peak_tops,
        "peak_source": "measured in this run: back-to-back tcgen05.mma.cta_group::2.kind::i8 256x256x32 on all 74 CTA pairs, operands in "
                       "shared memory, no loads or epilogue (csrc/mma_peak.cu); MEASURED_PEAKS.json has no int8 entry",
        "frac_of_2x_bf16_sustained": achieved_tops / peak_2x_bf16,
        "peak_2x_bf16_sustained": peak_2x_bf16, "peaks_file": peaks["source"],
        "frac_of_nominal_int8_4500": achieved_tops / NOMINAL_INT8_TOPS,
        "algorithmic_ops_per_launch": alg_ops, "kernel_ms": gemm_avg_ms, "kernel_share_of_step": gemm_avg_ms / ms_per_step,
        "traffic": TRAFFIC_BYTES_PER_LAUNCH if sym is None else None,
        "hbm_write_gbs": 8.0 * R * n / (gemm_avg_ms * 1e-3) / 1e9,
        "gemm_launches_per_step": gemm_launches, "entries_issued_over_entries_held": float(issued[0]) / (float(R) * n),
    }

    # ---- end to end through the reference-facing C-ABI with host buffers (kmg_spectrum_host)
    e2e_rows = min(2048, R)
    rows_h = np.ascontiguousarray(codes[row0:row0 + e2e_rows])
    kh.spectrum_gram(rows_h[:256], KS, cols=codes)  # warm the path (allocations, tile list)
    barrier()
    t0 = time.perf_counter()
    Kh = kh.spectrum_gram(rows_h, KS, cols=codes)  # first full-size call of the process: cold result memory
    cold_dt = time.perf_counter() - t0
    for _ in range(max(args.warmup - 1, 0)):
        del Kh
        Kh = kh.spectrum_gram(rows_h, KS, cols=codes)
    barrier()
    e2e_steps, dts = 5, []
    for _ in range(e2e_steps):
        del Kh  # the caller drops the previous 3.3 GB result before asking for the next one (not part of the call)
        t0 = time.perf_counter()
        Kh = kh.spectrum_gram(rows_h, KS, cols=codes)
        dts.append(time.perf_counter() - t0)
    dt = float(np.mean(dts))
    te = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e = {"value": e2e_rows * float(n) * world / float(te.item()), "unit": "entries/s",
           "h2d_bytes_per_step": int((e2e_rows + n) * L), "d2h_bytes_per_step": int(e2e_rows * n * (2 if float(Kh.max()) <= 65535.0 else 4)),
           "sample": f"{e2e_rows} x {n} rows of the block-row per GPU per call through kmg_spectrum_host (numpy in, numpy fp64 out; "
                     "pageable host memory; H2D of the sequences and D2H of the Gram inside the timed region; the counts cross "
                     "PCIe as u16 when every entry of the block fits (checked on the device), else as the GEMM's s32 accumulators, "
                     "and are widened to fp64 by the copy threads; result arrays come "
                     "from libkmg's recycled host blocks, warm after the first call)",
           "first_call_value": e2e_rows * float(n) * world / cold_dt,
           "calls_timed": e2e_steps, "checksum": float(Kh[0, :8].sum())}
    del Kh

    line = {
        "metric": "gram_entries_per_sec", "value": value, "unit": "entries/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "s8->s32->f64", "data": "synthetic", "config": workload_config(world, sym_note, sym_checked), "e2e": e2e,
        "gpu_launches": (1 + gemm_launches) * args.steps, "roofline": roofline, "clocks": clocks,
    }

    if rank == 0 and world == 1 and not args.no_extras:
        del out
        torch.cuda.empty_cache()
        line["kernels"] = extras(kd, torch, codes, planes, phi, peaks)
    if rank == 0 and world == 1 and not args.no_cpu:
        import oracle_c as oc
        oc.build()
        rows, cols = cpu_sample_shape(8.0, oc)
        v, dtc = run_cpu(rows, cols, 3, 0, oc)
        line["cpu_baseline"] = {"value": v, "unit": "entries/s", "cores": oc.num_threads(), "kind": "port",
                                "sample": f"{rows} x {cols} entries of the block-row, 3 timed passes of {dtc:.1f} s (oracle/kmg_oracle.c: "
                                          "dense Phi + dot products as kernels.py:12-47, all host threads)"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if sym is not None:
        out = None
        sym.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, from the ncu --set full
# capture summarised in profiles/r1_gemm_ncu_full_summary.txt (106.30 GB read + 39.99 GB written; the algorithmic
# minimum is 40 GB written + 4.4 GB of Phi read once -- the reads are the B panels re-streamed once per 8-row-tile band)
TRAFFIC_BYTES_PER_LAUNCH = 146_276_310_000


def extras(kd, torch, codes, planes, phi, peaks):
    """Device-resident numbers for the other BASELINE configs (N=1 only): each timed with CUDA events over 3 launches."""
    def timed(fn, iters=3):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    res = {}
    hbm = peaks["hbm_gbs"]
    # spectrum k=6 (configs[0] kernel at scale): 32768 x 32768 block, symmetric, fp64
    n6 = 32768
    phi6 = kd.spectrum_phi(planes[:n6], L, [6])
    out = torch.empty((n6, n6), dtype=torch.float64, device="cuda")
    ms = timed(lambda: kd.gram_i8(phi6, phi6, out_dtype=1, symmetric=True, out=out))
    res["spectrum_k6_sym_32768"] = {"entries_per_s": n6 * n6 / ms * 1e3, "ms": ms, "hbm_write_gbs": n6 * n6 * 8 / ms / 1e6,
                                    "frac_hbm_write": n6 * n6 * 8 / ms / 1e6 / hbm}
    del out, phi6
    # configs[1]: (k,m)=(10,1) mismatch over 9000 sequences, symmetric, normalised fp64 (pairwise bit-vector kernel)
    n2 = 9000
    sd = kd.mismatch_diag_sqrt(planes[:n2], L, 10, 1)
    out = torch.empty((n2, n2), dtype=torch.float64, device="cuda")
    ms = timed(lambda: kd.mismatch_block(planes[:n2], planes[:n2], L, 10, 1, symmetric=True, sd_rows=sd, sd_cols=sd, out=out))
    res["mismatch_k10_m1_sym_9000"] = {"entries_per_s": n2 * n2 / ms * 1e3, "ms": ms,
                                       "window_pair_tests_per_s": (n2 * (n2 + 1) / 2) * 92 * 92 / ms * 1e3}
    del out
    # configs[3]: weighted degree d=10, 100k sequences: one 12500 x 100000 block-row
    n3, r3 = 100_000, 12_500
    out = torch.empty((r3, n3), dtype=torch.float64, device="cuda")
    ms = timed(lambda: kd.wd_block(planes[:r3], planes[:n3], L, 10, out=out))
    res["wd_d10_blockrow_12500x100000"] = {"entries_per_s": r3 * n3 / ms * 1e3, "ms": ms, "hbm_write_gbs": r3 * n3 * 8 / ms / 1e6,
                                           "frac_hbm_write": r3 * n3 * 8 / ms / 1e6 / hbm}
    del out
    # configs[4]: local alignment (intended recursion), 20k sequences: one 1024 x 20000 block
    n4, r4 = 20_000, 1024
    out = torch.empty((r4, n4), dtype=torch.float64, device="cuda")
    ms = timed(lambda: kd.la_block(planes[:r4], planes[:n4], L, 11, 1, 0.5, 0, out=out), iters=1)
    res["la_affine_block_1024x20000"] = {"entries_per_s": r4 * n4 / ms * 1e3, "ms": ms, "dp_cells_per_s": r4 * n4 * 10201 / ms * 1e3}
    return res


if __name__ == "__main__":
    main()
