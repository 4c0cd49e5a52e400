#!/usr/bin/env python
"""
bench.py -- Gram entries/sec for the B200-native Gram construction (contract: see the task statement).

Headline workload (BASELINE.json configs[2]): sum of spectrum kernels k=1..7 over n = 200 000 synthetic
uniform-random 101-bp sequences (PCG64 seed 3), Gram block-rows sharded over the GPUs.  The n x n fp64 Gram
(320 GB) does not fit one GPU, so the unit of work is ONE BLOCK-ROW of 25 000 x 200 000 entries per GPU
(weak scaling: at --gpus 8 the ranks together build the complete 200k x 200k Gram; at --gpus N < 8 they
build its first N block-rows).  Every rank holds all packed sequences (6.4 MB) and builds its own int8
feature matrix Phi (4.4 GB) locally.

The job at every N is the SAME computation: rows [0, N*R) of the Gram.  Its leading N*R x N*R square is symmetric,
so -- like the reference, which computes the upper triangle and mirrors it (kernels.py:41-45) -- it is built from its
upper triangle: at N = 1 one symmetric launch with in-place mirror stores, at N > 1 the ranks share the square
(kmg/dist.py SymmetricShards: each computes about half of its part and delivers the transposed blocks to their
owners over NVLink); the columns [N*R, n) are a plain cross-Gram launch.

A "step" = spectrum_phi (packed sequences -> Phi, all 200k rows) + the tcgen05 int8 Gram GEMM launches of the
rank's block-row with the fp64 epilogue, inputs (packed sequences) resident in HBM.  Phi (4.4 GB) and the
40 GB output are far larger than the 126 MB L2, so no L2 flush is needed between iterations.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
  python bench.py --impl reference ...                          # the CPU restatement on the host cores

One JSON line on stdout (rank 0).  Extra keys beyond the contract: "roofline", "cpu_baseline", "parity", "kernels"
(device-resident numbers for the other BASELINE configs, every N), "exchange", "clocks".
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

# Warp-level instructions per Gram entry on the pipe that bounds each pairwise kernel, from the ncu captures summarised in
# profiles/r2_pipe_instructions.txt (smsp__inst_executed_pipe_*.sum / entries of the launch; uniform-random 101-bp
# sequences, the code is straight-line per pair so the count does not depend on the block shape).  Thread-level
# instructions = 32 x these / 32 entries per warp-row... see extras(): achieved = entries/s x INST_PER_ENTRY.
PIPE_INST = {
    # kernel: (pipe, thread-level instructions per computed Gram entry, microbenchmark kind that measures the pipe)
    "mismatch_k10_m1": ("alu", None, "lop3"),
    "wd_d10": ("alu", None, "lop3"),
    "la_affine": ("fp64", 10 * 10201.0, "dfma"),  # 10 FP64 instructions per DP cell by construction (csrc/la_kernel.cu)
}
try:
    with open(os.path.join(ROOT, "profiles", "r2_pipe_instructions.json")) as _f:
        for _k, _v in json.load(_f).items():
            if _k in PIPE_INST and _v.get("inst_per_entry"):
                PIPE_INST[_k] = (PIPE_INST[_k][0], float(_v["inst_per_entry"]), PIPE_INST[_k][2])
except (OSError, ValueError):
    pass


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


def workload_config(n_gpus):
    """Identical for both arms at a given N: nothing run-dependent goes in here (the exchange report is a separate key)."""
    return {"workload": "BASELINE configs[2]: sum of spectrum kernels k=1..7, n=200000 synthetic 101-bp sequences (PCG64 seed 3), "
                        "one 25000 x 200000 fp64 Gram block-row per GPU",
            "n": N_SEQ, "rows_per_gpu": ROWS_PER_GPU, "block_rows_built": n_gpus, "L": L, "ks": KS, "feature_width": D_ALG,
            "output": "fp64 (exact integers)",
            "sharding": f"block-row x{n_gpus}; the leading {n_gpus * ROWS_PER_GPU}^2 square is built from its upper triangle and mirrored "
                        "(kernels.py:41-45), the remaining columns are a plain cross-Gram",
            "l2": "inputs (Phi 4.4 GB) and output (40 GB) exceed the 126 MB L2; no flush needed"}


def sym_issued_entries(R):
    """Entries a symmetric launch over an R x R square issues to the tensor cores: the 256 x 256 tiles that touch col >= row."""
    t = -(-R // 256)
    w = [min(256, R - 256 * i) for i in range(t)]
    return float(sum(w[i] * w[j] for i in range(t) for j in range(i, t)))


# ----------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip the per-kernel numbers for the other BASELINE configs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="(debug) skip the end-to-end leg")
    ap.add_argument("--n", type=int, default=N_SEQ, help="(debug) number of sequences")
    ap.add_argument("--rows", type=int, default=ROWS_PER_GPU, help="(debug) rows per GPU")
    ap.add_argument("--no-sym", action="store_true", help="plain block-rows: no symmetric sharing of the leading square")
    ap.add_argument("--exchange", default=None, help="N > 1: staged | direct | single (default: kmg/dist.py's, KMG_SYM_EXCHANGE)")
    ap.add_argument("--single-buffer", action="store_true", help="N > 1: one set of block-row buffers (a step waits for its own peer copies)")
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

    def allmax(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allmin_int(x):
        t = torch.tensor([int(x)], dtype=torch.int32, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return int(t.item())

    n, R = args.n, args.rows
    peaks = load_peaks()
    codes = synthetic_codes(n, SEED)
    planes = kd.pack(codes, 0)  # every rank holds all packed sequences
    W = kd.phi_width(KS)
    phi = torch.empty((n, W), dtype=torch.int8, device="cuda")
    R_tot = min(n, world * R)  # the job: rows [0, R_tot) of the Gram
    sym, sym_b, exchange = None, None, {"mode": "n/a (1 GPU: in-place mirror stores)"}
    if world > 1 and not args.no_sym:
        from kmg import dist as kdist
        try:
            sym = kdist.SymmetricShards(R_tot, ldo=n, exchange=args.exchange)
            ok, note = 1, sym.exchange
        except Exception as exc:  # noqa: BLE001 - e.g. CUDA IPC not permitted in this container
            ok, note = 0, f"plain block-rows ({type(exc).__name__}: {exc})"
        if allmin_int(ok) == 0:
            if sym is not None:
                sym.close()
            sym, note = None, note if ok == 0 else "plain block-rows (another rank could not map peer memory)"
        exchange = {"mode": note}
        # Two sets of block-row buffers (2 x 57 GB per GPU at N = 8): consecutive Grams are built into alternate sets, so a
        # build does not wait for its own outgoing peer copies -- they drain under the next build -- and the build after
        # that waits for exactly those copies (kmg_gram_sharded_mark / wait_mark) before it reuses the buffers.
        if sym is not None and sym.exchange == "staged" and not args.single_buffer:
            try:
                sym_b = kdist.SymmetricShards(R_tot, ldo=n, exchange=args.exchange)
                ok = 1
            except Exception:  # noqa: BLE001 - not enough memory: single buffering
                ok = 0
            if allmin_int(ok) == 0:
                if sym_b is not None:
                    sym_b.close()
                sym_b = None
        exchange["buffers"] = 2 if sym_b is not None else 1
    elif world > 1:
        exchange = {"mode": "plain block-rows (--no-sym)"}
    syms = [s for s in (sym, sym_b) if s is not None]
    step_no = [0]
    if sym is not None:
        row0, R, out = sym.r0, sym.r1 - sym.r0, sym.block
        token = torch.zeros(1, dtype=torch.int32, device="cuda")
    else:
        row0 = (rank * R) % max(n - R + 1, 1)
        out = torch.empty((R, n), dtype=torch.float64, device="cuda")
    sym1 = world == 1 and not args.no_sym and row0 == 0 and R <= n  # N = 1: the leading R x R square is symmetric too
    issued = [0.0]
    launches = [0]

    def build():
        if sym is not None:
            cur = syms[step_no[0] % len(syms)]
            pipelined = len(syms) == 2
            issued[0] = float(cur.build_spectrum(phi[:R_tot], defer_join=pipelined or R_tot < n))
            launches[0] = cur.launches
            if R_tot < n:  # the plain remainder runs under the last peer copies of the shared square
                kd.gram_i8(phi[row0:row0 + R], phi[R_tot:], row_index0=row0, col_index0=R_tot, out_dtype=1, m_sub=0, out=cur.block[:, R_tot:])
                issued[0] += float(R) * (n - R_tot)
                launches[0] += 1
                if not pipelined:
                    cur.join()
            if pipelined:
                kd.sharded_mark(step_no[0] % 2)
        elif sym1:
            kd.gram_i8(phi[:R], phi[:R], out_dtype=1, symmetric=True, m_sub=0, out=out[:, :R])
            issued[0], launches[0] = sym_issued_entries(R), 1
            if R < n:
                kd.gram_i8(phi[:R], phi[R:], row_index0=0, col_index0=R, out_dtype=1, m_sub=0, out=out[:, R:])
                issued[0] += float(R) * (n - R)
                launches[0] += 1
        else:
            kd.gram_i8(phi[row0:row0 + R], phi, row_index0=row0, col_index0=0, out_dtype=1, m_sub=0, out=out)
            issued[0], launches[0] = float(R) * n, 1

    # Phi rows this rank's launches read: at N > 1 the column blocks beyond cyclic distance N/2 only enter tiles that other
    # ranks compute and deliver, so their features are not built here (N = 8: 5 of 8 blocks)
    phi_ranges = [(0, n)]
    if sym is not None:
        phi_ranges = sym.needed_row_ranges() + ([(R_tot, n)] if R_tot < n else [])
    phi_rows_built = sum(hi - lo for lo, hi in phi_ranges)

    def build_phi():
        for lo, hi in phi_ranges:
            kd.spectrum_phi(planes[lo:hi], L, KS, out=phi[lo:hi])

    def sync_before_step():
        """What must have happened before this step may overwrite its buffers.  One buffer set: every rank's deliveries of
        the previous step (each rank joined its copies at the end of its build).  Two sets: every rank's deliveries of the
        step before the previous one -- this rank's stream waits for its own copies of that step, the all-reduce of a
        token then orders the step after every rank's wait."""
        if sym is None:
            return
        if len(syms) == 2:
            kd.sharded_wait_mark(step_no[0] % 2)
        dist.all_reduce(token)

    def step():
        sync_before_step()
        build_phi()
        build()
        step_no[0] += 1

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
    for i in range(args.steps):
        sync_before_step()
        build_phi()
        ev[2 * i + 1].record()
        build()
        ev[2 * i + 2].record()
        step_no[0] += 1
    if sym is not None:  # the timed region ends when the last builds' deliveries are complete on every rank
        if len(syms) == 2:
            kd.sharded_wait_mark(0)
            kd.sharded_wait_mark(1)
        dist.all_reduce(token)
    ev[-1].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = allmax(ev[0].elapsed_time(ev[-1]))
    gemm_ms = [ev[2 * i + 1].elapsed_time(ev[2 * i + 2]) for i in range(args.steps)]
    ms_per_step = total_ms / args.steps
    entries_per_step = float(R_tot) * n if sym is not None else float(R) * n * world  # Gram entries DELIVERED by all ranks
    value = entries_per_step / (ms_per_step * 1e-3)
    gemm_avg_ms = float(np.mean(gemm_ms))
    gemm_max_ms = allmax(gemm_avg_ms)

    # ---- parity of the TIMED output against the CPU oracle (plain C, oracle/kmg_oracle.c): sampled tiles of this rank's
    # block-row -- in the mirrored part of its own diagonal block, in a block delivered by a peer, in a block it computed
    # for a peer, in the plain remainder and at the ragged end -- bit for bit.  Every rank checks its own rows.
    if sym is not None:
        sym.finish()
    parity = check_parity(torch, out, codes, row0, R, n, R_tot, rank, world)
    if sym_b is not None:  # both buffer sets hold a complete Gram block-row
        sym_b.finish()
        pb = check_parity(torch, sym_b.block, codes, row0, R, n, R_tot, rank, world)
        parity["ok"] = parity["ok"] and pb["ok"]
        parity["tiles_per_rank"] += pb["tiles_per_rank"]
        parity["entries_per_rank"] += pb["entries_per_rank"]
    parity["ok"] = bool(allmin_int(1 if parity["ok"] else 0))
    parity["ranks"] = world
    if not parity["ok"]:
        raise SystemExit(f"bench.py: the timed Gram block-row disagrees with the oracle: {parity}")

    # ---- roofline of the dominant kernel (gram_i8_2cta_kernel, the CTA-pair tcgen05 GEMM): tensor bound
    # 2*D ops per entry ISSUED to the tensor cores: the symmetric part of the job issues about half of
    # the entries it delivers.  Denominator: the int8 tensor-core rate measured in this run, right after the timed steps
    # (same clocks and power state), by issuing the kernel's own MMA instruction back to back with operands resident in
    # shared memory (kmg_mma_peak_i8_dev).  MEASURED_PEAKS.json has a cuBLAS bf16 rate but no int8 entry; twice its
    # sustained bf16 figure and a cuBLASLt int8 GEMM (torch._int_mm) timed in this run are reported beside it.
    alg_ops = 2.0 * D_ALG * issued[0]
    achieved_tops = alg_ops / (gemm_avg_ms * 1e-3) / 1e12
    peak_tops = kd.mma_peak_i8(iters=200000, repeats=2)
    peak_2x_bf16 = 2.0 * peaks["bf16_sustained"]
    cublaslt = cublaslt_int8_tops(torch, phi) if rank == 0 else None
    roofline = {
        "kernel": "gram_i8_2cta_kernel (tcgen05.mma.cta_group::2.kind::i8, 256x256 pair tile)", "bound": "tensor", "achieved": achieved_tops, "peak": peak_tops,
        "unit": "TOP/s (int8)", "frac": achieved_tops / peak_tops,
        "peak_source": "measured in this run: back-to-back tcgen05.mma.cta_group::2.kind::i8 256x256x32 on all 74 CTA pairs, operands in "
                       "shared memory, no loads or epilogue (csrc/mma_peak.cu); MEASURED_PEAKS.json has no int8 entry",
        "frac_of_2x_bf16_sustained": achieved_tops / peak_2x_bf16,
        "peak_2x_bf16_sustained": peak_2x_bf16, "peaks_file": peaks["source"],
        "frac_of_nominal_int8_4500": achieved_tops / NOMINAL_INT8_TOPS,
        "cublaslt_int8_same_run": cublaslt,
        "algorithmic_ops_per_step": alg_ops, "kernel_ms": gemm_avg_ms, "kernel_ms_max_over_ranks": gemm_max_ms,
        "kernel_share_of_step": gemm_avg_ms / ms_per_step,
        "traffic": TRAFFIC_BYTES_N1_STEP if sym1 else (TRAFFIC_BYTES_PER_LAUNCH if sym is None else None),
        "traffic_source": TRAFFIC_SOURCE,
        "hbm_write_gbs": 8.0 * R * n / (gemm_avg_ms * 1e-3) / 1e9,
        "gemm_launches_per_step": launches[0], "entries_issued_over_entries_held": issued[0] / (float(R) * n),
        "phi_rows_built_per_step": int(phi_rows_built), "phi_launches_per_step": len(phi_ranges),
        "entries_issued_per_s_per_gpu": issued[0] / (ms_per_step * 1e-3),
    }

    # The block-row buffers (2 x 57 GB per GPU at N = 8) are checked and done with: give the memory back before the
    # end-to-end leg and the per-kernel numbers allocate theirs.  close() synchronises and barriers, so no peer still
    # writes into a buffer that is being unmapped.
    out = None
    for s_ in syms:
        s_.close()
    sym, sym_b, syms = None, None, []
    torch.cuda.empty_cache()

    # ---- end to end through the reference-facing C-ABI with host buffers (kmg_spectrum_host)
    e2e = None
    if not args.no_e2e:
        e2e = e2e_leg(torch, kh, codes, row0, R, n, world, args, barrier, allmax)

    line = {
        "metric": "gram_entries_per_sec", "value": value, "unit": "entries/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "s8->s32->f64", "data": "synthetic", "config": workload_config(world), "e2e": e2e,
        "gpu_launches": (len(phi_ranges) + launches[0]) * args.steps, "roofline": roofline, "clocks": clocks,
        "parity_checked": parity["ok"], "parity": parity, "exchange": exchange,
        "entries_issued_per_s_per_gpu": issued[0] / (ms_per_step * 1e-3),
    }

    if not args.no_extras:
        kernels = extras(kd, torch, dist, codes, planes, peaks, rank, world, allmax, allmin_int)
        line["kernels"] = kernels
    if rank == 0 and world == 1 and not args.no_e2e:
        line["e2e"]["reference_sized_calls"] = reference_sized_calls(kh)
    if rank == 0 and world == 1 and not args.no_cpu:
        import oracle_c as oc
        oc.build()
        rows, cols = cpu_sample_shape(8.0, oc)
        v, dtc = run_cpu(rows, cols, 3, 0, oc)
        line["cpu_baseline"] = {"value": v, "unit": "entries/s", "cores": oc.num_threads(), "kind": "port",
                                "sample": f"{rows} x {cols} entries of the block-row, 3 timed passes of {dtc:.1f} s (oracle/kmg_oracle.c: "
                                          "dense Phi + dot products as kernels.py:12-47, all host threads)"}
        if "kernels" in line:  # the same CPU port on the other configs, next to their GPU numbers
            for name, base in cpu_kernel_baselines(oc).items():
                if name in line["kernels"]:
                    line["kernels"][name]["cpu_baseline"] = base
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum of one PLAIN 25000 x 200000 launch of the dominant kernel, from the
# ncu captures summarised in profiles/ (the reads are the operand panels streamed once per wave of 74 tiles; L2 eviction
# hints and other band heights were measured in round 2 and change nothing: profiles/r2_gemm_l2_hints.txt)
TRAFFIC_BYTES_PER_LAUNCH = 146_286_971_904
# the two launches of the N = 1 step (25 000^2 symmetric with TMA mirror stores + 25 000 x 175 000 plain):
# 6.70 + 4.96 GB and 93.65 + 34.98 GB, profiles/r2_gemm_ncu_full_summary.txt
TRAFFIC_BYTES_N1_STEP = 140_297_595_000
TRAFFIC_SOURCE = ("ncu --set full dram__bytes_read.sum + dram__bytes_write.sum: N = 1 step (symmetric 25000^2 + plain 25000 x 175000 launch) "
                  "100.35 + 39.95 GB, profiles/r2_gemm_ncu_full_summary.txt; one plain 25000 x 200000 launch (--no-sym) 106.30 + 39.99 GB, "
                  "profiles/r2_gemm_l2_hints.txt; null for the sharded build at N > 1 (ncu cannot follow a torchrun job); algorithmic minimum "
                  "40 GB written + 4.4 GB of Phi read once")


def check_parity(torch, out, codes, row0, R, n, R_tot, rank, world):
    """Sampled tiles of the timed output vs oracle/kmg_oracle.c, bit for bit."""
    import oracle_c as oc
    oc.build()
    rng = np.random.default_rng(1000 + rank)
    TR, TC = 24, 96
    spots = []
    r_mid = int(rng.integers(256, max(R - TR - 256, 257)))
    spots.append((max(R - TR, 0), max(row0 + 0, 0)))                         # last rows x first columns of the own diagonal block (mirrored part)
    spots.append((r_mid, min(row0 + int(rng.integers(0, max(R - TC, 1))), n - TC)))  # own diagonal block, random
    if world > 1:
        for peer in ((rank + 1) % world, (rank - 1) % world, (rank + world // 2) % world):
            c0 = peer * (R_tot // world) + int(rng.integers(0, max(R_tot // world - TC, 1)))
            spots.append((int(rng.integers(0, max(R - TR, 1))), min(c0, n - TC)))
    if R_tot < n:
        spots.append((int(rng.integers(0, max(R - TR, 1))), int(rng.integers(R_tot, n - TC))))  # plain remainder
    spots.append((max(R - TR, 0), n - TC))                                   # ragged end: last rows x last columns
    spots.append((0, 0))
    ok, entries = True, 0
    for (lr, c0) in spots:
        rows = codes[row0 + lr: row0 + lr + TR]
        cols = codes[c0: c0 + TC]
        want = oc.spectrum_block(rows, cols, KS)
        got = out[lr: lr + rows.shape[0], c0: c0 + cols.shape[0]].cpu().numpy()
        entries += want.size
        if not np.array_equal(got, want):
            ok = False
    return {"ok": ok, "tiles_per_rank": len(spots), "entries_per_rank": entries, "tile": [TR, TC],
            "oracle": "oracle/kmg_oracle.c (dense Phi + dot products, kernels.py:12-47)", "bar": "bit-exact"}


def cublaslt_int8_tops(torch, phi):
    """A library int8 GEMM (cuBLASLt through torch._int_mm, s32 out) on the same Phi, timed in this run: an independent
    second denominator for the tcgen05 kernel.  Not a product path."""
    try:
        m = 16384
        a = phi[:m].contiguous()
        b = phi[m:2 * m].t()  # (W, m) column-major view
        for _ in range(2):
            torch._int_mm(a, b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            torch._int_mm(a, b)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        return {"tops": 2.0 * m * m * phi.shape[1] / (ms * 1e-3) / 1e12, "shape": [m, m, int(phi.shape[1])], "ms": ms,
                "what": "torch._int_mm (cuBLASLt IMMA), s32 output, no fp64 epilogue"}
    except Exception as exc:  # noqa: BLE001
        return {"tops": None, "error": f"{type(exc).__name__}: {exc}"}


def e2e_leg(torch, kh, codes, row0, R, n, world, args, barrier, allmax):
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
    dt = allmax(float(np.mean(dts)))
    import oracle_c as oc
    oc.build()
    ok = bool(np.array_equal(Kh[100:116, 5000:5064], oc.spectrum_block(rows_h[100:116], codes[5000:5064], KS)))
    res = {"value": e2e_rows * float(n) * world / dt, "unit": "entries/s",
           "h2d_bytes_per_step": int((e2e_rows + n) * L), "d2h_bytes_per_step": int(e2e_rows * n * (2 if float(Kh.max()) <= 65535.0 else 4)),
           "sample": f"{e2e_rows} x {n} rows of the block-row per GPU per call through kmg_spectrum_host (numpy in, numpy fp64 out; "
                     "pageable host memory; H2D of the sequences and D2H of the Gram inside the timed region; the counts cross "
                     "PCIe as u16 when every entry of the block fits (checked on the device), else as the GEMM's s32 accumulators, "
                     "and are widened to fp64 by the copy threads; result arrays come "
                     "from libkmg's recycled host blocks, warm after the first call)",
           "first_call_value": e2e_rows * float(n) * world / cold_dt,
           "calls_timed": e2e_steps, "parity_checked": ok}
    del Kh
    if not ok:
        raise SystemExit("bench.py: the end-to-end result disagrees with the oracle")
    # The same call with the copy engine writing fp64 straight into the (pinned) result array -- no CPU in the data path.
    # Which delivery is faster depends on how many processes share the host's cores and memory system (1 GPU: widen;
    # profiles/r2_host_dram_ceiling.txt), so the library's "auto" setting times both on the first large results and keeps
    # the faster; the value reported here is that steady state.
    try:
        kh.set_d2h_mode("dma")
        Kd = kh.spectrum_gram(rows_h, KS, cols=codes)  # pins the result block
        okd = bool(np.array_equal(Kd[100:116, 5000:5064], oc.spectrum_block(rows_h[100:116], codes[5000:5064], KS)))
        dtd = []
        for _ in range(3):
            del Kd
            barrier()
            t0 = time.perf_counter()
            Kd = kh.spectrum_gram(rows_h, KS, cols=codes)
            dtd.append(time.perf_counter() - t0)
        del Kd
        dtd = allmax(float(np.mean(dtd)))
        widen_value, dma_value = res["value"], e2e_rows * float(n) * world / dtd
        res["delivery"] = {"widen_value": widen_value, "dma_value": dma_value, "dma_parity_checked": okd,
                           "widen_host_write_gbs": widen_value * 8 / 1e9, "dma_host_write_gbs": dma_value * 8 / 1e9,
                           "chosen": "dma" if (dma_value > widen_value and okd) else "widen",
                           "what": "widen = u16/s32 over PCIe + copy threads widening to fp64; dma = fp64 by the copy engine into the pinned result array; "
                                   "kmg.host.set_d2h_mode('auto') keeps the faster of the two"}
        if res["delivery"]["chosen"] == "dma":
            res["value"] = dma_value
            res["d2h_bytes_per_step"] = int(e2e_rows * n * 8)
    finally:
        kh.set_d2h_mode("widen")
    return res


def cpu_kernel_baselines(oc, budget_s=2.0):
    """The oracle's plain-C restatement (all host threads) on a bounded sample of every other BASELINE config, entries/s:
    the CPU side of the `kernels.*` entries (N = 1, rank 0; the reference itself is pure Python and cannot travel).
    Each sample is sized from a small calibration call to about `budget_s` seconds."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "dna9000.npz"))
    legs = {
        "mismatch_k10_m1_n9000": (z["codes"], lambda r, c: oc.mismatch_raw_block(r, c, 10, 1), "mismatch_raw_block (k,m)=(10,1): W^2 window tests per entry, raw counts"),
        "wd_d10_n100000": (synthetic_codes(4096, 4), lambda r, c: oc.wd_block(r, c, 10, row_index0=0, col_index0=8192), "wd_block d=10 (kernels.py:64-81 per pair)"),
        "la_affine_n20000": (synthetic_codes(4096, 5), lambda r, c: oc.la_block(r, c, 11, 1, 0.5, 0, row_index0=0, col_index0=8192), "la_block affine e=11 d=1 beta=0.5 (10 201 DP cells per entry)"),
    }
    res = {}
    for name, (codes, fn, what) in legs.items():
        try:
            t0 = time.perf_counter()
            fn(codes[:8], codes[64:128])
            rate = 8 * 64 / max(time.perf_counter() - t0, 1e-6)  # entries/s, pessimistic (thread start-up included)
            cols = 512
            rows = int(max(8, min(2048, budget_s * rate * 1.5 / cols)))
            t0 = time.perf_counter()
            fn(codes[:rows], codes[2048:2048 + cols])
            dt = time.perf_counter() - t0
            res[name] = {"value": rows * cols / dt, "unit": "entries/s", "cores": oc.num_threads(), "kind": "port",
                         "sample": f"{rows} x {cols} entries in {dt:.2f} s, oracle/kmg_oracle.c {what}"}
        except Exception as exc:  # noqa: BLE001 - a reported baseline must not cost the bench line
            res[name] = {"value": None, "error": f"{type(exc).__name__}: {exc}"}
    return res


def reference_sized_calls(kh):
    """The calls the reference's own scripts make, through the same host API (numpy in -> numpy fp64 out, best of 5 warm
    calls): BASELINE configs[0] (spectrum k=6 on Xtr0 + Xte0, 3000 sequences), configs[1] (mismatch (10,1) on all 9000
    sequences, normalised) and the nine kernels run.py builds on the 9000 sequences (run.py:6)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "dna9000.npz"))
    codes = z["codes"]
    c1 = np.ascontiguousarray(np.concatenate((codes[:2000], codes[6000:7000])))  # Xtr0 then Xte0

    def best(fn, reps=5):
        fn()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            K = fn()
            ts.append(time.perf_counter() - t0)
            del K
        return min(ts)

    res = {}
    t = best(lambda: kh.spectrum_gram(c1, [6]))
    res["configs0_spectrum_k6_n3000"] = {"ms": t * 1e3, "entries_per_s": 9e6 / t, "reference_s": 360.4,
                                         "reference_source": "BASELINE.md section 2 (unmodified kernels.py, 1 core, survey container)"}
    t = best(lambda: kh.mismatch_gram(codes, 10, 1), reps=3)
    res["configs1_mismatch_k10_m1_n9000"] = {"ms": t * 1e3, "entries_per_s": 81e6 / t, "reference_s": None,
                                             "reference_source": "not runnable by the reference (BASELINE.md: 300-420 s per sequence)"}
    nine = [("SP", 4), ("SP", 5), ("SP", 6), ("MM", 4), ("MM", 5), ("MM", 6), ("WD", 4), ("WD", 5), ("WD", 10)]

    def run_nine():
        out = []
        for kind, v in nine:
            if kind == "SP":
                out.append(kh.spectrum_gram(codes, [v]))
            elif kind == "MM":
                out.append(kh.mismatch_gram(codes, v, 1))
            else:
                out.append(kh.wd_gram(codes, v))
        return out
    t = best(run_nine, reps=2)
    res["run_py_nine_kernels_n9000"] = {"ms": t * 1e3, "entries_per_s": 9 * 81e6 / t, "methods": "SP_k4 SP_k5 SP_k6 MM_k4_m1 MM_k5_m1 MM_k6_m1 WD_d4 WD_d5 WD_d10 (run.py:6)"}
    return res


def extras(kd, torch, dist, codes, planes, peaks, rank, world, allmax, allmin_int):
    """Device-resident numbers for the other BASELINE configs at this N: every rank builds ITS block-row of the config's
    Gram (block-row sharding, no exchange: every entry depends on two sequences only), time = max over ranks (CUDA
    events), each with a sampled-tile check against the oracle and a fraction of the MEASURED peak of the pipe that
    bounds it (csrc/alu_peak.cu microbenchmarks, run here)."""
    import oracle_c as oc
    import oracle_np as onp
    oc.build()

    def timed(fn, iters=3):
        fn(); torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record(); torch.cuda.synchronize()
        return allmax(a.elapsed_time(b) / iters)

    def span(n, per):
        """this rank's rows [r0, r1) when every rank takes `per` rows (weak scaling), clipped to n"""
        r0 = min(n, rank * per)
        return r0, min(n, r0 + per)

    res = {}
    hbm = peaks["hbm_gbs"]
    pk = {k: kd.alu_peak(k) for k in ("lop3", "shf", "popc", "dfma", "dadd", "dmul")}
    res["measured_issue_peaks"] = {"unit": "thread-level instructions/s over 148 SMs", **pk,
                                   "source": "csrc/alu_peak.cu, dependent chains of the one instruction, 32 warps per SM, run in this process"}

    def pipe_frac(name, entries_computed_per_s):
        pipe, inst, kind = PIPE_INST[name]
        if inst is None:
            return {"frac": None, "peak_source": "instructions per entry not profiled yet"}
        ach = entries_computed_per_s * inst
        return {"bound": f"{pipe} pipe issue", "achieved": ach, "peak": pk[kind], "unit": "thread-level instructions/s", "frac": ach / pk[kind],
                "inst_per_entry": inst, "peak_source": f"csrc/alu_peak.cu '{kind}' microbenchmark, this run; instructions per entry from profiles/r2_pipe_instructions.txt"}

    # configs[0] kernel at scale: spectrum k=6, 32768 x 32768, symmetric, fp64 (N = 1 only: HBM-write bound)
    if world == 1:
        n6 = 32768
        phi6 = kd.spectrum_phi(planes[:n6], L, [6])
        out = torch.empty((n6, n6), dtype=torch.float64, device="cuda")
        ms = timed(lambda: kd.gram_i8(phi6, phi6, out_dtype=1, symmetric=True, out=out))
        res["spectrum_k6_sym_32768"] = {"entries_per_s": n6 * n6 / ms * 1e3, "ms": ms, "bound": "hbm", "achieved": n6 * n6 * 8 / ms / 1e6,
                                        "peak": hbm, "unit": "GB/s", "frac": n6 * n6 * 8 / ms / 1e6 / hbm, "peak_source": "MEASURED_PEAKS.json hbm_gbs"}
        del out, phi6

    # configs[1]: (k,m)=(10,1) mismatch over the 9000 real sequences, normalised fp64 (pairwise bit-vector kernel).
    # N = 1: the symmetric Gram; N > 1: block-rows of ceil(9000/N) rows (strong scaling: the config has a fixed size).
    z = np.load(os.path.join(ROOT, "tests", "golden", "dna9000.npz"))
    c2 = z["codes"]
    n2 = c2.shape[0]
    p2 = kd.pack(c2, 0)
    sd = kd.mismatch_diag_sqrt(p2, L, 10, 1)
    if world == 1:
        out = torch.empty((n2, n2), dtype=torch.float64, device="cuda")
        ms = timed(lambda: kd.mismatch_block(p2, p2, L, 10, 1, symmetric=True, sd_rows=sd, sd_cols=sd, out=out))
        r0, r1, computed = 0, n2, n2 * (n2 + 1) / 2
    else:
        per = -(-n2 // world)
        r0, r1 = span(n2, per)
        out = torch.empty((r1 - r0, n2), dtype=torch.float64, device="cuda")
        ms = timed(lambda: kd.mismatch_block(p2[r0:r1], p2, L, 10, 1, row_index0=r0, sd_rows=sd[r0:r1], sd_cols=sd, out=out))
        computed = float(per) * n2  # per GPU (the slowest rank has `per` rows)
    lr = min(100, r1 - r0 - 8)
    raw = oc.mismatch_raw_block(c2[r0 + lr:r0 + lr + 8], c2[4000:4032], 10, 1).astype(np.float64)
    dr = np.array([oc.mismatch_raw_block(c2[i:i + 1], c2[i:i + 1], 10, 1)[0, 0] for i in range(r0 + lr, r0 + lr + 8)], np.float64)
    dc = np.array([oc.mismatch_raw_block(c2[i:i + 1], c2[i:i + 1], 10, 1)[0, 0] for i in range(4000, 4032)], np.float64)
    want = raw / (np.sqrt(dr)[:, None] * np.sqrt(dc)[None, :])
    ok = bool(allmin_int(1 if np.array_equal(out[lr:lr + 8, 4000:4032].cpu().numpy(), want) else 0))
    ent = {"entries_per_s": n2 * n2 / ms * 1e3, "ms": ms, "window_pair_tests_per_s": computed * (world if world > 1 else 1) * 92 * 92 / ms * 1e3,
           "scaling": "strong (fixed 9000 x 9000 Gram)", "parity_checked": ok, "parity_bar": "bit-exact vs oracle/kmg_oracle.c on a sampled tile"}
    ent.update(pipe_frac("mismatch_k10_m1", computed / (ms * 1e-3)))
    res["mismatch_k10_m1_n9000"] = ent
    del out

    # configs[3]: weighted degree d=10, 100k synthetic sequences (seed 4): one 12500 x 100000 block-row per GPU
    n3, per3 = 100_000, 12_500
    c3 = synthetic_codes(n3, 4)
    p3 = kd.pack(c3, 0)
    r0, r1 = span(n3, per3)
    out = torch.empty((r1 - r0, n3), dtype=torch.float64, device="cuda")
    ms = timed(lambda: kd.wd_block(p3[r0:r1], p3, L, 10, row_index0=r0, out=out))
    lr = 77
    want = oc.wd_block(c3[r0 + lr:r0 + lr + 16], c3[r0 + lr - 5:r0 + lr + 59], 10, row_index0=r0 + lr, col_index0=r0 + lr - 5)  # straddles the diagonal (closed-form entries)
    ok = np.array_equal(out[lr:lr + 16, r0 + lr - 5:r0 + lr + 59].cpu().numpy(), want)
    want = oc.wd_block(c3[r1 - 16:r1], c3[n3 - 64:], 10, row_index0=r1 - 16, col_index0=n3 - 64)
    ok = bool(allmin_int(1 if (ok and np.array_equal(out[-16:, n3 - 64:].cpu().numpy(), want)) else 0))
    ent = {"entries_per_s": float(per3) * n3 * world / ms * 1e3, "ms": ms, "rows_per_gpu": per3, "scaling": "weak",
           "hbm_write_gbs": per3 * n3 * 8 / ms / 1e6, "frac_hbm_write": per3 * n3 * 8 / ms / 1e6 / hbm,
           "parity_checked": ok, "parity_bar": "bit-exact vs oracle/kmg_oracle.c on sampled tiles"}
    ent.update(pipe_frac("wd_d10", float(per3) * n3 / (ms * 1e-3)))
    res["wd_d10_n100000"] = ent
    del out

    # configs[4]: local alignment (intended recursion, affine, e=11 d=1 beta=0.5 taken literally), 20k synthetic sequences
    # (seed 5): 2500 x 20000 per GPU (8 GPUs = the whole Gram)
    n4, per4 = 20_000, 2_500
    c4 = synthetic_codes(n4, 5)
    p4 = kd.pack(c4, 0)
    r0, r1 = span(n4, per4)
    out = torch.empty((r1 - r0, n4), dtype=torch.float64, device="cuda")
    ms = timed(lambda: kd.la_block(p4[r0:r1], p4, L, 11, 1, 0.5, 0, row_index0=r0, out=out), iters=1)
    want = oc.la_block(c4[r0 + 40:r0 + 44], c4[9000:9008], 11, 1, 0.5, 0, row_index0=r0 + 40, col_index0=9000)
    got = out[40:44, 9000:9008].cpu().numpy()
    ok = bool(allmin_int(1 if np.all(np.abs(got - want) <= 1e-12 * np.abs(want)) else 0))
    ent = {"entries_per_s": float(per4) * n4 * world / ms * 1e3, "ms": ms, "rows_per_gpu": per4, "scaling": "weak",
           "dp_cells_per_s": float(per4) * n4 * world * 10201 / ms * 1e3,
           "parity_checked": ok, "parity_bar": "1e-12 relative vs oracle/kmg_oracle.c (la_block) on a sampled tile"}
    ent.update(pipe_frac("la_affine", float(per4) * n4 / (ms * 1e-3)))
    res["la_affine_n20000"] = ent
    return res


if __name__ == "__main__":
    main()
