"""bench.py contract pieces that run without a GPU: the reference arm prints one JSON line with the agreed keys, and
the product arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "gram_entries_per_sec" and d["unit"] == "entries/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "entries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("BASELINE configs[2]") and d["gpu_launches"] == 0


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)


def _dryrun(world, extra=(), n="2048"):
    """bench.py's main() with the device layer replaced by oracle-backed stand-ins (tests/bench_flow_dryrun.py)."""
    script = os.path.join(ROOT, "tests", "bench_flow_dryrun.py")
    flags = ["--gpus", str(world), "--steps", "2", "--no-extras", "--no-e2e", "--no-cpu", *extra]
    if world == 1:
        cmd = [sys.executable, script, *flags]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
               "--master-port", str(29700 + world + (7 if extra else 0)), script, *flags]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env={**os.environ, "DRY_N": n, "DRY_ROWS": "512"})
    assert out.returncode == 0, (out.stdout + out.stderr)[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]  # rank 0 alone prints, one line
    return json.loads(lines[0])


def _check_line(d, world):
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "e2e", "gpu_launches", "roofline", "clocks", "parity_checked", "exchange"):
        assert key in d, key
    assert d["metric"] == "gram_entries_per_sec" and d["unit"] == "entries/s" and d["n_gpus"] == world and d["warmup"] >= 3
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["higher_is_better"] is True and d["value"] > 0
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["peak"] > 0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and "traffic" in r
    assert d["parity_checked"] is True and d["parity"]["ranks"] == world and d["parity"]["bar"] == "bit-exact"
    assert d["gpu_launches"] == (r["phi_launches_per_step"] + r["gemm_launches_per_step"]) * d["steps"] > 0
    assert d["config"]["block_rows_built"] == world and "model" not in d["config"]


def test_product_arm_control_flow_one_gpu():
    d = _dryrun(1)
    _check_line(d, 1)
    assert d["roofline"]["gemm_launches_per_step"] == 2        # symmetric leading square + plain remainder
    assert 0.5 < d["roofline"]["entries_issued_over_entries_held"] < 1.0
    assert d["roofline"]["traffic"] is not None


def test_product_arm_control_flow_two_ranks_two_buffer_sets():
    d = _dryrun(2)
    _check_line(d, 2)
    assert d["exchange"] == {"mode": "staged", "buffers": 2}
    assert d["parity"]["tiles_per_rank"] == 16                  # both buffer sets went through the oracle check
    assert d["roofline"]["traffic"] is None                     # no ncu capture of a torchrun job


def test_product_arm_control_flow_two_ranks_single_buffer_no_remainder():
    d = _dryrun(2, extra=("--single-buffer",), n="1024")
    _check_line(d, 2)
    assert d["exchange"] == {"mode": "staged", "buffers": 1} and d["roofline"]["phi_rows_built_per_step"] == 1024


def test_reference_arm_under_torchrun_prints_on_rank_zero_only():
    """N > 1: the driver launches the reference arm like the product arm; rank 0 alone runs and prints, the others exit 0."""
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29741", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, (out.stdout + out.stderr)[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["config"]["block_rows_built"] == 2 and d["value"] > 0


def test_per_kernel_cpu_baselines():
    """bench.py's CPU side of the kernels.* entries: bounded samples of the oracle's C port, one per BASELINE config."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import bench
    import oracle_c as oc
    oc.build()
    res = bench.cpu_kernel_baselines(oc, budget_s=0.2)
    assert set(res) == {"mismatch_k10_m1_n9000", "wd_d10_n100000", "la_affine_n20000"}
    for name, base in res.items():
        assert base["value"] > 0 and base["kind"] == "port" and base["cores"] >= 1 and base["unit"] == "entries/s", (name, base)
