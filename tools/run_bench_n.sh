#!/bin/bash
# bench.py at N GPUs of one box, JSON line kept under gpurun_out/, one-screen summary on stdout
N=${1:-1}; TAG=${2:-r2h}
mkdir -p gpurun_out
if [ "$N" = 1 ]; then CMD="python bench.py"; else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N"; fi
timeout 900 $CMD 2> gpurun_out/${TAG}_bench_n$N.err | grep "^{" > gpurun_out/${TAG}_bench_n$N.json
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_n$N.json")); r=d["roofline"]; e=d["e2e"]
print("N=%d value %.4g ms/step %.2f frac %.3f kernel_ms %.2f (max %.2f) share %.3f launches %d issued/s/gpu %.3g parity %s %s" % (d["n_gpus"], d["value"], d["ms_per_step"], r["frac"], r["kernel_ms"], r["kernel_ms_max_over_ranks"], r["kernel_share_of_step"], r["gemm_launches_per_step"], d["entries_issued_per_s_per_gpu"], d["parity_checked"], d["exchange"]))
print("   e2e %.3g (widen %.3g, dma %.3g, chosen %s)  clocks %s" % (e["value"], e["delivery"]["widen_value"], e["delivery"]["dma_value"], e["delivery"]["chosen"], d["clocks"]))
for k,v in d["kernels"].items():
    if k!="measured_issue_peaks": print("   ", k, "%.4g entries/s" % v["entries_per_s"], "frac", v.get("frac"), "parity", v.get("parity_checked"))
PY
