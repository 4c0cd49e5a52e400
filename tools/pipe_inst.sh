#!/bin/bash
# warp-level instruction counts per pipe of the pairwise kernels (ncu), for the per-kernel roofline fractions of bench.py
M="smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_xu.sum,smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_lsu.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active"
for spec in "mm 4096 4096" "wd 8192 16384" "la 512 4096"; do
  set -- $spec
  python tools/prof_one.py --kind $1 --rows $2 --cols $3 --iters 2 > gpurun_out/r2_pipe_$1.plain 2>&1 && \
  ncu --metrics $M --clock-control none -k regex:'mismatch_kernel|wd_kernel|la_kernel' -c 1 --csv --log-file gpurun_out/r2_pipe_$1.csv python tools/prof_one.py --kind $1 --rows $2 --cols $3 --iters 1 > gpurun_out/r2_pipe_$1.log 2>&1
  cat gpurun_out/r2_pipe_$1.plain | tail -1
done
