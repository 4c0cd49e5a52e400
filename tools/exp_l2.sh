#!/bin/bash
# L2 / rasterisation experiments on the 25000 x 200000 block-row GEMM: time without ncu, then DRAM bytes under ncu.
M="dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum"
run() { # name -- args
  name=$1; shift
  python tools/kbench.py --what gemm --quick "$@" > gpurun_out/exp_$name.plain 2>&1 && KMG_GEMM_COOP=0 ncu --metrics $M --clock-control none -k regex:gram_i8_2cta -s 2 -c 1 --csv --log-file gpurun_out/exp_$name.csv python tools/kbench.py --what gemm --quick "$@" > gpurun_out/exp_$name.log 2>&1
  echo "== $name: $(grep 'm_sub=3' gpurun_out/exp_$name.plain)"
  grep -o '"dram__bytes_read.sum","[a-zA-Z]*","[0-9.,]*"\|"dram__bytes_write.sum","[a-zA-Z]*","[0-9.,]*"\|"gpu__time_duration.sum","[a-zA-Z]*","[0-9.,]*"\|"lts__t_sector_hit_rate.pct","%","[0-9.,]*"' gpurun_out/exp_$name.csv | tr '\n' ' '; echo
}
for hint in 0 1 2 3; do
  for band in 8 12; do
    KMG_GEMM_HINT=$hint KMG_GEMM_BAND=$band run h${hint}_b${band} --n 25000 --cols 200000
  done
done
