#!/bin/bash
M="dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum"
run() { name=$1; shift
  python tools/kbench.py --what gemm --quick "$@" > gpurun_out/exp_$name.plain 2>&1 && KMG_GEMM_COOP=0 ncu --metrics $M --clock-control none -k regex:gram_i8_2cta -s 2 -c 1 --csv --log-file gpurun_out/exp_$name.csv python tools/kbench.py --what gemm --quick "$@" > gpurun_out/exp_$name.log 2>&1
  echo "== $name: $(grep 'm_sub=3' gpurun_out/exp_$name.plain)"
  grep -o '"dram__bytes_read.sum","[a-zA-Z]*","[0-9.,]*"\|"dram__bytes_write.sum","[a-zA-Z]*","[0-9.,]*"\|"gpu__time_duration.sum","[a-zA-Z]*","[0-9.,]*"\|"lts__t_sector_hit_rate.pct","%","[0-9.,]*"' gpurun_out/exp_$name.csv | tr '\n' ' '; echo
}
KMG_GEMM_STCS=1 KMG_GEMM_HINT=0 run stcs_h0 --n 25000 --cols 200000
KMG_GEMM_STCS=1 KMG_GEMM_HINT=3 run stcs_h3 --n 25000 --cols 200000
KMG_GEMM_STCS=1 KMG_GEMM_HINT=1 run stcs_h1 --n 25000 --cols 200000
KMG_GEMM_STCS=1 KMG_GEMM_HINT=1 KMG_GEMM_BAND=6 run stcs_h1_b6 --n 25000 --cols 200000
