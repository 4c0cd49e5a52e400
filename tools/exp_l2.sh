M="dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_sectors_srcunit_ltcfabric.sum,lts__t_sectors_srcnode_gpc_op_read.sum"
run() { # name env... -- args
  name=$1; shift
  python tools/kbench.py --what gemm --quick "$@" > gpurun_out/exp_$name.plain 2>&1 && ncu --metrics $M --clock-control none -k regex:gram_i8_tcgen05 -s 2 -c 1 --csv --log-file gpurun_out/exp_$name.csv python tools/kbench.py --what gemm --quick "$@" > gpurun_out/exp_$name.log 2>&1
  grep gemm gpurun_out/exp_$name.plain
}
KMG_GEMM_BAND=8 run a16k --n 16384
KMG_GEMM_BAND=8 run b25k --n 25000 --cols 200000
KMG_GEMM_BAND=4 run c25k_g4 --n 25000 --cols 200000
KMG_GEMM_BAND=1 run d25k_g1 --n 25000 --cols 200000
KMG_GEMM_BAND=2 run e25k_g2 --n 25000 --cols 200000
KMG_GEMM_BAND=8 run f4k --n 4096 --cols 200000
KMG_GEMM_BAND=8 run g25k16k --n 25000 --cols 16384
