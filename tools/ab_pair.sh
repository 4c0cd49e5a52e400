#!/bin/bash
# timing of the pairwise kernels' variants on ONE box (round-2 A/B results: profiles/r2_la_wd_variants.txt)
echo -n "wd           : "; python tools/kbench.py --what wd --n 32768 2>&1 | grep "d=10 32768x32768 full"
echo -n "mm           : "; python tools/kbench.py --what mm --n 4096 2>&1 | grep "(10,1)"
echo -n "la 16x7      : "; KMG_LA_SHAPE=16 python tools/prof_one.py --kind la --rows 2048 --cols 4096 --iters 3 | tail -1
echo -n "la 8x13      : "; KMG_LA_SHAPE=8 python tools/prof_one.py --kind la --rows 2048 --cols 4096 --iters 3 | tail -1
