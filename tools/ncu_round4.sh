#!/bin/bash
# ncu evidence (round 1, v6 = end of round): launch list of a short bench run + full capture of the BN=32 conv kernel
# (four accumulator stages) and of the fused reward all-gather kernel (one rank).
set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-variants"
$B > gpurun_out/plain6.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_v6.csv $B > gpurun_out/ncu_l6.log 2>&1
C1="python tools/conv_bench.py --b 64 --s 256 --c0 32 --c1 0 --cout 32 --iters 2"
$C1 > gpurun_out/cb_r.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv3x3_umma -s 1 -c 1 -f -o gpurun_out/prof5_conv32_nacc4 $C1 > gpurun_out/ncu_r5.log 2>&1
ls -la gpurun_out/prof5*.ncu-rep
