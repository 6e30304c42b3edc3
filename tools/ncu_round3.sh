#!/bin/bash
# ncu evidence (round 1, v5 = final kernels): launch list of a short bench run + full capture of the CTA-pair conv kernel.
set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-variants"
$B > gpurun_out/plain5.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_v5.csv $B > gpurun_out/ncu_l5.log 2>&1
C1="python tools/conv_bench.py --b 64 --s 32 --c0 256 --c1 0 --cout 256 --iters 2"
$C1 > gpurun_out/cb_p.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv3x3_pair -s 1 -c 1 -f -o gpurun_out/prof4_pair256 $C1 > gpurun_out/ncu_p4.log 2>&1
C2="python tools/conv_bench.py --b 64 --s 128 --c0 64 --c1 128 --cout 64 --iters 2"
$C2 > gpurun_out/cb_q.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv3x3_pair -s 1 -c 1 -f -o gpurun_out/prof4_pair192_64 $C2 > gpurun_out/ncu_q4.log 2>&1
ls -la gpurun_out/prof4*.ncu-rep
