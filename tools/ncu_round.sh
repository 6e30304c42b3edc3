#!/bin/bash
# ncu evidence for profiles/: launch list of a short bench run + full captures of three conv kernels.
set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-variants"
$B > gpurun_out/plain3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_v3.csv $B > gpurun_out/ncu_l3.log 2>&1
C1="python tools/conv_bench.py --b 64 --s 256 --c0 32 --c1 0 --cout 32 --iters 2"
$C1 > gpurun_out/cb_a.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv3x3 -s 1 -c 1 -f -o gpurun_out/prof2_conv32 $C1 > gpurun_out/ncu_a.log 2>&1
C2="python tools/conv_bench.py --b 64 --s 256 --c0 32 --c1 64 --cout 32 --iters 2"
$C2 > gpurun_out/cb_b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv3x3 -s 1 -c 1 -f -o gpurun_out/prof2_kws96 $C2 > gpurun_out/ncu_b.log 2>&1
C3="python tools/conv_bench.py --b 64 --s 32 --c0 256 --c1 0 --cout 256 --iters 2"
$C3 > gpurun_out/cb_c.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv3x3 -s 1 -c 1 -f -o gpurun_out/prof2_conv256 $C3 > gpurun_out/ncu_c.log 2>&1
tail -2 gpurun_out/ncu_a.log gpurun_out/ncu_b.log gpurun_out/ncu_c.log gpurun_out/ncu_l3.log
ls -la gpurun_out/*.ncu-rep
