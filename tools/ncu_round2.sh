#!/bin/bash
# ncu evidence (round 1, v4): launch list of a short bench run + full captures of the two FFT-prox kernels.
set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-variants"
$B > gpurun_out/plain4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_v4.csv $B > gpurun_out/ncu_l4.log 2>&1
P1="python tools/prox_bench.py --cases 256x256c --iters 2"
$P1 > gpurun_out/pb_a.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fftprox_rows256 -s 1 -c 1 -f -o gpurun_out/prof3_prox_rows $P1 > gpurun_out/ncu_p1.log 2>&1
P2="python tools/prox_bench.py --cases 256x256r --iters 2"
$P2 > gpurun_out/pb_b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fftprox_fused2 -s 1 -c 1 -f -o gpurun_out/prof3_prox_fused2 $P2 > gpurun_out/ncu_p2.log 2>&1
P3="python tools/psnr_bench.py"
$P3 > gpurun_out/pb_c.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:psnr_kernel -s 1 -c 1 -f -o gpurun_out/prof3_psnr $P3 > gpurun_out/ncu_p3.log 2>&1
ls -la gpurun_out/prof3*.ncu-rep
