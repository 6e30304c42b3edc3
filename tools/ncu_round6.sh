#!/bin/bash
# ncu evidence, round 2 (session 2): per-kernel DRAM bytes / durations of one step with the current kernels (refreshes
# profiles/r02_ncu_step_b64_256.csv, which bench.py parses for roofline.traffic), and full captures of the kernels changed in
# this session: the pipelined general-mask cluster prox (v7), the row-only prox with its batched epilogue loads, the
# dense-DFT any-size kernel.  Each target first runs plain.
set -x
S="python tools/ncu_step.py 64 256"
$S > gpurun_out/n6_plain1.log 2>&1 && \
ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    --csv --log-file gpurun_out/r02_ncu_step_b64_256_s2.csv $S > gpurun_out/n6_a.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:fftprox_cl_kernel -c 1 -f \
    -o gpurun_out/r02_prox_cl_v7 $S > gpurun_out/n6_b.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:fftprox_rows256 -c 1 -f \
    -o gpurun_out/r02_prox_rows256_s2 $S > gpurun_out/n6_c.log 2>&1
P="python tools/prox_bench.py --iters 2 --cases 64x130r"
$P > gpurun_out/n6_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dft_any -s 6 -c 3 -f -o gpurun_out/r02_prox_any_130 $P > gpurun_out/n6_d.log 2>&1
ls -la gpurun_out/r02_*s2* gpurun_out/r02_prox_cl_v7* gpurun_out/r02_prox_any_130*
