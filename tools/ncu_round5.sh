#!/bin/bash
# ncu evidence, round 2 (late): full captures of the kernels added after r02_ncu_step_b64_256.csv - the two-row first conv,
# the split-K cluster conv (B=1), the fused observation kernel and the policy step.  Each target first runs plain.
set -x
L="python tools/layer_profile.py --reps 1"
$L > gpurun_out/n5_plain1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_first -s 1 -c 1 -f -o gpurun_out/r02_first_conv $L > gpurun_out/n5_a.log 2>&1
L1="python tools/layer_profile.py --reps 1 --batch 1"
$L1 > gpurun_out/n5_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv3x3_splitk -s 20 -c 1 -f -o gpurun_out/r02_splitk $L1 > gpurun_out/n5_b.log 2>&1
R="python tools/rollout_once.py"
$R > gpurun_out/n5_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:policy_ -s 4 -c 2 -f -o gpurun_out/r02_policy $R > gpurun_out/n5_c.log 2>&1
ls -la gpurun_out/r02_*.ncu-rep
