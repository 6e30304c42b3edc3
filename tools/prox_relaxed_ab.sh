#!/bin/bash
# A/B of the relaxed execution-only cluster barriers in the 256x256 cluster FFT-prox kernel, same box, alternating.
mkdir -p gpurun_out
PNP_PROX_RELAXED=1 timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_env.py -m gpu -q -x -k "prox or engine or step" 2>&1 | tail -2
for d in ${RX_LIST:-0 1 0 1 0 1}; do
  echo "== PNP_PROX_RELAXED=$d"
  PNP_PROX_RELAXED=$d timeout 300 python tools/prox_bench.py --cases 64x256r,256x256r,1024x256r --iters 100 | sed 's/of 6541 GB\/s//'
done | tee gpurun_out/prox_relaxed_ab.txt
