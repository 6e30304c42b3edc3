#!/bin/bash
# A/B of the general 256x256 FFT-prox cluster kernel: transposing (PNP_PROX_DIRECT=0) vs direct DSMEM gather/scatter (=1).
mkdir -p gpurun_out
PNP_PROX_DIRECT=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "prox" > gpurun_out/pytest_prox_direct.log 2>&1
echo "pytest(direct) rc=$?"; tail -n 4 gpurun_out/pytest_prox_direct.log
for d in 0 1 0 1; do
  echo "== PNP_PROX_DIRECT=$d"
  PNP_PROX_DIRECT=$d timeout 300 python tools/prox_bench.py --cases 64x256r,256x256r,1024x256r --iters 50
done | tee gpurun_out/prox_direct_ab.txt
