#!/bin/bash
# A/B of the L2 prefetches in the 256x256 FFT-prox kernels, same box, alternating.
# PNP_PROX_PREFETCH=0|1: prefetch of the blend operand rows (Yt / y0T).
mkdir -p gpurun_out
for d in ${PF_LIST:-0 1 0 1 0 1}; do
  echo "== PNP_PROX_PREFETCH=$d"
  PNP_PROX_PREFETCH=$d timeout 300 python tools/prox_bench.py --cases ${PF_CASES:-64x256c,256x256c,1024x256c,64x256r,256x256r} --iters 100 | sed 's/of 6541 GB\/s//'
done | tee gpurun_out/prox_prefetch_ab.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "prox" 2>&1 | tail -2
