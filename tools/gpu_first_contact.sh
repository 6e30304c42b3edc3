#!/bin/bash
# First-contact run on the B200 box: every group in its own process under a timeout, logs to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/fc_gpu.txt 2>&1
run() { # name, timeout, cmd...
  local name=$1; local to=$2; shift 2
  echo "=== $name ===" | tee -a gpurun_out/fc_summary.txt
  timeout $to "$@" > gpurun_out/fc_$name.log 2>&1
  echo "exit=$? $(tail -n 3 gpurun_out/fc_$name.log | tr '\n' ' ')" | tee -a gpurun_out/fc_summary.txt
}
: > gpurun_out/fc_summary.txt
run psnr 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "psnr"
run fft 300 python -m pytest tests/test_gpu_kernels.py -q -k "fft2c"
run prox 300 python -m pytest tests/test_gpu_kernels.py -q -k "prox"
PNP_DESC_MODE=0 run conv_mode0 300 python -m pytest tests/test_gpu_kernels.py -q -k "conv3x3"
PNP_DESC_MODE=1 run conv_mode1 300 python -m pytest tests/test_gpu_kernels.py -q -k "conv3x3"
run unet 300 python -m pytest tests/test_gpu_kernels.py -q -k "unet"
cat gpurun_out/fc_summary.txt
