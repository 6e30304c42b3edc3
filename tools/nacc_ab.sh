#!/bin/bash
# A/B helper: conv tests + per-layer table + short bench for the current build.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "conv or unet" > gpurun_out/pytest_nacc.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/pytest_nacc.log
timeout 300 python tools/layer_profile.py | tee gpurun_out/layers_nacc4.txt | grep -E "32 @256|64 @128|total"
for i in 1 2; do timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-variants 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'])"; done
