#!/bin/bash
# A/B of the first conv (2->32): one-pixel-per-thread kernel (PNP_FIRST_QUAD=0) vs lane-quad kernel (default).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_env.py -m gpu -q -x -k "unet or env or engine or step or traj" > gpurun_out/pytest_first_quad.log 2>&1
echo "pytest(quad) rc=$?"; tail -n 4 gpurun_out/pytest_first_quad.log
for q in 0 1; do
  echo "== PNP_FIRST_QUAD=$q"
  PNP_FIRST_QUAD=$q timeout 300 python tools/layer_profile.py | grep -E "inc.conv|total"
done | tee gpurun_out/first_quad_ab.txt
for q in 0 1 0 1; do
  echo "== bench PNP_FIRST_QUAD=$q"
  PNP_FIRST_QUAD=$q timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-variants 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
done | tee -a gpurun_out/first_quad_ab.txt
