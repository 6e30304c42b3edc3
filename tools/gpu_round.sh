#!/bin/bash
# Standard GPU round: gpu tests, smoke, bench (+ optional extra args), logs in gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/round_summary.txt
tail -n 15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/round_summary.txt
tail -n 3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/round_summary.txt
cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
