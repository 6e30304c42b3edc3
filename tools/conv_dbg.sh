#!/bin/bash
# Stall breakdown + CTA timeline of representative conv layers (PNP_CONV_DBG instrumentation), B=64 at 256x256 pyramid sizes.
export PNP_CONV_DBG=1
export PNP_CONV_KWS=0
run() { python tools/conv_bench.py --b 64 --s $1 --c0 $2 --c1 $3 --cout $4 --iters 3 2>&1 | grep "conv dbg\|timeline\|TFLOP" | tail -3; }
run 256 32 0 32
run 128 64 0 64
run 64 128 0 128
run 32 256 0 256
