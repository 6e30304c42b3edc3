export PNP_CONV_KWS=0
for b in 8 16 32 64 128; do python tools/conv_bench.py --b $b --s 256 --c0 32 --c1 0 --cout 32 --iters 10 2>&1 | tail -1; done
for b in 8 16 32 64 128; do python tools/conv_bench.py --b $b --s 64 --c0 128 --c1 0 --cout 128 --iters 10 2>&1 | tail -1; done
export PNP_CONV_KWS=1
for b in 8 16 32 64 128; do python tools/conv_bench.py --b $b --s 256 --c0 32 --c1 0 --cout 32 --iters 10 2>&1 | tail -1; done
