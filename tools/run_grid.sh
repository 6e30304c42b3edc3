export PNP_CONV_DBG=1
for g in 148 74 37; do
export PNP_CONV_GRID=$g
echo "GRID=$g"
for cfg in "64 128 0 128" "32 256 0 256"; do set -- $cfg; python tools/conv_bench.py --b 64 --s $1 --c0 $2 --c1 $3 --cout $4 --iters 2 2>&1 | grep "conv dbg" | tail -1 | cut -c1-260; done
done
