for sa in 3 2; do
export PNP_CONV_SA=$sa
echo "SA=$sa"
for cfg in "64 128 0 128" "32 256 0 256" "16 512 0 512" "32 256 512 256" "128 64 128 64"; do set -- $cfg; python tools/conv_bench.py --b 64 --s $1 --c0 $2 --c1 $3 --cout $4 --iters 5 2>&1 | tail -1; done
done
