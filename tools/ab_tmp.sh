timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for pr in 0 1 auto; do echo -n "PAIR=$pr: "; if [ $pr = auto ]; then unset PNP_CONV_PAIR; else export PNP_CONV_PAIR=$pr; fi; timeout 300 python tools/layer_profile.py --reps 10 > gpurun_out/layers_pair$pr.txt 2>&1; tail -1 gpurun_out/layers_pair$pr.txt; done
unset PNP_CONV_PAIR
python bench.py --no-cpu --no-variants --steps 100 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('auto sustained', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
