for rep in 1 2; do for m in 64 32; do echo -n "rep $rep KWS_MIN=$m: "; PNP_CONV_KWS_MIN=$m python bench.py --no-cpu --no-variants --steps 100 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'])"; done; done
