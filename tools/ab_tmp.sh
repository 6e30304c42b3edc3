for rep in 1 2; do for sa in 3 2; do echo -n "rep $rep SA=$sa: "; PNP_CONV_SA=$sa timeout 300 python tools/layer_profile.py --reps 10 2>&1 | tail -1; done; done
