run() { python bench.py --no-cpu --no-sweep --no-e2e --steps 5 --warmup 3 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$*', '| value %.4g' % d['value'], 'conv', d['converged_frac'], 'iters', d['mean_iters'], 'exact', d['exact_match_frac'], 'frac21', round(r.get('frac_executed_21',0),3), d.get('kernel_profile_cycles_per_warp_iteration'))"; }
run
run --opt var_pipe=1
run --opt var_pipe=1 --kernel-profile
run --kernel-profile
run --opt var_pipe=1 --per 0.1 --batch 2000000
run --per 0.1 --batch 2000000
python -m pytest tests -q -m gpu -x -k "parity_configs or golden" 2>&1 | tail -2
