# quick C3 / C2 check after a kernel change: parity subset, then throughput at the bench operating points
python -m pytest tests -q -m gpu -x -k "parity_configs or ragged or golden or filter or semantic or tiny or irregular or sharding or two_live" 2>&1 | tail -2
run() { python bench.py --no-cpu --no-sweep --no-e2e --steps 5 --warmup 3 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$*', '| value %.4g' % d['value'], 'frac', round(r['frac'],3), {k:round(v,3) for k,v in r.items() if k.startswith('frac_ex') or k.startswith('whole')}, 'kernel share', round(r['timing']['kernel_share_of_step'],3), d.get('kernel_profile_cycles_per_warp_iteration'))"; }
run
run --kernel-profile
run --per 0.1 --batch 2000000
run --workload C2
run --opt first_iteration_filter=0
