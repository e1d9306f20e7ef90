# quick C3 / C2 check after a kernel change: parity subset, then throughput at the bench operating points
python -m pytest tests -q -m gpu -x -k "parity_configs or ragged or golden or filter or semantic or tiny or irregular or sharding or two_live or direct_bit" 2>&1 | tail -2
run() { python bench.py --no-cpu --no-sweep --steps 5 --warmup 3 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$*', '| value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], 'frac21', round(r.get('frac_executed_21', 0),3), 'conv', d['converged_frac'], d['mean_iters'])"; }
run
run --opt dynamic_queue=0
run --per 0.1 --batch 2000000
run --per 0.1 --batch 2000000 --opt dynamic_queue=0
run --workload C2
run --workload C2 --opt dynamic_queue=0
run --batch 1000000
run --batch 1000000 --opt dynamic_queue=0
