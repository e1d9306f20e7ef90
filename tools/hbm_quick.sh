python -m pytest tests -q -m gpu -x -k "c5 or parity_configs or staging or irregular or decision_fields or minsum or fast32 or tiny or ragged or semantic or bposd_matches or auto_family" 2>&1 | tail -2
run() { python bench.py --no-cpu --no-sweep --no-e2e --steps 2 --warmup 3 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['kernel']
print('$*', '| value %.4g frac %.3f' % (d['value'], d['roofline']['frac']), 'warps', k['threads_per_cta']//32, 'pd', k['prefetch_distance'], 'mode', k['kernel_mode'])"; }
run --workload C4 --batch 1000000
run --workload C4 --batch 1000000 --opt dynamic_queue=0
run --workload C5 --batch 65536
run --workload C5 --batch 65536 --opt dynamic_queue=0
run --workload C5 --batch 262144 --tile 262144 --steps 1
run --workload C1 --batch 1000000
run --workload C1 --batch 1000000 --opt dynamic_queue=0
