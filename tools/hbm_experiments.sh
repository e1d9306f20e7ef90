# Ring-depth / CTA-width sweep of the HBM-resident modes (C5: mode 2, C4: mode 1).  value = syndromes/s, frac = HBM roofline fraction.
run() { python bench.py --no-cpu --no-sweep --no-e2e --steps 2 --warmup 3 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['kernel']
print('$*', '| value %.4g frac %.3f' % (d['value'], d['roofline']['frac']), 'warps', k['threads_per_cta']//32, 'ctas', k['ctas_per_sm'], 'pd', k['prefetch_distance'], 'mode', k['kernel_mode'])"; }
for pd in 3 4 5 6; do for w in 12 10 8; do run --workload C5 --batch 65536 --prefetch $pd --warps $w; done; done
run --workload C5 --batch 65536 --prefetch 6 --warps 16 --max-ctas 1
for pd in 3 4 6; do for w in 12 8; do run --workload C4 --batch 1000000 --prefetch $pd --warps $w; done; done
