# run one bench configuration against several experiment builds of the library (LDPCB200_LIB override)
run() { python bench.py --no-cpu --no-sweep --no-e2e --steps 2 --warmup 3 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['kernel']
print('$LDPCB200_LIB $*', '| value %.4g frac %.3f' % (d['value'], d['roofline']['frac']), 'warps', k['threads_per_cta']//32, 'pd', k['prefetch_distance'])"; }
for lib in "" $(ls ldpcdecoders.jl_b200/lib/exp_*.so); do
  export LDPCB200_LIB=$lib
  [ -n "$lib" ] && export LDPCB200_LIB=$PWD/$lib
  run --workload C5 --batch 65536
  run --workload C5 --batch 65536 --warps 8
done
