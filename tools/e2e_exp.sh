python -m pytest tests -q -m gpu -x -k "pinned_and_pageable or element_formats or bit_formats or sharding or filter or ragged or bposd_large" 2>&1 | tail -2
run() { python bench.py --no-cpu --no-sweep --steps 5 --warmup 3 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*', '| device %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], d['e2e']['converged_check'], d['e2e']['gpu_launches'])"; }
run --opt overlap_chunks=0
run
run --opt chunk=1666688
run --opt chunk=1250016
run --opt chunk=833344
run --opt chunk=625024
run --workload C2
run --workload C2 --opt overlap_chunks=0
