# Phase-timing experiments of the shared-memory kernel on C3 (per 0.03): SM cycles per warp-iteration by phase for
# the shipped shape, one CTA per SM, and the two-teams form.  Diagnostics only (clock reads inside the kernel).
run() { python bench.py --no-cpu --no-sweep --no-e2e --steps 3 --warmup 3 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*', '| value %.4g' % d['value'], d.get('kernel_profile_cycles_per_warp_iteration'), d['kernel']['ctas_per_sm'], d['kernel']['threads_per_cta'])"; }
run
run --kernel-profile
run --kernel-profile --max-ctas 1
run --max-ctas 1
run --kernel-profile --opt dual=1
run --opt dual=1
run --opt first_iteration_filter=0
run --per 0.1 --batch 2000000
run --per 0.1 --batch 2000000 --kernel-profile
run --per 0.1 --batch 2000000 --kernel-profile --max-ctas 1
