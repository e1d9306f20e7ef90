set -x
python __graft_entry__.py smoke 2>&1 | tail -6
python bench.py --steps 5 --warmup 3 > gpurun_out/f_c3.json 2> gpurun_out/f_c3.err; tail -c 300 gpurun_out/f_c3.err
python bench.py --workload C4 --steps 3 --warmup 3 > gpurun_out/f_c4.json 2> gpurun_out/f_c4.err; tail -c 300 gpurun_out/f_c4.err
python bench.py --workload C5 --steps 3 --warmup 3 --no-sweep > gpurun_out/f_c5.json 2> gpurun_out/f_c5.err; tail -c 300 gpurun_out/f_c5.err
python bench.py --workload C2 --steps 5 --warmup 3 --no-sweep > gpurun_out/f_c2.json 2> gpurun_out/f_c2.err; tail -c 300 gpurun_out/f_c2.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_ref.json 2> gpurun_out/f_ref.err; tail -c 300 gpurun_out/f_ref.err
python bench.py --steps 2 --warmup 3 --no-cpu --no-sweep > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-sweep > gpurun_out/f_ncu.log 2>&1
for f in gpurun_out/f_c*.json; do python -c "
import json,sys
d=json.load(open('$f')); print('$f', d['value'], d['e2e']['value'], d['roofline']['bound'], round(d['roofline']['frac'],3), d.get('cpu_baseline',{}).get('value'))"; done
# ncu --set full capture of the headline kernel (report stays on the box; only the text summary comes back)
timeout 700 ncu --set full --clock-control none --import-source on -k regex:bp_persistent -c 1 -o /tmp/c3 python bench.py --batch 2000000 --steps 1 --warmup 1 --no-cpu --no-sweep --no-e2e > /dev/null 2> gpurun_out/f_ncu_c3.err
python tools/ncu_summary.py /tmp/c3.ncu-rep /tmp/c3_sass.txt > gpurun_out/f_c3_ncu.txt 2>&1
head -30 gpurun_out/f_c3_ncu.txt | cut -c1-130
