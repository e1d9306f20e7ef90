# Final 1-GPU evidence run of a round: test-suite, smoke, the default bench line, launch list, ncu captures.
set -x
python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee gpurun_out/f_pytest.txt
python __graft_entry__.py smoke 2>&1 | tail -4 | tee gpurun_out/f_smoke.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/f_default.json 2> gpurun_out/f_default.err; tail -c 300 gpurun_out/f_default.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_ref.json 2> gpurun_out/f_ref.err; tail -c 300 gpurun_out/f_ref.err
python bench.py --steps 2 --warmup 3 --no-cpu --no-sweep > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-sweep > gpurun_out/f_ncu.log 2>&1
python -c "
import json
d=json.load(open('gpurun_out/f_default.json')); r=d['roofline']
print('C3', d['value'], d['e2e']['value'], r['frac'], r.get('frac_executed_21'), r['timing']['kernel_share_of_step'], r['traffic'])
for k in ('config_C4_hgp1600_10M','config_C5_gallager100k_1M'):
    x=d[k]; print(k, x['value'], x['roofline']['frac'], x['roofline']['traffic'], x['roofline']['algorithmic_bytes_per_launch'])
"
bash tools/ncu_capture.sh r2 ${1:-unknown} ${FINAL_NCU:-C3 C4 C5}
