#!/bin/bash
# Multi-GPU evidence run (gpurun --gpus N): the GPU test-suite with the in-library sharding tests on distinct devices,
# the single-process strong-scaling line (one ldpcb200 handle over 1/2/4/8 devices, one 10M-syndrome host batch) and
# the torchrun weak-scaling lines.  Outputs under gpurun_out/mg_*.
N=${1:-8}
nvidia-smi -L | tee gpurun_out/mg_gpus.txt
python -m pytest tests -q -m gpu -rs ${MG_PYTEST_ARGS:-} 2>&1 | tail -8 | tee gpurun_out/mg_pytest.txt
for n in 1 2 4 8; do
  [ $n -le $N ] || continue
  python bench.py --single-process --gpus $n --steps 5 --warmup 3 > gpurun_out/mg_single_process_n$n.json 2> gpurun_out/mg_single_process_n$n.err || tail -5 gpurun_out/mg_single_process_n$n.err
  python - <<PY
import json
d=json.load(open("gpurun_out/mg_single_process_n$n.json"))
print("single-process N=$n", d["value"], d["ms_per_step"], "counters ok", d["counters_match_single_device_run"], "conv ok", d["converged_check"], "nccl", d["counters_via_nccl"])
PY
done
for n in 2 4 8; do
  [ $n -le $N ] || continue
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 --no-sweep --no-cpu > gpurun_out/mg_torchrun_c3_n$n.json 2> gpurun_out/mg_torchrun_c3_n$n.err || tail -5 gpurun_out/mg_torchrun_c3_n$n.err
  python -c "
import json; d=json.load(open('gpurun_out/mg_torchrun_c3_n$n.json')); print('torchrun C3 N=$n', d['value'], d['e2e']['value'])"
done
# BASELINE config 5 (n = 100k, 1M streamed) across the GPUs: weak scaling, 1M syndromes per rank would take minutes -> 262144 per rank
for n in ${MG_C5_NS:-1 $N}; do
  if [ $n -eq 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512"; fi
  $L bench.py --workload C5 --gpus $n --batch 262144 --steps 1 --warmup 3 --no-sweep --no-cpu --no-e2e > gpurun_out/mg_torchrun_c5_n$n.json 2> gpurun_out/mg_torchrun_c5_n$n.err || tail -5 gpurun_out/mg_torchrun_c5_n$n.err
  python -c "
import json; d=json.load(open('gpurun_out/mg_torchrun_c5_n$n.json')); print('C5 N=$n', d['value'], d['roofline']['frac'])"
done
