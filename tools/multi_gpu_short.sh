#!/bin/bash
# Short multi-GPU evidence run (gpurun --gpus 8): the in-library sharding tests on distinct devices, single-process strong
# scaling of one 10M-syndrome host batch at N = 1 / 4 / 8, torchrun weak scaling at N = 8 for C3 and C5.
nvidia-smi -L | tee gpurun_out/mg_gpus.txt
python -m pytest tests -q -m gpu -rs -k "sharding or harness or two_live" 2>&1 | tail -4 | tee gpurun_out/mg_pytest.txt
for n in 1 4 8; do
  python bench.py --single-process --gpus $n --steps 5 --warmup 3 > gpurun_out/mg_single_process_n$n.json 2> gpurun_out/mg_single_process_n$n.err || tail -5 gpurun_out/mg_single_process_n$n.err
  python -c "
import json
d=json.load(open('gpurun_out/mg_single_process_n$n.json'))
print('single-process N=$n', d['value'], d['ms_per_step'], 'counters ok', d['counters_match_single_device_run'], 'conv ok', d['converged_check'], 'nccl', d['counters_via_nccl'])"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 --no-sweep --no-cpu > gpurun_out/mg_torchrun_c3_n8.json 2> gpurun_out/mg_torchrun_c3_n8.err || tail -5 gpurun_out/mg_torchrun_c3_n8.err
python -c "
import json; d=json.load(open('gpurun_out/mg_torchrun_c3_n8.json')); print('torchrun C3 N=8', d['value'], d['e2e']['value'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --workload C5 --gpus 8 --batch 262144 --steps 1 --warmup 3 --no-sweep --no-cpu --no-e2e > gpurun_out/mg_torchrun_c5_n8.json 2> gpurun_out/mg_torchrun_c5_n8.err || tail -5 gpurun_out/mg_torchrun_c5_n8.err
python -c "
import json; d=json.load(open('gpurun_out/mg_torchrun_c5_n8.json')); print('C5 N=8', d['value'], d['roofline']['frac'])"
