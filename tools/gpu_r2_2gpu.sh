#!/bin/bash
# 2 GPUs: sharded == single-GPU (NCCL), then the C3 strong-scaling bench line at N = 2
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -x > gpurun_out/r2_multigpu_2gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_multigpu_2gpu.log; tail -5 gpurun_out/r2_multigpu_2gpu.log | cut -c1-300
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.json 2>gpurun_out/r2_bench_n2.err; echo "bench rc $?"
tail -4 gpurun_out/r2_bench_n2.err | cut -c1-300
python - <<'PY'
import json
s=open('gpurun_out/r2_bench_n2.json').read(); d=json.loads(s[s.index('{\"metric'):].splitlines()[0])
print('value', d['value'], 'ms', d['ms_per_step'], d['scaling'], d['stage_ms'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['latency_ms'])
print('parity', d['parity']); print('eig', d['eig']); print('roofline', d['roofline']['frac'], d['roofline']['ms_per_launch'])
PY
