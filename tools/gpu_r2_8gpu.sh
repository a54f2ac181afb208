#!/bin/bash
# 8 GPUs: the C3 strong-scaling line at N = 8 with the C4 leg (4M cells, seconds per phase)
mkdir -p gpurun_out
if [ "$(df --output=avail -BG /dev/shm | tail -1 | tr -dc 0-9)" -ge 8 ]; then export SCF_BENCH_TMP=/dev/shm; else export SCF_BENCH_TMP=/tmp; fi; echo "tmp: $SCF_BENCH_TMP"
df -h /tmp /dev/shm | tail -2
free -g | head -2
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_n8.json 2>gpurun_out/r2_bench_n8.err; echo "bench rc $?"
tail -4 gpurun_out/r2_bench_n8.err | cut -c1-300
python - <<'PY'
import json
s=open('gpurun_out/r2_bench_n8.json').read(); d=json.loads(s[s.index('{\"metric'):].splitlines()[0])
print('value', d['value'], 'ms', d['ms_per_step'], d['scaling'], d['stage_ms'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['latency_ms'])
print('parity', d['parity']); print('eig', d['eig']); print('roofline', d['roofline']['frac'], d['roofline']['ms_per_launch'])
print('C4', json.dumps(d['legs'].get('C4'), indent=1))
PY
