#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "eig or chain or kmeans or symmetris" 2>&1 | tail -4
timeout 300 python tools/eig_probe.py 50 100 2>&1 | tail -3 | tee gpurun_out/r2_eig_probe.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2>gpurun_out/r2_bench_n1.err; echo "bench rc $?"
tail -3 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n1.json'))
print('value', d['value'], 'ms', d['ms_per_step'], d['stage_ms'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['latency_ms'])
print('parity', d['parity']['ok'], d['parity']['hash']); print('eig', d['eig'])
c3=d['legs']['C3']; print('C3', c3['value'], c3['ms_per_step'], c3['stage_ms'], c3['eig'], c3['parity']['hash'], c3['roofline']['frac'])
print('datastore', d['legs']['datastore_e2e'])
PY
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_c2.csv python bench.py --steps 1 --warmup 2 --no-e2e --no-cpu-baseline --no-parity --legs none --profiler-range > gpurun_out/r2_bench_ncu.log 2>&1
python tools/ncu_times.py gpurun_out/r2_launches_c2.csv > gpurun_out/r2_launches_c2.txt 2>&1; head -40 gpurun_out/r2_launches_c2.txt | cut -c1-140
