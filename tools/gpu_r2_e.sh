#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest_gpu.log; tail -15 gpurun_out/r2_pytest_gpu.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python tools/eig_probe.py 50 100 2>&1 | tail -4 | tee gpurun_out/r2_eig_probe.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2>gpurun_out/r2_bench_n1.err; echo "bench rc $?"
tail -5 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2_bench_n1.json'))
    print('value', d['value'], 'ms', d['ms_per_step'], d['stage_ms'])
    print('e2e', d['e2e'])
    print('roofline', {k:v for k,v in d['roofline'].items() if k in ('frac','achieved','ms_per_launch','entry_point_ms','hbm_side','peaks_measured_in_run')})
    print('parity', d['parity']); print('eig', d['eig']); print('clocks', d['clocks']); print('cpu', d['cpu_baseline'])
    for k,v in d['legs'].items():
        print('LEG', k, json.dumps(v)[:1500])
except Exception as e:
    print('bench parse failed', e)
PY
