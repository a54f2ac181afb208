#!/bin/bash
# round 2, run n: full validation of the tree on one GPU -> pytest log, smoke, bench line (hashes), launch list
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2>gpurun_out/bench_full.err; echo "bench rc $?"
python - <<'PY'
import json
s=open('gpurun_out/bench_full.json').read(); d=json.loads(s[s.index('{"metric'):].splitlines()[0])
print(d['value'], d['ms_per_step'], d['stage_ms']); print('e2e', d['e2e']); print('roofline', d['roofline']); print('parity', d['parity']); print(d['clocks'])
for k,v in d['legs'].items(): print('leg', k, {kk: v[kk] for kk in v if kk in ('ms_per_step','value','stage_ms','parity','mark_hvgs_ms','make_graph_ms','roofline')})
PY
tail -3 gpurun_out/bench_full.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --legs none --no-parity --profiler-range > gpurun_out/bench_ncu.log 2>&1
python tools/ncu_times.py gpurun_out/launches.csv > gpurun_out/launches.txt 2>&1; head -30 gpurun_out/launches.txt | cut -c1-140
) 2>&1 | tee gpurun_out/r2_n.log
