#!/bin/bash
# round 2, run s: tree check after the 576-thread single-issuer launch and the batched k-means seeding
mkdir -p gpurun_out
export KNN_PROBE_NO_EXACT=1
( timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | cut -c1-300
timeout 300 python tools/knn_probe.py 100000 50 11 2>&1 | tail -1 | cut -c1-200
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --legs datastore > gpurun_out/bench_s.json 2>/dev/null; echo "bench rc $?"
python - <<'PY'
import json
s=open('gpurun_out/bench_s.json').read(); d=json.loads(s[s.index('{"metric'):].splitlines()[0])
print(d['value'], d['ms_per_step'], d['stage_ms']); print('e2e', d['e2e']['ms_per_step']); print('roofline', d['roofline']['frac'], d['roofline']['ms_per_launch']); print('parity', d['parity']['ok'], d['parity']['hash_match'])
print(d['legs']['datastore_e2e'])
PY
) 2>&1 | tee gpurun_out/r2_s.log
