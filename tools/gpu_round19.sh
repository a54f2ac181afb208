#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python tools/knn_probe.py 5000 25 11 100000 50 11 200000 100 21 1000000 100 21 2>&1 | tee gpurun_out/knn_probe.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'])"
