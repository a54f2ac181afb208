#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -q -x -k "lowess or hvg or eig or pca or chain" 2>&1 | tail -4
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'])"
