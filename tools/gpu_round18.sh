#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
VARIANTS="4" SHAPES="100000 50 11" bash tools/gpu_round13.sh
timeout 300 python tools/csr_probe.py 2>&1 | grep -E "dense_scale|mark_hvgs|compact"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'])"
