#!/bin/bash
# quick check after a change: selected tests (pytest -k "$1") under a short timeout, then the stage times of a C2 step
timeout 180 python -m pytest tests -m gpu -q -x -k "$1" 2>&1 | tail -4
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'])"
