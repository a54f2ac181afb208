#!/bin/bash
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms']); r=d['roofline']; print({k:r[k] for k in ('achieved','peak','frac','ms_per_launch','entry_point_ms','entry_point_achieved','tmem_readout_floor_ms')})"
