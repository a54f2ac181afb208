#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py --workload C3 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
python -c "
import json
d=json.load(open('gpurun_out/bench_c3.json')); print(d['value'], d['ms_per_step'], d['stage_ms']); print(d['roofline']['achieved'], d['roofline']['frac'], d['clocks'])"
tail -3 gpurun_out/bench_c3.err
