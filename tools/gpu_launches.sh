#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "lowess or hvg" 2>&1 | tail -3
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --profiler-range > gpurun_out/bench_ncu.log 2>&1
python tools/ncu_times.py gpurun_out/launches.csv > gpurun_out/launches.txt 2>&1; head -70 gpurun_out/launches.txt | cut -c1-140
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'])"
