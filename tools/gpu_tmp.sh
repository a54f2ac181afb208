#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "eig or chain or pca or compact" 2>&1 | tail -4
timeout 120 python tools/eig_probe.py 2>&1 | tee gpurun_out/eig_probe.log | grep -E "qr|cholqr2|cheb|eig_topk|eigh f64 164"
timeout 200 python tools/csr_probe.py 2>&1 | tee gpurun_out/csr_probe.log | grep -E "compact \(variant None|dense_scale"
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'])"
