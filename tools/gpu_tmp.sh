#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "compact or chain_tensor" 2>&1 | tail -3
timeout 200 python tools/csr_probe.py 2>&1 | tee gpurun_out/csr_probe.log | grep -E "dense_scale"
