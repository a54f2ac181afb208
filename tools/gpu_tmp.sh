#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "compact or chain or mu_sigma or normalised or run_mapping" 2>&1 | tail -5
timeout 200 python tools/csr_probe.py 2>&1 | tee gpurun_out/csr_probe.log | grep -E "compact|norm_scale|dense_scale \(Z \+"
