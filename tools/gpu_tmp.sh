#!/bin/bash
mkdir -p gpurun_out
( SCF_EIG_DEBUG=1 timeout 300 python tools/eig_probe.py 100 2>&1 | grep tridiag | head -2
timeout 300 python tools/eig_probe.py 50 100 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "eig or jacobi or tridiag or chain or pca" 2>&1 | tail -4 | cut -c1-250 ) | tee gpurun_out/r2_tmp.log
