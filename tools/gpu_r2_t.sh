#!/bin/bash
# round 2, run t: eigenpairs of the tridiagonal matrix over the whole GPU, cooperative split-K reduction of the FP64 GEMM
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "eig or tridiag or jacobi or chain or pca" 2>&1 | tail -3 | cut -c1-300
export EIG_PROBE_COV=tools/build/c3_cov.npy EIG_PROBE_N=1000000
timeout 300 python tools/eig_probe.py 100 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/eig_c3_launches.csv python tools/eig_probe.py 100 > gpurun_out/eig_c3_ncu.log 2>&1
python tools/ncu_times.py gpurun_out/eig_c3_launches.csv > gpurun_out/eig_c3_launches.txt 2>&1; grep -v "sytrd\|laed\|cutlass\|syherk\|larft\|transpose\|stedc" gpurun_out/eig_c3_launches.txt | head -12 | cut -c1-120
unset EIG_PROBE_COV EIG_PROBE_N
timeout 300 python tools/eig_probe.py 50 2>&1 | tail -1
) 2>&1 | tee gpurun_out/r2_t.log
