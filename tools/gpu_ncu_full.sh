#!/bin/bash
# ncu --set full of the four heaviest kernels of one C2 step (one launch each)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:"knn_tc_kernel|project_tc_kernel|gene_stats_win_kernel|gram_tc_kernel|hvg_compact_kernel" -c 6 -o gpurun_out/c2_step_full -f \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --profiler-range > gpurun_out/ncu_step_full.log 2>&1
tail -2 gpurun_out/ncu_step_full.log; ls -la gpurun_out/c2_step_full.ncu-rep
