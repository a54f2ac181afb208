#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gene_stats_win_kernel -c 1 -o gpurun_out/gene_stats_win -f python tools/csr_probe.py > gpurun_out/ncu_gs.log 2>&1
tail -2 gpurun_out/ncu_gs.log
