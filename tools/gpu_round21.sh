#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/knn_probe.py 5000 25 11 100000 50 11 200000 100 21 1000000 100 21 2>&1 | tee gpurun_out/knn_probe.log
