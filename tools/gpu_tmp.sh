#!/bin/bash
timeout 120 python tools/jacobi_probe.py 2>&1 | tail -2
bash tools/gpu_quick.sh "jacobi or eig or pca or chain or subset"
