#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "project" 2>&1 | tail -8
timeout 300 python tools/csr_probe.py 2>&1 | grep -E "project|gene_stats w"
