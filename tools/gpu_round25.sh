#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "gene_stats or hvg or row_sums or smoke" 2>&1 | tail -3
timeout 300 python tools/csr_probe.py 2>&1 | grep -E "gene_|mark_hvgs|row_sums"
