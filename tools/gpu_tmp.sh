#!/bin/bash
for v in 5 6 7 8 9; do
  echo "== SCF_GS_VARIANT=$v"
  SCF_GS_VARIANT=$v timeout 120 python tools/csr_probe.py 2>&1 | grep -E "gene_stats w|gene_ncells w"
done
SCF_GS_VARIANT=6 timeout 120 python -m pytest tests -m gpu -q -x -k "gene_stats or hvg" 2>&1 | tail -2
