#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_refstore.py -m gpu -q -x 2>&1 | tail -15
timeout 120 python tools/dense_probe.py 8000 30000 2>&1 | tail -3
