#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "kmeans" 2>&1 | tail -30 | cut -c1-250
