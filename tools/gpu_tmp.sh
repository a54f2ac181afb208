#!/bin/bash
timeout 200 python -m pytest tests -m gpu -q -x -k "knn" 2>&1 | tail -3
KNN_PROBE_NO_EXACT=1 timeout 200 python tools/knn_probe.py 100000 50 11 1000000 100 21 2>&1 | tail -2
