timeout 150 python -m pytest tests -m gpu -q -x -k "knn" 2>&1 | tail -3
timeout 300 python tools/c3_knn_probe.py 2>&1 | tail -3
