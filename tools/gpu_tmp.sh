#!/bin/bash
# scratch: the last ad-hoc GPU call of the session (front-end and mapping tests after the host-side refactors)
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_datastore.py tests/test_gpu_parity.py tests/test_refstore.py -m gpu -q -x -k "datastore or mu_sigma or run_mapping or reference_written" 2>&1 | tail -3
