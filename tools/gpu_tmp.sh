#!/bin/bash
# scratch: the last ad-hoc GPU call of the session (reference-store tests + ncu --set full of the two CSR -> Z kernels)
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_refstore.py -m gpu -q -x 2>&1 | tail -3
NCU_KERNELS="hvg_compact_kernel|hvg_dense_scale_kernel" NCU_COUNT=2 NCU_OUT=c2_csr_side_full NCU_TIMEOUT=150 bash tools/gpu_ncu_full.sh
