#!/bin/bash
# round 2, run z2: racecheck of the tridiagonal solver's kernels and the Cholesky kernel (small matrices), full log kept
mkdir -p gpurun_out
timeout 330 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_sym_eig_tridiag_equals_eigh or test_eig_topk_small" > gpurun_out/racecheck_eig.log 2>&1
echo "rc $?"
grep -c "Race reported" gpurun_out/racecheck_eig.log
grep "Race reported\|and Write\|and Read\|RACECHECK SUMMARY\|passed\|failed" gpurun_out/racecheck_eig.log | sed 's/+0x[0-9a-f]*//; s/(const double.*)//' | sort | uniq -c | sort -rn | head -20 | cut -c1-200
