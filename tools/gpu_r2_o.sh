#!/bin/bash
# round 2, run o: ncu --set full of the C3-shard kNN kernel (knn_tc_kernel<32, 2, false>, 125k x 1M, D 100, k 21), FP64 yardsticks,
# k-means timing through the DataStore leg
mkdir -p gpurun_out
export KNN_PROBE_NO_EXACT=1
( timeout 300 python tools/fp64_probe.py 2>&1 | tail -5
KNN_PROBE_NQ=125000 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel -s 2 -c 1 -o gpurun_out/r2_knn_c3 python tools/knn_probe.py 1000000 100 21 > gpurun_out/r2_ncu_knn_c3.log 2>&1
tail -2 gpurun_out/r2_ncu_knn_c3.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "kmeans" 2>&1 | tail -2
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --legs datastore --no-parity 2>/dev/null | python -c "
import json,sys
s=sys.stdin.read(); d=json.loads(s[s.index('{\"metric'):].splitlines()[0]); print(d['legs']['datastore_e2e'])"
) 2>&1 | tee gpurun_out/r2_o.log
