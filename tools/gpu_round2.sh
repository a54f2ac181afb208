#!/bin/bash
# gpurun call 2: new kernels (fix-scan repair, tcgen05 Gram) + ncu full capture of the kNN tensor-core kernel
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python tools/knn_probe.py > gpurun_out/knn_probe.log 2>&1; cat gpurun_out/knn_probe.log
for f in 4 8; do SCF_KNN_FLAGS=$f timeout 120 python tools/knn_probe.py 100000 50 11 > gpurun_out/knn_probe_flags$f.log 2>&1; cat gpurun_out/knn_probe_flags$f.log; done
timeout 600 python bench.py --steps 3 --warmup 3 --knn-method 1 --gram-mode 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_m1g3.json 2> gpurun_out/bench_m1g3.err; cat gpurun_out/bench_m1g3.json; tail -3 gpurun_out/bench_m1g3.err
timeout 600 python bench.py --steps 3 --warmup 3 --knn-method 1 --gram-mode 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_m1g1.json 2> gpurun_out/bench_m1g1.err; cat gpurun_out/bench_m1g1.json; tail -3 gpurun_out/bench_m1g1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel -c 1 -o gpurun_out/prof_knn_tc python tools/knn_probe.py 100000 50 11 > gpurun_out/ncu_knn.log 2>&1; tail -3 gpurun_out/ncu_knn.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_tc_kernel -c 1 -o gpurun_out/prof_gram_tc python bench.py --steps 1 --warmup 0 --knn-method 1 --gram-mode 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_gram.log 2>&1; tail -3 gpurun_out/ncu_gram.log
ls -la gpurun_out
