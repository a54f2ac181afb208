#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/c5_mapping_probe.py 20000 10000 2>&1 | tail -5
timeout 420 python tools/c5_mapping_probe.py > gpurun_out/c5_mapping.json 2> gpurun_out/c5_mapping.err; tail -3 gpurun_out/c5_mapping.err; cat gpurun_out/c5_mapping.json
