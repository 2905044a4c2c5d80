#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/tma_probe.log
for v in 0 1 2 3 4 5 6 7; do timeout 60 tools/bin/tma_probe $v >> gpurun_out/tma_probe.log 2>&1; done
cat gpurun_out/tma_probe.log
timeout 900 python -m pytest tests/test_gpu_map.py tests/test_gpu_shim.py -q -m gpu > gpurun_out/call6_map_tests.log 2>&1
echo "map+shim pytest rc=$?"
tail -n 25 gpurun_out/call6_map_tests.log
