#!/bin/bash
mkdir -p gpurun_out
timeout 60 tools/bin/tma_probe > gpurun_out/tma_probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/tma_probe.log
cat gpurun_out/tma_probe.log
timeout 900 python -m pytest tests/test_gpu_map.py -q -m gpu -x > gpurun_out/call5_map_tests.log 2>&1
echo "map pytest rc=$?"
tail -n 30 gpurun_out/call5_map_tests.log
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_loop_modes.py -q -m gpu -x -k tma > gpurun_out/tma_sanitizer.log 2>&1
echo "sanitizer rc=$?"
grep -E "Illegal|Invalid|at 0x|in esikf|error|Error|passed|failed" gpurun_out/tma_sanitizer.log | head -20
