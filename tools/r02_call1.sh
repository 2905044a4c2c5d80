#!/bin/bash
# Round-2 re-baseline of the round-1 build on one B200: full GPU suite with the gated tests enabled, timing of the opt-in
# tuning variants (never timed in round 1), and the default bench line.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_call1_gpu.txt 2>&1
ESIKF_EXPERIMENTAL=1 timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/r02_call1_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_call1_tests.log
TUNING=1,2,4,6,7 STAMPS=1 MODES=2,1 STEPS=30 timeout 300 python tools/loop_mode_check.py > gpurun_out/r02_call1_tuning.log 2>&1
echo "check rc=$?" >> gpurun_out/r02_call1_tuning.log
timeout 300 python bench.py > gpurun_out/r02_call1_bench.json 2> gpurun_out/r02_call1_bench.err
echo "bench rc=$?" >> gpurun_out/r02_call1_bench.err
tail -n 5 gpurun_out/r02_call1_tests.log
grep -E "it/s|LOOP MODES|rc=" gpurun_out/r02_call1_tuning.log
head -c 600 gpurun_out/r02_call1_bench.json
