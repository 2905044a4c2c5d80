#!/bin/bash
# New kernels (compact records, resident invariants, compact partials): GPU suite, sanitizer pass on the smoke frame, timing.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02_call2_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_call2_tests.log
tail -n 30 gpurun_out/r02_call2_tests.log
TUNING=1 STAMPS=1 MODES=2,0 STEPS=30 timeout 300 python tools/loop_mode_check.py > gpurun_out/r02_call2_timing.log 2>&1
echo "check rc=$?" >> gpurun_out/r02_call2_timing.log
grep -E "it/s|LOOP MODES|rc=|differing" gpurun_out/r02_call2_timing.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_call2_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02_call2_memcheck.log
tail -n 8 gpurun_out/r02_call2_memcheck.log
