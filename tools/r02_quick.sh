#!/bin/bash
# quick correctness + timing check of the current build (one B200); optional A/B library in /tmp
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lio.py tests/test_gpu_loop_modes.py tests/test_gpu_vio.py tests/test_gpu_z_new_entry_points.py -q -m gpu -x > gpurun_out/quick_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/quick_tests.log
tail -n 6 gpurun_out/quick_tests.log
TUNING=1 STAMPS=1 MODES=2,0 STEPS=30 timeout 300 python tools/loop_mode_check.py > gpurun_out/quick_timing.log 2>&1
echo "check rc=$?" >> gpurun_out/quick_timing.log
grep -E "it/s|LOOP MODES|rc=|differing|^LIO|^VIO [0-3] |per-CTA" gpurun_out/quick_timing.log
if [ -f ab/libesikf_inl.so ]; then
  ESIKF_LIB=$PWD/ab/libesikf_inl.so STAMPS=1 MODES=2 STEPS=30 timeout 300 python tools/loop_mode_check.py > gpurun_out/quick_timing_inl.log 2>&1
  echo "== inlined cold paths:"; grep -E "it/s|^LIO|per-CTA" gpurun_out/quick_timing_inl.log
fi
