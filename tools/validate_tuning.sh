#!/bin/bash
# One GPU call that re-checks and TIMES the opt-in kernel variants behind esikf_set_tuning (written at the end of round 1;
# parity passed once, profiles/gpu_tests_r01_new_paths.txt, timing still missing):   gpurun --timeout 600 -- 'bash tools/validate_tuning.sh'
# 1. the gated parity tests (bit-identity of DEFER / VIO_FAST, association + tolerance checks of DEAL_POINTS, the
#    inverse-compositional VIO variant against the oracle);
# 2. resident LIO + VIO step time of the default and of every flag combination on BASELINE config 2, with phase stamps.
mkdir -p gpurun_out
ESIKF_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_loop_modes.py tests/test_gpu_vio_inverse.py tests/test_gpu_z_new_entry_points.py -q -m gpu > gpurun_out/tuning_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/tuning_tests.log
TUNING=1,2,4,6,7 STAMPS=1 MODES=2 STEPS=30 timeout 240 python tools/loop_mode_check.py > gpurun_out/tuning_timing.log 2>&1
echo "check rc=$?" >> gpurun_out/tuning_timing.log
tail -n 15 gpurun_out/tuning_tests.log
grep -E "it/s|LOOP MODES|rc=" gpurun_out/tuning_timing.log
