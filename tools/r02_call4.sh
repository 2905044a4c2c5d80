#!/bin/bash
# one B200: the new VIO variants (inverse-compositional loop in the persistent kernel, TMA tap footprints), the shim e2e fix
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vio_inverse.py tests/test_gpu_loop_modes.py tests/test_gpu_vio.py tests/test_gpu_shim.py -q -m gpu > gpurun_out/call4_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/call4_tests.log
tail -n 8 gpurun_out/call4_tests.log
INVERSE=1 TUNING=2 STAMPS=1 MODES=2,0 STEPS=30 timeout 400 python tools/loop_mode_check.py > gpurun_out/call4_timing.log 2>&1
echo "check rc=$?" >> gpurun_out/call4_timing.log
grep -E "it/s|LOOP MODES|rc=|differing|inverse|^VIO [0-9] " gpurun_out/call4_timing.log
timeout 600 python bench.py > gpurun_out/call4_bench.json 2> gpurun_out/call4_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/call4_bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"], "e2e_shim", d.get("e2e_shim"), "frac", d["roofline"]["frac"], "parity", d.get("parity", {}).get("ok"))
except Exception as e:
    print("parse", e)
PY
tail -n 5 gpurun_out/call4_bench.err
