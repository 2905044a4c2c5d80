#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_map.py tests/test_gpu_shim.py tests/test_gpu_loop_modes.py tests/test_gpu_vio_inverse.py -q -m gpu > gpurun_out/call7_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/call7_tests.log
tail -n 12 gpurun_out/call7_tests.log
INVERSE=1 TUNING=2 STAMPS=1 MODES=2,0 STEPS=30 timeout 400 python tools/loop_mode_check.py > gpurun_out/call7_timing.log 2>&1
echo "check rc=$?" >> gpurun_out/call7_timing.log
grep -E "it/s|LOOP MODES|rc=|differing|inverse" gpurun_out/call7_timing.log
timeout 900 python bench.py > gpurun_out/call7_bench.json 2> gpurun_out/call7_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/call7_bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "parity", d.get("parity_vs_oracle", {}).get("ok"))
    print("e2e_shim", json.dumps(d.get("e2e_shim"))[:900])
    print("map_update", json.dumps(d.get("map_update"))[:1500])
except Exception as e:
    print("parse", e)
PY
tail -n 5 gpurun_out/call7_bench.err
