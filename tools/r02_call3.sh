#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02_call3_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_call3_tests.log
tail -n 25 gpurun_out/r02_call3_tests.log
TUNING=1 STAMPS=1 MODES=2,0 STEPS=30 timeout 300 python tools/loop_mode_check.py > gpurun_out/r02_call3_timing.log 2>&1
echo "check rc=$?" >> gpurun_out/r02_call3_timing.log
grep -E "it/s|LOOP MODES|rc=|differing|^LIO|^VIO|per-CTA" gpurun_out/r02_call3_timing.log
timeout 600 python bench.py > gpurun_out/r02_call3_bench.json 2> gpurun_out/r02_call3_bench.err
echo "bench rc=$?" >> gpurun_out/r02_call3_bench.err
tail -n 5 gpurun_out/r02_call3_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_call3_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d.get('e2e_shim'), d['roofline']['frac'], d['parity_vs_oracle'])
except Exception as e: print('bench parse', e)
PY
timeout 900 bash tools/capture_ncu.sh > gpurun_out/r02_call3_ncu.log 2>&1
tail -n 6 gpurun_out/r02_call3_ncu.log
