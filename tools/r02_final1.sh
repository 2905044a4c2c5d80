#!/bin/bash
# Final single-GPU validation of the round: full GPU suite, smoke, bench (both arms), variant timings, ncu evidence.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/final_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/final_gpu_tests.log
tail -n 8 gpurun_out/final_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/final_smoke.log
timeout 900 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "reference arm rc=$?"
python - <<'PY'
import json
for f in ("final_bench_n1", "final_bench_reference"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 1), "e2e", d.get("e2e", {}).get("value"), "frac", d.get("roofline", {}).get("frac"), "parity", d.get("parity_vs_oracle", {}).get("ok"),
              "shim", (d.get("e2e_shim") or {}).get("value"), "map", (d.get("map_update") or {}).get("device_ms_per_tick"))
    except Exception as e:
        print(f, "parse", e)
PY
INVERSE=1 TUNING=1,2 STAMPS=1 MODES=2,0 STEPS=30 timeout 400 python tools/loop_mode_check.py > gpurun_out/final_timing.log 2>&1
echo "check rc=$?" >> gpurun_out/final_timing.log
grep -E "it/s|LOOP MODES|rc=|differing|inverse" gpurun_out/final_timing.log
bash tools/capture_ncu.sh > gpurun_out/final_ncu.log 2>&1
ESIKF_MAP=1 STEPS=2 ncu --clock-control none --set full --import-source on -k regex:map_replay_kernel -s 1 -c 1 -f -o gpurun_out/prof_map_replay python tools/profile_driver.py > gpurun_out/prof_d.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
