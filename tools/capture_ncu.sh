#!/bin/bash
# ncu evidence for profiles/ in one GPU call (never a bench value: everything here runs under the profiler):
#   gpurun --timeout 900 -- 'bash tools/capture_ncu.sh'
# 1. launch list of the bench command (per-launch gpu__time_duration, cold cache, serialised) -> gpurun_out/launches.csv
# 2. --set full captures of the persistent update kernels and of the per-iteration LIO residual kernel (-lineinfo is on,
#    so `ncu -i X.ncu-rep --page source` maps to the .cu files); read them here with tools/ncu_extract.py raw|stall.
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-shim > gpurun_out/launches_bench.log 2>&1
STEPS=2 $NCU --set full --import-source on -k regex:lio_update_kernel -s 1 -c 1 -f -o gpurun_out/prof_lio_update python tools/profile_driver.py > gpurun_out/prof_a.log 2>&1
STEPS=2 $NCU --set full --import-source on -k regex:vio_update_kernel -s 1 -c 1 -f -o gpurun_out/prof_vio_update python tools/profile_driver.py > gpurun_out/prof_b.log 2>&1
ESIKF_LOOP=0 STEPS=2 $NCU --set full --import-source on -k regex:lio_residual_kernel -s 3 -c 1 -f -o gpurun_out/prof_lio_residual python tools/profile_driver.py > gpurun_out/prof_c.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches.csv
