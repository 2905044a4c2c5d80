#!/bin/bash
# A/B timing of library variants built from the same sources with different -D switches (ab/*.so), one B200.
mkdir -p gpurun_out
STAMPS=1 MODES=2 STEPS=30 timeout 300 python tools/loop_mode_check.py > gpurun_out/ab_default.log 2>&1
echo "== default build:"; grep -E "it/s|^LIO|per-CTA|LOOP" gpurun_out/ab_default.log
for v in ab/libesikf_*.so; do
  n=$(basename $v .so)
  ESIKF_LIB=$PWD/$v STAMPS=1 MODES=2 STEPS=30 timeout 300 python tools/loop_mode_check.py > gpurun_out/ab_$n.log 2>&1
  echo "== $n:"; grep -E "it/s|^LIO|per-CTA|LOOP" gpurun_out/ab_$n.log
done
