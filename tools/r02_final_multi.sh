#!/bin/bash
# Multi-GPU lines of the round on a box with NG GPUs: bench.py under torchrun exactly as the driver launches it.
NG=${NG:-2}
CFGS=${CFGS:-"cfg2 cfg5"}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
if [ "${TESTS:-1}" = "1" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_vio.py tests/test_gpu_z_new_entry_points.py -q -m gpu > gpurun_out/final_multi_tests_n$NG.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/final_multi_tests_n$NG.log; tail -n 5 gpurun_out/final_multi_tests_n$NG.log
fi
port=29730
for cfg in $CFGS; do
  port=$((port + 1))
  timeout 600 $TR --nproc-per-node $NG --master-port $port bench.py --gpus $NG --steps 20 --warmup 3 --config $cfg > gpurun_out/final_bench_${cfg}_n$NG.json 2> gpurun_out/final_bench_${cfg}_n$NG.err
  echo "bench $cfg n=$NG rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/final_bench_${cfg}_n$NG.json').read().strip().splitlines()[-1])
    print('$cfg n=$NG value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'median',round(d['ms_per_step_median'],4),'max',round(d['ms_per_step_max'],4),'e2e',round(d['e2e']['value']),'parity ok',d['parity_vs_oracle']['ok'],d['parity_vs_oracle'].get('matched_points'))
except Exception as e:
    print('parse',e)
PY
  tail -n 3 gpurun_out/final_bench_${cfg}_n$NG.err
done
