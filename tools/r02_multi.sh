#!/bin/bash
# Multi-GPU validation on one box with NG GPUs: sharded-update parity against the un-sharded result (new build, and the
# round-1 build for the root cause of its N = 4 / 8 divergence), then the bench at N = 2 .. NG.
NG=${NG:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 2 $NG; do
  PARITY_CFG=cfg2 PARITY_MODES=p2p,nccl PARITY_STEPS=30 timeout 600 $TR --nproc-per-node $n --master-port 2960$n tools/multi_gpu_parity.py > gpurun_out/multi_parity_n$n.log 2>&1
  echo "parity n=$n rc=$?" | tee -a gpurun_out/multi_parity_n$n.log
  grep -E "PARITY|MISMATCH|N=1 reference" gpurun_out/multi_parity_n$n.log | head -20
done
if [ -d ab/old_tree ]; then
  (cd ab/old_tree && PARITY_CFG=cfg2 PARITY_MODES=p2p,nccl PARITY_STEPS=10 timeout 600 $TR --nproc-per-node $NG --master-port 29650 tools/multi_gpu_parity.py > ../../gpurun_out/multi_parity_round1_build_n$NG.log 2>&1; echo "round-1 build parity n=$NG rc=$?" >> ../../gpurun_out/multi_parity_round1_build_n$NG.log)
  grep -E "PARITY|MISMATCH|N=1 reference|rc=" gpurun_out/multi_parity_round1_build_n$NG.log | head -30
fi
for n in 2 $NG; do
  timeout 600 $TR --nproc-per-node $n --master-port 2970$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/multi_bench_cfg2_n$n.json 2> gpurun_out/multi_bench_cfg2_n$n.err
  echo "bench cfg2 n=$n rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/multi_bench_cfg2_n$n.json').read().strip().splitlines()[-1])
    print('n=$n value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'median',round(d['ms_per_step_median'],4),'max',round(d['ms_per_step_max'],4),'e2e',round(d['e2e']['value']),'parity ok',d['parity_vs_oracle']['ok'],d['parity_vs_oracle']['matched_points'],d['parity_vs_oracle'].get('vio_iters_per_level'))
except Exception as e: print('parse',e)
PY
done
timeout 600 $TR --nproc-per-node $NG --master-port 29710 bench.py --gpus $NG --steps 20 --warmup 3 --config cfg4 > gpurun_out/multi_bench_cfg4_n$NG.json 2> gpurun_out/multi_bench_cfg4_n$NG.err
echo "bench cfg4 n=$NG rc=$?"; head -c 400 gpurun_out/multi_bench_cfg4_n$NG.json; echo
timeout 600 python bench.py --config cfg4 --no-shim > gpurun_out/multi_bench_cfg4_n1.json 2> gpurun_out/multi_bench_cfg4_n1.err
echo "bench cfg4 n=1 rc=$?"; head -c 300 gpurun_out/multi_bench_cfg4_n1.json; echo
