#!/bin/bash
# 8-GPU validation: sharded-update parity (cfg2, cfg5), the bench at N = 8 (cfg2, cfg5) and N = 4 (cfg4), and the round-1
# build's own bench at N = 4 / 8 (does its matched_points / iteration-count divergence reproduce on cached frames?).
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
PARITY_CFG=cfg2 PARITY_MODES=p2p,nccl PARITY_STEPS=30 timeout 600 $TR --nproc-per-node 8 --master-port 29608 tools/multi_gpu_parity.py > gpurun_out/multi_parity_n8.log 2>&1
echo "parity cfg2 n=8 rc=$?" | tee -a gpurun_out/multi_parity_n8.log
grep -E "PARITY|MISMATCH|N=1 reference" gpurun_out/multi_parity_n8.log | head -20
PARITY_CFG=cfg5 PARITY_MODES=p2p PARITY_STEPS=10 timeout 600 $TR --nproc-per-node 8 --master-port 29609 tools/multi_gpu_parity.py > gpurun_out/multi_parity_cfg5_n8.log 2>&1
echo "parity cfg5 n=8 rc=$?" | tee -a gpurun_out/multi_parity_cfg5_n8.log
grep -E "PARITY|MISMATCH|N=1 reference" gpurun_out/multi_parity_cfg5_n8.log | head -20
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1],'value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'median',round(d['ms_per_step_median'],4),'max',round(d['ms_per_step_max'],4),'e2e',round(d['e2e']['value']),'parity ok',d['parity_vs_oracle']['ok'],d['parity_vs_oracle']['matched_points'],d['parity_vs_oracle'].get('vio_iters_per_level'),'frac',round(d['roofline']['frac'],3))
except Exception as e: print('parse',sys.argv[1],e)
PY
}
for c in cfg2 cfg5; do
  timeout 600 $TR --nproc-per-node 8 --master-port 29708 bench.py --gpus 8 --steps 20 --warmup 3 --config $c > gpurun_out/multi_bench_${c}_n8.json 2> gpurun_out/multi_bench_${c}_n8.err
  echo "bench $c n=8 rc=$?"; summ gpurun_out/multi_bench_${c}_n8.json
done
timeout 600 $TR --nproc-per-node 4 --master-port 29704 bench.py --gpus 4 --steps 20 --warmup 3 --config cfg4 > gpurun_out/multi_bench_cfg4_n4.json 2> gpurun_out/multi_bench_cfg4_n4.err
echo "bench cfg4 n=4 rc=$?"; summ gpurun_out/multi_bench_cfg4_n4.json
timeout 600 $TR --nproc-per-node 2 --master-port 29702 bench.py --gpus 2 --steps 20 --warmup 3 --config cfg5 > gpurun_out/multi_bench_cfg5_n2.json 2> gpurun_out/multi_bench_cfg5_n2.err
echo "bench cfg5 n=2 rc=$?"; summ gpurun_out/multi_bench_cfg5_n2.json
timeout 600 python bench.py --config cfg5 --no-shim > gpurun_out/multi_bench_cfg5_n1.json 2> gpurun_out/multi_bench_cfg5_n1.err
echo "bench cfg5 n=1 rc=$?"; summ gpurun_out/multi_bench_cfg5_n1.json
if [ -d ab/old_tree ]; then
  for n in 4 8; do
    (cd ab/old_tree && timeout 300 $TR --nproc-per-node $n --master-port 2980$n bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline > ../../gpurun_out/round1_build_bench_n$n.json 2> ../../gpurun_out/round1_build_bench_n$n.err; echo "round-1 build bench n=$n rc=$?")
    python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/round1_build_bench_n$n.json').read().strip().splitlines()[-1])
    print('round-1 build n=$n value',round(d['value']),'matched',d['config']['matched_points'],'lio',d['config']['lio_iters'],'vio',d['config']['vio_iters'])
except Exception as e: print('parse',e)
PY
  done
fi
