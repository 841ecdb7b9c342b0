#!/bin/bash
# 2-GPU session: multi-GPU parity tests, CCSD exchange overlap A/B, full bench at N=2
set -u
TAG=${1:-r02g}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi_${TAG}.log 2>&1
echo "pytest multi rc=$?"; tail -5 gpurun_out/pytest_multi_${TAG}.log
for ch in 1 2 4; do
  timeout 600 $TR --master-port $((29700+ch)) bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --target 0 --dist-chunks $ch > gpurun_out/bench2_${TAG}_ch${ch}.json 2> gpurun_out/bench2_${TAG}_ch${ch}.err
  echo "bench N=2 chunks=$ch rc=$?"
done
timeout 1200 $TR --master-port 29711 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench2_${TAG}_full.json 2> gpurun_out/bench2_${TAG}_full.err
echo "bench N=2 full rc=$?"; tail -3 gpurun_out/bench2_${TAG}_full.err
python - <<P
import json,glob
for f in sorted(glob.glob('gpurun_out/bench2_${TAG}_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value',round(d['value'],4),'ccsd',round(d['ccsd_s_per_iter'],4),'T',round(d['t_wall_s'],4),'e2e',round(d['e2e']['value'],4),d['e2e']['breakdown_s'],'parity',d['parity'].get('ok'),d['parity'].get('abs_diff'))
        t=d.get('target_config')
        if t: print('   target',t.get('value'),t.get('ccsd_s_per_iter'),t.get('t_wall_s'),t.get('e2e',{}).get('value'),t.get('parity'),t.get('error'))
    except Exception as e: print(f,'ERR',e)
P
du -sh gpurun_out
