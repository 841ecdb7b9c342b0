#!/bin/bash
# 2-GPU: multi-GPU parity tests + all-gather vs grouped-broadcast exchange A/B
set -u
TAG=${1:-r02m}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi_${TAG}.log 2>&1
echo "pytest multi rc=$?"; tail -3 gpurun_out/pytest_multi_${TAG}.log
for ag in 1 0; do
  AFESP_DIST_ALLGATHER=$ag timeout 600 $TR --master-port $((29720+ag)) bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --target 0 > gpurun_out/bench2_${TAG}_ag${ag}.json 2> gpurun_out/bench2_${TAG}_ag${ag}.err
  echo "bench N=2 allgather=$ag rc=$?"
done
python - <<P
import json,glob
for f in sorted(glob.glob('gpurun_out/bench2_${TAG}_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value',round(d['value'],4),'ccsd',round(d['ccsd_s_per_iter'],4),'T',round(d['t_wall_s'],4),'e2e',round(d['e2e']['value'],4),'parity',d['parity'].get('ok'),d['parity'].get('abs_diff'))
    except Exception as e: print(f,'ERR',e)
P
