#!/bin/bash
# 8-GPU session: bench at N=8 (full, with target_config), exchange A/B at N=8, N=4 (headline shape only)
set -u
TAG=${1:-r02k}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29801 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench8_${TAG}.json 2> gpurun_out/bench8_${TAG}.err
echo "bench N=8 rc=$?"; tail -4 gpurun_out/bench8_${TAG}.err | cut -c1-300
AFESP_DIST_ALLGATHER=1 timeout 600 $TR --nproc-per-node 8 --master-port 29803 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu --target 0 > gpurun_out/bench8ag_${TAG}.json 2> gpurun_out/bench8ag_${TAG}.err
echo "bench N=8 allgather rc=$?"
timeout 600 $TR --nproc-per-node 4 --master-port 29802 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu --target 0 > gpurun_out/bench4_${TAG}.json 2> gpurun_out/bench4_${TAG}.err
echo "bench N=4 rc=$?"
python - <<P
import json,glob
for f in sorted(glob.glob('gpurun_out/bench[48]*_${TAG}.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value',round(d['value'],4),'ccsd',round(d['ccsd_s_per_iter'],4),'T',round(d['t_wall_s'],4),'e2e',round(d['e2e']['value'],4),d['e2e']['breakdown_s'],'parity',d['parity'].get('ok'),d['parity'].get('abs_diff'))
        t=d.get('target_config')
        if t: print('   target',t.get('value'),t.get('ccsd_s_per_iter'),t.get('t_wall_s'),t.get('ao2mo_s'),t.get('e2e',{}).get('value'),t.get('e2e',{}).get('breakdown_s'),t.get('parity'),t.get('error'))
    except Exception as e: print(f,'ERR',e)
P
