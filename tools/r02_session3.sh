#!/bin/bash
set -u
TAG=${1:-r02d}
mkdir -p gpurun_out
timeout 900 python tools/gemm_soak.py --reps 300 --out gpurun_out/soak_${TAG}.json > /dev/null 2> gpurun_out/soak_${TAG}.err
echo "soak rc=$?"; tail -2 gpurun_out/soak_${TAG}.err | cut -c1-400
python - <<P
import json
d=json.load(open('gpurun_out/soak_${TAG}.json'))
print('mismatches',d['total_mismatches'],'elements',d['total_elements'])
for r in d['shapes']: print("%-45s bad %d tma %.2f cpasync %.2f"%(r['shape'],r['mismatches'],r['tma_tflops'],r['cpasync_tflops']))
P
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_${TAG}.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_${TAG}.err
python - <<P
import json
try:
    d=json.loads(open('gpurun_out/bench_${TAG}.json').read().strip().splitlines()[-1])
    print('value',d['value'],'ccsd',d['ccsd_s_per_iter'],'T',d['t_wall_s'],'frac',d['roofline']['frac'],d['roofline'].get('ms_per_launch'),d['gemm_tflops_executed'],d['energies'], d['e2e']['value'])
    t=d['target_config']; print('target',t.get('value'),t.get('ccsd_s_per_iter'),t.get('t_wall_s'),t.get('roofline',{}).get('frac'),t.get('gemm_tflops_executed'),t.get('energies'),t.get('error'))
except Exception as e: print('ERR',e)
P
du -sh gpurun_out
