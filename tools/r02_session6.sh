#!/bin/bash
set -u
TAG=${1:-r02h}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_${TAG}.log 2>&1
echo "pytest rc=$?"; tail -16 gpurun_out/pytest_${TAG}.log
timeout 600 python tools/hbm_probe.py > gpurun_out/hbm_${TAG}.json 2> gpurun_out/hbm_${TAG}.err
echo "hbm rc=$?"; python - <<P
import json
try:
    d=json.load(open('gpurun_out/hbm_${TAG}.json'))
    for k,v in d.get('kernels',d).items():
        if isinstance(v,dict) and 'gbs' in v: print("%-36s %8.1f GB/s  %.3f"%(k,v['gbs'],v['frac']))
except Exception as e: print('ERR',e)
P
for w in n2 f2 h2o h2o-spinorb h2o-tz-spinorb h2o-tz; do
  timeout 900 python bench.py --workload $w --steps 3 --warmup 1 > gpurun_out/sample_${w}_${TAG}.json 2> gpurun_out/sample_${w}_${TAG}.err
  echo "sample $w rc=$?"
  python - <<P
import json
try:
    d=json.loads(open('gpurun_out/sample_${w}_${TAG}.json').read().strip().splitlines()[-1])
    print('  ', d['config']['workload'], 'value', d['value'], 'stages', d['stages_s'], 'device_ms', d['device_ms'], 'published', d['reference_published_s'], 'parity', d['parity'])
except Exception as e: print('ERR',e)
P
done
for w in n2 f2 h2o; do
  timeout 900 python bench.py --workload $w --impl reference --steps 3 --warmup 1 > gpurun_out/sampleref_${w}_${TAG}.json 2> gpurun_out/sampleref_${w}_${TAG}.err
  echo "sample ref $w rc=$?"; python - <<P
import json
try:
    d=json.loads(open('gpurun_out/sampleref_${w}_${TAG}.json').read().strip().splitlines()[-1])
    print('  ', d['value'], d['stages_s'], d['cpu_baseline']['cores'], d.get('ccsd_iterations'))
except Exception as e: print('ERR',e)
P
done
du -sh gpurun_out
