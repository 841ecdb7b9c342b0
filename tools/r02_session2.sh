#!/bin/bash
# Round-2 GPU session: parity tests, bench at nbf=200, ncu of the (T) kernels.
set -u
TAG=${1:-r02b}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_${TAG}.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_${TAG}.err
python - <<P
import json
try:
    d=json.loads(open('gpurun_out/bench_${TAG}.json').read().strip().splitlines()[-1])
    print('value',d['value'],'ccsd',d['ccsd_s_per_iter'],'T',d['t_wall_s'],'frac',d['roofline']['frac'],d['roofline'].get('ms_per_launch'),d['gemm_tflops_executed'],d['energies'], d['e2e']['value'])
except Exception as e: print('ERR',e)
P
PROFILE=T timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'gemm_f64_tma|k_triples_fused' -c 4 -f \
  -o gpurun_out/prof_T_${TAG} python tools/ncu_target.py > gpurun_out/ncu_T_${TAG}.log 2>&1
echo "ncu T rc=$?"; tail -3 gpurun_out/ncu_T_${TAG}.log
python tools/ncu_summary.py gpurun_out/prof_T_${TAG}.ncu-rep gpurun_out/ncu_T_${TAG}_summary
du -sh gpurun_out
