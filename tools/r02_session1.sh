#!/bin/bash
# Round-2 GPU session 1 (one B200): soak of the TMA kernel, bench at TMA scope 1 and 2, ncu launch list, ncu --set full of
# the shipped (T) GEMM + fused epilogue and of the CCSD GEMM shapes through the TMA kernel.
set -u
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,serial,clocks.max.sm --format=csv,noheader > gpurun_out/gpu_${TAG}.txt
timeout 900 python tools/gemm_soak.py --reps 1000 --out gpurun_out/soak_${TAG}.json > /dev/null 2> gpurun_out/soak_${TAG}.err
echo "soak rc=$?"; tail -3 gpurun_out/soak_${TAG}.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_${TAG}_scope1.json 2> gpurun_out/bench_${TAG}_scope1.err
echo "bench scope1 rc=$?"
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --tma 2 > gpurun_out/bench_${TAG}_scope2.json 2> gpurun_out/bench_${TAG}_scope2.err
echo "bench scope2 rc=$?"
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_*scope*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ccsd_s_per_iter'], d['t_wall_s'], d['roofline']['frac'], d['gemm_tflops_executed'], d['energies'])
    except Exception as e: print(f, 'ERR', e)
P
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_${TAG}.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "launch list rc=$?"
python tools/summarize_launches.py gpurun_out/launches_${TAG}.csv gpurun_out/launches_${TAG}.txt > /dev/null 2>&1
head -30 gpurun_out/launches_${TAG}.txt
PROFILE=T timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'gemm_f64_tma|k_triples_fused' -c 4 -f \
  -o gpurun_out/prof_T_${TAG} python tools/ncu_target.py > gpurun_out/ncu_T_${TAG}.log 2>&1
echo "ncu T rc=$?"
python tools/ncu_summary.py gpurun_out/prof_T_${TAG}.ncu-rep gpurun_out/ncu_T_${TAG}_summary
PROFILE=CCSD TMA_SCOPE=2 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:gemm_f64 -c 4 -f -o gpurun_out/prof_CCSD_${TAG} python tools/ncu_target.py > gpurun_out/ncu_CCSD_${TAG}.log 2>&1
echo "ncu CCSD rc=$?"
python tools/ncu_summary.py gpurun_out/prof_CCSD_${TAG}.ncu-rep gpurun_out/ncu_CCSD_${TAG}_summary
du -sh gpurun_out
