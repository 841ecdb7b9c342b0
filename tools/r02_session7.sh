#!/bin/bash
# Final-kernel evidence: plain bench -> ncu launch list of the same command -> ncu --set full of the (T) kernels at
# nbf=200 and nbf=400 and of the CCSD GEMM shapes.
set -u
TAG=${1:-r02j}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu --target 0 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_${TAG}.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --target 0 > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "launch list rc=$?"
python tools/summarize_launches.py gpurun_out/launches_${TAG}.csv gpurun_out/launches_${TAG}.txt > /dev/null 2>&1
head -24 gpurun_out/launches_${TAG}.txt
PROFILE=T timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'gemm_f64_tma|k_triples_fused' -c 6 -f -o gpurun_out/prof_T200_${TAG} python tools/ncu_target.py > gpurun_out/ncu_T200_${TAG}.log 2>&1
echo "ncu T200 rc=$?"
python tools/ncu_summary.py gpurun_out/prof_T200_${TAG}.ncu-rep gpurun_out/ncu_T200_${TAG}_summary
NBF=400 NOCC=40 PROFILE=T timeout 1500 ncu --set full --clock-control none --profile-from-start off \
  -k regex:'gemm_f64_tma|k_triples_fused' -c 3 -f -o gpurun_out/prof_T400_${TAG} python tools/ncu_target.py > gpurun_out/ncu_T400_${TAG}.log 2>&1
echo "ncu T400 rc=$?"; tail -2 gpurun_out/ncu_T400_${TAG}.log
python tools/ncu_summary.py gpurun_out/prof_T400_${TAG}.ncu-rep gpurun_out/ncu_T400_${TAG}_summary
PROFILE=CCSD timeout 900 ncu --set full --clock-control none --profile-from-start off \
  -k regex:gemm_f64 -c 6 -f -o gpurun_out/prof_CCSD_${TAG} python tools/ncu_target.py > gpurun_out/ncu_CCSD_${TAG}.log 2>&1
echo "ncu CCSD rc=$?"
python tools/ncu_summary.py gpurun_out/prof_CCSD_${TAG}.ncu-rep gpurun_out/ncu_CCSD_${TAG}_summary
rm -f gpurun_out/prof_T400_${TAG}.ncu-rep gpurun_out/prof_CCSD_${TAG}.ncu-rep
du -sh gpurun_out
