#!/bin/bash
# One GPU session: plain bench (exit 0 first), then the ncu launch list of the same command, then ncu --set full of the
# (T) kernels and of the CCSD-iteration GEMMs.  Outputs under gpurun_out/ (kept under the 64 MiB return limit: few
# kernels per capture; text/json summaries are made on the box as well).
set -u
TAG=${1:-v2}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench rc=$?"; cat gpurun_out/bench_${TAG}.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_${TAG}.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "launch list rc=$?"
python tools/summarize_launches.py gpurun_out/launches_${TAG}.csv gpurun_out/launches_${TAG}.txt > /dev/null 2>&1
head -30 gpurun_out/launches_${TAG}.txt
PROFILE=T timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'gemm_f64_tma|k_triples_fused' -c 4 -f \
  -o gpurun_out/prof_T_${TAG} python tools/ncu_target.py > gpurun_out/ncu_T_${TAG}.log 2>&1
echo "ncu T rc=$?"
python tools/ncu_summary.py gpurun_out/prof_T_${TAG}.ncu-rep gpurun_out/ncu_T_${TAG}_summary
PROFILE=CCSD timeout 900 ncu --set full --clock-control none --profile-from-start off \
  -k regex:gemm_f64 -c 4 -f -o gpurun_out/prof_CCSD_${TAG} python tools/ncu_target.py > gpurun_out/ncu_CCSD_${TAG}.log 2>&1
echo "ncu CCSD rc=$?"
python tools/ncu_summary.py gpurun_out/prof_CCSD_${TAG}.ncu-rep gpurun_out/ncu_CCSD_${TAG}_summary
du -sh gpurun_out
