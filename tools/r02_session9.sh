#!/bin/bash
set -u
TAG=${1:-r02l}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_${TAG}.err
python - <<P
import json
d=json.loads(open("gpurun_out/bench_${TAG}.json").read().strip().splitlines()[-1])
print("value",d["value"],"ccsd",d["ccsd_s_per_iter"],"T",d["t_wall_s"],"frac",d["roofline"]["frac"],d["gemm_tflops_executed"],"e2e",d["e2e"]["value"],d["e2e"]["breakdown_s"],"parity",d["parity"])
t=d["target_config"]; print("target",t.get("value"),t.get("ccsd_s_per_iter"),t.get("t_wall_s"),t.get("roofline",{}).get("frac"),t.get("e2e",{}).get("value"),t.get("parity"),t.get("error"))
P
