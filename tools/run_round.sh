#!/bin/bash
# usage: tools/run_round.sh [tag]   — the round's single-GPU evidence on one fresh B200 box:
#   pytest -m gpu, smoke(), every bench workload (one JSON line each), the reference arm.
# Outputs under gpurun_out/<tag>_*; copy the JSON lines you want judged into profiles/.
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x -s > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/${TAG}_tests.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/${TAG}_smoke.log)"
run() {  # name, args...
  name=$1; shift
  timeout 600 python bench.py "$@" > gpurun_out/${TAG}_bench_${name}.json 2> gpurun_out/${TAG}_bench_${name}.err
  echo "$name rc=$? $(head -c 200 gpurun_out/${TAG}_bench_${name}.json)"
}
run n1
run n1_bf16 --precision bf16
run reference_arm --impl reference
run mvfex_b64 --workload mvfex
run pose3d_b1024 --workload pose3d --batch 1024
run rw_e2e_b512 --workload rw_e2e --steps 4 --warmup 3
run generate_target --workload generate_target
run decode --workload decode
run eval_heatmap --workload eval_heatmap
run eval_pose --workload eval_pose
run preprocess --workload preprocess
