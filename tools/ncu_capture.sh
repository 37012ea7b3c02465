#!/bin/bash
# usage: tools/ncu_capture.sh   (on a GPU box, after tools/profile_step.py has exited 0 without ncu)
#   launch list of one fp16 step + ncu --set full captures of the kernels profiles/r02_ncu_summary.md tabulates;
#   outputs under gpurun_out/r2B_*  -> tools/ncu_summary.py
set -x
mkdir -p gpurun_out
timeout 200 python tools/profile_step.py --precision fp16 > gpurun_out/r2B_plain.log 2>&1; tail -2 gpurun_out/r2B_plain.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2B_launches_step.csv python tools/profile_step.py --precision fp16 > gpurun_out/r2B_ncu1.log 2>&1
for spec in "H1tail:head_tail_mma_kernel:1" "F1b:gemm_tc_kernel:16" "fold:gemm_tc_kernel:10" "tok_sample:tok_sample_kernel:1"; do
  name=${spec%%:*}; rest=${spec#*:}
  if [ "${rest%%:*}" = "regex" ]; then rest=${rest#regex:}; kn="regex:${rest%:*}"; else kn="${rest%:*}"; fi
  skip=${spec##*:}
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k "$kn" -s $((skip-1)) -c 1 -o gpurun_out/r2B_$name -f python tools/profile_step.py --precision fp16 > gpurun_out/r2B_ncu_$name.log 2>&1
  ls -la gpurun_out/r2B_$name.ncu-rep
done
