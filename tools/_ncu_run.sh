set -x
mkdir -p gpurun_out
timeout 200 python tools/profile_step.py --precision fp16 > gpurun_out/r2v_plain.log 2>&1; tail -2 gpurun_out/r2v_plain.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2v_launches_step.csv python tools/profile_step.py --precision fp16 > gpurun_out/r2v_ncu1.log 2>&1
for spec in "H1tail:head_tail_mma_kernel:1" "F1b:regex:gemm_tc_kernel<256, __half, __half>:3" "fold:regex:gemm_tc_kernel<64, __half, float>:2" "tok_sample:tok_sample_kernel:1"; do
  name=${spec%%:*}; rest=${spec#*:}
  if [ "${rest%%:*}" = "regex" ]; then rest=${rest#regex:}; kn="regex:${rest%:*}"; else kn="${rest%:*}"; fi
  skip=${spec##*:}
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k "$kn" -s $((skip-1)) -c 1 -o gpurun_out/r2v_$name -f python tools/profile_step.py --precision fp16 > gpurun_out/r2v_ncu_$name.log 2>&1
  ls -la gpurun_out/r2v_$name.ncu-rep
done
