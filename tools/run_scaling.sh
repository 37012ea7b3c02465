#!/bin/bash
# usage: tools/run_scaling.sh N [tag]   — the multi-GPU lines of BASELINE.json's configs on N GPUs of one box (torchrun, NCCL):
#   headline (config 2 + pose3d lifting, 64 frames / GPU / step), config 4 (generate_target sweep) and config 5 (rw end to end
#   from images, 512 frames / GPU / step).  Outputs: gpurun_out/<tag>_n<N>_{headline,gt,rw_e2e}.json
N=${1:-2}; TAG=${2:-r02}
mkdir -p gpurun_out
run() {  # name, args...
  name=$1; shift
  if [ "$N" = "1" ]; then
    timeout 400 python bench.py --gpus 1 "$@" > gpurun_out/${TAG}_n${N}_${name}.json 2> gpurun_out/${TAG}_n${N}_${name}.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N "$@" > gpurun_out/${TAG}_n${N}_${name}.json 2> gpurun_out/${TAG}_n${N}_${name}.err
  fi
  echo "$name rc=$? $(head -c 300 gpurun_out/${TAG}_n${N}_${name}.json)"
}
run headline --steps 30 --warmup 5
run gt --workload generate_target --steps 20 --warmup 3
run rw_e2e --workload rw_e2e --steps 4 --warmup 3 --sustain-s 0
