#!/usr/bin/env python
"""One steady-state step of the bench workload inside a cudaProfilerStart/Stop window (for ncu --profile-from-start off).

    python tools/profile_step.py [--batch 64] [--precision bf16] [--steps 1]

Prints the step time measured with CUDA events (never quote a number taken under ncu).
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--all-outputs", action="store_true", help="materialise the NCHW fp32 refined features (module outputs)")
    args = ap.parse_args()
    import torch
    from egorear_b200 import synth
    from egorear_b200.pipeline import HotPathPipeline
    dev = torch.device("cuda", 0)
    # default: the bench's headline configuration (chained model, refined features stay channels-last)
    pipe = HotPathPipeline(4, "ego4view_syn", args.precision, dev, materialize_features=args.all_outputs)
    feat, bfb = synth.synth_features(min(args.batch, 64), 4, seed=100)
    if args.batch > 64:
        feat, bfb = feat.repeat(args.batch // 64, 1, 1, 1, 1), bfb.repeat(args.batch // 64, 1, 1, 1, 1)
    feat, bfb = feat.to(dev), bfb.to(dev)
    pipe.freeze()
    for _ in range(3):
        pipe(feat, bfb)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.cudart().cudaProfilerStart()
    e0.record()
    for _ in range(args.steps):
        pipe(feat, bfb)
    e1.record()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("step %.3f ms (batch %d, %s)" % (e0.elapsed_time(e1) / args.steps, args.batch, args.precision))


if __name__ == "__main__":
    main()
