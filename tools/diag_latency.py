"""diagnostic: single-stream latency of the chained forward at small batch sizes (CUDA events, median of 50)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from egorear_b200 import synth
from egorear_b200.pipeline import HotPathPipeline
dev = torch.device("cuda", 0)
pipe = HotPathPipeline(4, "ego4view_syn", "bf16", dev, materialize_features=False)
pipe.freeze()
for B in (1, 2, 4, 8, 16, 32, 64):
    feat, bfb = [t.to(dev) for t in synth.synth_features(B, 4, seed=1)]
    for _ in range(10):
        pipe(feat, bfb)
    torch.cuda.synchronize()
    ts = []
    for _ in range(50):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pipe(feat, bfb); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print("B=%2d  median %.3f ms  min %.3f ms  -> %.0f frames/s" % (B, ts[25], ts[0], B / ts[25] * 1e3))
print("--- CUDA graph replay ---")
for B in (1, 4, 16, 64):
    feat, bfb = [t.to(dev) for t in synth.synth_features(B, 4, seed=1)]
    want = pipe(feat, bfb)["packed"].clone()
    pipe.capture(B)
    got = pipe.replay(feat, bfb)["packed"]
    torch.cuda.synchronize()
    assert torch.equal(got, want), "graph replay differs from the eager forward"
    ts = []
    for _ in range(50):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pipe.replay(feat, bfb); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print("B=%2d  graph median %.3f ms  min %.3f ms  -> %.0f frames/s" % (B, ts[25], ts[0], B / ts[25] * 1e3))
