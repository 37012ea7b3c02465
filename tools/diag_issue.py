"""diagnostic: host-side enqueue time per step vs device time per step"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from egorear_b200 import synth
from egorear_b200.pipeline import HotPathPipeline
dev = torch.device("cuda", 0)
pipe = HotPathPipeline(4, "ego4view_syn", "bf16", dev)
feat, bfb = [t.to(dev) for t in synth.synth_features(64, 4, seed=1)]
pipe.freeze()
for lanes in (1, 2):
    fn = (lambda: pipe(feat, bfb)) if lanes == 1 else (lambda: pipe.forward_async(feat, bfb, lanes=2))
    for _ in range(5): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(30): fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("lanes %d: host enqueue %.3f ms/step, total %.3f ms/step" % (lanes, (t1 - t0) / 30 * 1e3, (t2 - t0) / 30 * 1e3))
