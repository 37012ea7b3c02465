"""diagnostic: H2D bandwidth from pinned memory on the default vs a side stream, alone and under a concurrent forward"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from egorear_b200 import synth
from egorear_b200.pipeline import HotPathPipeline
dev = torch.device("cuda", 0)
feat_h, bfb_h = synth.synth_features(64, 4, seed=1)
feat_h, bfb_h = feat_h.pin_memory(), bfb_h.pin_memory()
d = torch.empty_like(feat_h, device=dev)
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
print("default stream copy_  %.2f ms" % t(lambda: d.copy_(feat_h, non_blocking=True)))
side = torch.cuda.Stream(dev)
def side_copy():
    with torch.cuda.stream(side):
        d.copy_(feat_h, non_blocking=True)
print("side stream copy_     %.2f ms" % t(side_copy))
pipe = HotPathPipeline(4, "ego4view_syn", "bf16", dev)
feat, bfb = feat_h.to(dev), bfb_h.to(dev)
pipe.freeze()
print("forward alone         %.2f ms" % t(lambda: pipe(feat, bfb)))
def both():
    with torch.cuda.stream(side):
        d.copy_(feat_h, non_blocking=True)
    pipe(feat, bfb)
print("copy(side) || forward %.2f ms" % t(both))
def gen(n):
    return pipe.infer_host_batches(((feat_h, bfb_h) for _ in range(n)))
for _ in gen(2): pass
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in gen(8): pass
torch.cuda.synchronize(); print("infer_host_batches    %.2f ms/batch" % ((time.perf_counter() - t0) / 8 * 1e3))
def seq():
    f = feat_h.to(dev, non_blocking=True); b = bfb_h.to(dev, non_blocking=True)
    return pipe(f, b)["packed"].cpu()
print("sequential e2e        %.2f ms" % t(seq))
