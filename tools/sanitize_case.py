"""smallest case that touches every hot-path kernel once (for `compute-sanitizer --tool memcheck python tools/sanitize_case.py`)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from egorear_b200 import ops, synth
from egorear_b200.pipeline import HotPathPipeline
dev = torch.device("cuda", 0)
for mat in (False, True):
    pipe = HotPathPipeline(4, "ego4view_syn", "bf16", dev, materialize_features=mat)
    feat, bfb = [t.to(dev) for t in synth.synth_features(1, 4, seed=3)]
    out = pipe(feat, bfb)
    torch.cuda.synchronize()
    print("chained forward ok (materialize_features=%s):" % mat, tuple(out["packed"].shape), float(out["packed"].abs().sum()))
kp = torch.from_numpy(synth.synth_keypoints(2, 4, 16, seed=0)).to(dev)
hm = ops.generate_target_batch(kp)
p, m, v = ops.get_max_preds(hm.view(8, 16, 64, 64), 0.5, True)
ps, ms = ops.get_max_preds_soft_pytorch(hm.view(8, 16, 64, 64))
torch.cuda.synchronize()
print("G1 / D1 / D1s ok", float(p.sum()), float(ps.sum()))
