"""diagnostic: pose MPJPE delta of the chained bf16 pipeline vs the fp32 oracle lifting of the same refined features,
with the TF32 copy of the refined features (default) and with conv_frame_feat.0 reading the bf16 copy"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from egorear_b200 import synth, calib
from egorear_b200.pipeline import HotPathPipeline
from oracle import model_ref
dev = torch.device("cuda", 0)
for tf32_final in (True, False):
    pipe = HotPathPipeline(4, "ego4view_syn", "bf16", dev, tf32_final=tf32_final)
    sdp = {k: v.cpu() for k, v in pipe.pose3d.state_dict().items()}
    worst, mean = 0.0, []
    for seed in range(6):
        feat, bfb = synth.synth_features(4, 4, seed=seed)
        out = pipe(feat.to(dev), bfb.to(dev))
        with torch.no_grad():
            preds = model_ref.pose3d_forward(sdp, feat, out["list_ff"][-1].cpu(), calib.load_calibration(None), "ego4view_syn")
        for l in range(len(preds)):
            d = (out["list_pose3d"][l].cpu() - preds[l]).norm(dim=-1).mean(dim=-1)
            worst = max(worst, float(d.max())); mean.append(float(d.mean()))
    print("tf32_final=%s: MPJPE delta worst %.3e cm, mean %.3e cm (budget 1e-2 cm)" % (tf32_final, worst, sum(mean) / len(mean)))
