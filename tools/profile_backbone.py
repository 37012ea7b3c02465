#!/usr/bin/env python
"""One forward of the backbone engine at B frames for an ncu launch list:

    ncu --metrics gpu__time_duration.sum --clock-control none -c 140 --csv --log-file X.csv python tools/profile_backbone.py [B] [precision]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from egorear_b200.pipeline import HotPathPipeline  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16"
dev = torch.device("cuda", 0)
pipe = HotPathPipeline(4, "ego4view_rw", prec, dev, with_backbone=True, materialize_features=False)
img = torch.randn((B, 4, 3, 256, 256), generator=torch.Generator().manual_seed(0)).to(dev)
pipe.heatmap.backbone_engine()._sync_params()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
xh, bfb = pipe.backbone_staged(img)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok", tuple(xh.shape), float(xh.float().abs().mean()))
