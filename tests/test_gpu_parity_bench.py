"""GPU: parity AT THE BENCHMARKED batch sizes (BASELINE.json configs 2, 3 and 5), where the dense kernel picks other tile
widths / split-K factors and its persistent loop walks many tiles per CTA (the small-batch parity tests never get there).

The pipeline runs the full batch; the oracle (oracle/parity.py -> oracle/model_ref.py) runs a seeded sample of 16 of its
frames — frames are independent.  Next to the bounds on heatmaps / features / 3D joints each case prints the
fraction of decoded 2D joints whose argmax cell equals the reference's and the largest displacement of the rest.

Bounds: fp32 <= 1e-3 relative (measured 3e-6).  fp16 (the default tensor-core mode: fp16 operands, split 1x1 weights, 3x-TF32 /
fp16-pair token Linears, and four 32x32 activations of the refine path kept as fp16 pairs, option `asplit`) <= 1e-3:
measured 6.3e-4 / 9.1e-4 at B = 64, 5.7e-4 / 8.1e-4 at B = 512, 5.6e-4 / 6.9e-4 at B = 2.  The margin is real but not wide:
ten independent 10-bit roundings are left between the input features and the refined heatmap, and the max-norm of a 16-frame
sample scatters by +-15 % with the sample and with the summation order of any kernel on the path.  `asplit=2` (the 512-channel
F1b output as a pair too) measures 8.5e-4 at B = 64 for another 3 % of the step; without any pairs (`asplit=0`, 7 % faster:
fourteen roundings) the same case measures 9.9e-4 ... 1.1e-3, i.e. AT the line - that variant states 1.5e-3
(test_config2_b64_fp16_without_activation_pairs).  bf16 states 1e-2 (measured 5.7e-3 / 8.4e-3).  3D joints of the CHAINED
model (GPU features -> GPU lifting vs oracle features -> oracle lifting): <= 0.1 mm MPJPE delta in fp32 and fp16 (measured
0.04 mm); the bf16 mode does not meet that end to end (measured 0.18-0.21 mm: its feature error feeds the lifting) and states
0.3 mm - one more reason fp16 is the default precision.  Decoded argmax cells: random-init heatmaps are
almost flat (no trained peak), so two cells a few 1e-4 apart swap under any rounding; the fractions are reported and only
loosely bounded.
"""
import json

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N_SAMPLE = 16
HM_TOL = {"fp32": 1e-3, "fp16": 1e-3, "bf16": 1e-2}
# decoded argmax cells: random-init heatmaps have broad, flat maxima, so reduced-precision heatmaps may move a cell
ANCHOR_FRAC_MIN = {"fp32": 0.995, "fp16": 0.98, "bf16": 0.95}       # init-heatmap argmax cells (the cross-attention anchors)
JOINT_FRAC_MIN = {"fp32": 0.98, "fp16": 0.8, "bf16": 0.6}           # refined-heatmap argmax cells, oracle end to end
MPJPE_MM = {"fp32": 0.1, "fp16": 0.1, "bf16": 0.3}


def _device_features(B, seed, dev):
    g = torch.Generator(device=dev).manual_seed(seed)
    feat = torch.relu(torch.randn((B, 4, 128, 64, 64), generator=g, device=dev))
    bfb = torch.relu(torch.randn((B, 4, 512, 8, 8), generator=g, device=dev))
    return feat, bfb


def _check(res, precision, with_feat):
    print("parity %s: %s" % (precision, json.dumps(res)))
    tol = HM_TOL[precision]
    assert res["hm_init_rel"] < tol and res["hm_refined_rel"] < tol, res
    if with_feat:
        assert res["feat_refined_rel"] < tol, res
    assert res["mpjpe_delta_mm"] < MPJPE_MM[precision], res
    assert res["anchors_same_frac"] >= ANCHOR_FRAC_MIN[precision] and res["joints2d_same_frac"] >= JOINT_FRAC_MIN[precision], res


@pytest.mark.parametrize("precision", ["fp16", "bf16", "fp32"])
def test_config2_b64_chained(precision):
    """config 2 + lifting, batch 64, syn cameras: the headline workload of bench.py (chained, features internal)"""
    from egorear_b200 import calib, synth
    from egorear_b200.pipeline import HotPathPipeline
    from oracle import parity
    dev = torch.device("cuda", 0)
    B = 64
    feat, bfb = synth.synth_features(B, 4, seed=300)
    pipe = HotPathPipeline(4, "ego4view_syn", precision, dev, materialize_features=(precision == "fp32"))
    out = pipe(feat.to(dev), bfb.to(dev))
    idx = parity.sample_indices(B, N_SAMPLE, seed=1)
    sd_h = {k: v.cpu() for k, v in pipe.heatmap.state_dict().items()}
    sd_p = {k: v.cpu() for k, v in pipe.pose3d.state_dict().items()}
    res = parity.hot_path_parity(out, pipe.heatmap.last_anchors[0], idx, feat[idx], bfb[idx], sd_h, sd_p,
                                 calib.load_calibration(None), "ego4view_syn")
    _check(res, precision, precision == "fp32")


def test_config2_b64_fp16_without_activation_pairs():
    """the faster fp16 variant (option asplit=0: every activation a single fp16 tensor): same case, stated bound 1.5e-3"""
    from egorear_b200 import calib, engine, synth
    from egorear_b200.pipeline import HotPathPipeline
    from oracle import parity
    dev = torch.device("cuda", 0)
    B = 64
    feat, bfb = synth.synth_features(B, 4, seed=300)
    engine.set_option("asplit", 0)
    try:
        pipe = HotPathPipeline(4, "ego4view_syn", "fp16", dev, materialize_features=False)
        out = pipe(feat.to(dev), bfb.to(dev))
    finally:
        engine.set_option("asplit", 1)
    idx = parity.sample_indices(B, N_SAMPLE, seed=1)
    sd_h = {k: v.cpu() for k, v in pipe.heatmap.state_dict().items()}
    sd_p = {k: v.cpu() for k, v in pipe.pose3d.state_dict().items()}
    res = parity.hot_path_parity(out, pipe.heatmap.last_anchors[0], idx, feat[idx], bfb[idx], sd_h, sd_p,
                                 calib.load_calibration(None), "ego4view_syn")
    print("parity fp16 asplit=0: %s" % json.dumps(res))
    assert res["hm_init_rel"] < 1.5e-3 and res["hm_refined_rel"] < 1.5e-3, res
    assert res["mpjpe_delta_mm"] < MPJPE_MM["fp16"], res


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
def test_config5_b512_rw(precision):
    """config 5's hot path: rw cameras (per-frame device->camera transforms), 512 frames per GPU"""
    from egorear_b200 import calib, synth
    from egorear_b200.pipeline import HotPathPipeline
    from oracle import parity
    dev = torch.device("cuda", 0)
    B = 512
    feat, bfb = _device_features(B, 17, dev)
    ctm = synth.synth_coord_trans_mat(B, seed=3)
    pipe = HotPathPipeline(4, "ego4view_rw", precision, dev, materialize_features=False)
    out = pipe(feat, bfb, ctm.to(dev))
    idx = parity.sample_indices(B, N_SAMPLE, seed=2)
    ii = torch.as_tensor(idx, device=dev)
    sd_h = {k: v.cpu() for k, v in pipe.heatmap.state_dict().items()}
    sd_p = {k: v.cpu() for k, v in pipe.pose3d.state_dict().items()}
    res = parity.hot_path_parity(out, pipe.heatmap.last_anchors[0], idx, feat[ii].cpu(), bfb[ii].cpu(), sd_h, sd_p,
                                 calib.load_calibration(None), "ego4view_rw", ctm[idx])
    _check(res, precision, False)


@pytest.mark.parametrize("precision", ["fp16", "bf16", "fp32"])
def test_config3_b1024_pose3d(precision):
    """config 3: standalone 3D lifting from NCHW fp32 feature maps, batch 1024 (split-K of Linear(32768 -> 2048) is off at
    this size, the conv stages walk 20+ tiles per CTA)"""
    from egorear_b200 import calib
    from oracle import parity
    from test_oracle_model import build_pose3d
    dev = torch.device("cuda", 0)
    B = 1024
    g = torch.Generator(device=dev).manual_seed(23)
    fi = torch.relu(torch.randn((B, 4, 128, 64, 64), generator=g, device=dev))
    ff = torch.relu(torch.randn((B, 4, 128, 64, 64), generator=g, device=dev))
    m = build_pose3d("ego4view_syn", precision).to(dev)
    with torch.no_grad():
        preds = m(fi, ff, None)
    idx = parity.sample_indices(B, N_SAMPLE, seed=3)
    ii = torch.as_tensor(idx, device=dev)
    sd_p = {k: v.cpu() for k, v in m.state_dict().items()}
    res = parity.pose3d_parity(preds[-1], idx, fi[ii].cpu(), ff[ii].cpu(), sd_p, calib.load_calibration(None), "ego4view_syn")
    print("parity pose3d %s: %s" % (precision, json.dumps(res)))
    assert res["mpjpe_delta_mm"] < MPJPE_MM["fp16"], res
