"""ORACLE-side parity checker (test infrastructure, NOT product code).

Compares the outputs of a `HotPathPipeline` forward on a FULL batch with `oracle/model_ref.py` run on a seeded SAMPLE
of its frames (frames are independent, so the oracle only has to run the sample).  Used by tests/test_gpu_parity_bench.py
at the benchmarked batch sizes, by `__graft_entry__.smoke()` and by bench.py's untimed `"parity"` block — always as the
checker, never on a measured or shipped path.

What is reported (reference lines: decode `utils/loss.py:122-142`; mvfex forward `egoposeformer_heatmap_mvf_ex.py:236-437`;
chained model `egoposeformer_mvf_ex.py:422-452`):

  hm_init_rel / hm_refined_rel / feat_refined_rel   max|gpu - ref| / max|ref| with the oracle driven by the SAME anchors the
                                                    GPU used (its own init-heatmap decode is handed to the oracle through
                                                    `heatmap_for_anchor`, :293-296), so the bound is on the arithmetic
  anchors_same_frac, anchors_max_px                 decoded init-heatmap argmax cells (the cross-attention anchors) vs the
                                                    oracle decoding ITS OWN init heatmap: fraction of identical cells, and the
                                                    largest displacement (pixels, Chebyshev) among the rest
  joints2d_same_frac, joints2d_max_px               the same for the final decode of the refined heatmap (what the model
                                                    outputs as 2D joints), oracle fully on its own anchors = end-to-end
  mpjpe_delta_mm                                    3D joints, GPU chain vs oracle chain (oracle features -> oracle lifting)
"""
import numpy as np
import torch

from . import model_ref


def sample_indices(B, n, seed=0):
    n = min(n, B)
    idx = np.random.default_rng(seed).choice(B, size=n, replace=False)
    idx.sort()
    return [int(i) for i in idx]


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / max(float(b.double().abs().max()), 1e-30))


def _cells(a, b):
    """a, b [..., 2] integer-valued pixel coordinates -> (fraction of identical cells, max Chebyshev displacement)"""
    d = (a.double() - b.double()).abs().amax(dim=-1)
    return float((d == 0).double().mean()), float(d.max())


@torch.no_grad()
def hot_path_parity(out, anchors_2d, idx, feat_s, bfb_s, sd_h, sd_p, cams, camera_model="ego4view_syn", ctm_s=None):
    """out: HotPathPipeline.forward(...) of the full batch; anchors_2d: pipe.heatmap.last_anchors[0] ([B,V,J,2], normalised);
    idx: sampled frame indices; feat_s / bfb_s / ctm_s: CPU fp32 copies of the sampled frames' inputs; sd_*: CPU state dicts."""
    ii = torch.as_tensor(idx)
    hm0 = out["list_hm"][0][ii.to(out["list_hm"][0].device)].float().cpu()
    hm1 = out["list_hm"][-1][ii.to(out["list_hm"][-1].device)].float().cpu()
    ff1 = out["list_ff"][-1]
    ff1 = ff1[ii.to(ff1.device)].float().cpu() if ff1 is not None else None
    a_gpu = anchors_2d[ii.to(anchors_2d.device)].float().cpu()
    j2_gpu = out["joints2d"][ii.to(out["joints2d"].device)].float().cpu()
    p3_gpu = out["pose3d"][ii.to(out["pose3d"].device)].float().cpu()
    n, V, J, H, W = hm1.shape
    # (1) the oracle end to end on its own anchors
    lh_o, lf_o, a_o, _ = model_ref.mvfex_hot_path(sd_h, feat_s, bfb_s)
    res = {"frames": len(idx), "batch": int(out["list_hm"][0].shape[0])}
    res["anchors_same_frac"], px = _cells(a_gpu * W, a_o * W)
    res["anchors_max_px"] = px
    j2_o, _, _ = model_ref.get_max_preds(lh_o[-1].reshape(n * V, J, H, W), 0.5, False)
    res["joints2d_same_frac"], res["joints2d_max_px"] = _cells(j2_gpu.reshape(-1, 2), j2_o.reshape(-1, 2))
    # (2) arithmetic bound: the oracle on the anchors the GPU used (skipped when they are identical anyway)
    if res["anchors_same_frac"] < 1.0:
        lh_a, lf_a, _, _ = model_ref.mvfex_hot_path(sd_h, feat_s, bfb_s, heatmap_for_anchor=hm0)
    else:
        lh_a, lf_a = lh_o, lf_o
    res["hm_init_rel"] = _rel(hm0, lh_o[0])
    res["hm_refined_rel"] = _rel(hm1, lh_a[-1])
    if ff1 is not None:
        res["feat_refined_rel"] = _rel(ff1, lf_a[-1])
    # (3) the chained 3D joints: oracle features -> oracle lifting
    p3_o = model_ref.pose3d_forward(sd_p, feat_s, lf_a[-1], cams, camera_model, ctm_s)[-1]
    res["mpjpe_delta_mm"] = 10.0 * float((p3_gpu - p3_o).norm(dim=-1).mean(dim=-1).max())
    return {k: (round(v, 6) if isinstance(v, float) else v) for k, v in res.items()}


@torch.no_grad()
def pose3d_parity(preds_last, idx, feats_init_s, feats_final_s, sd_p, cams, camera_model="ego4view_syn", ctm_s=None):
    """standalone lifting (BASELINE config 3): preds_last [B,16,3] of the full batch vs the oracle on the sampled frames"""
    ii = torch.as_tensor(idx).to(preds_last.device)
    got = preds_last[ii].float().cpu()
    want = model_ref.pose3d_forward(sd_p, feats_init_s, feats_final_s, cams, camera_model, ctm_s)[-1]
    return {"frames": len(idx), "batch": int(preds_last.shape[0]),
            "mpjpe_delta_mm": round(10.0 * float((got - want).norm(dim=-1).mean(dim=-1).max()), 6)}
