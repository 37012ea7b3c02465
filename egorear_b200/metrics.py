"""Eval-time metrics of the reference's Lightning wrappers on the GPU.

Same names, argument meaning and return types as the reference:

  compute_mpjpe_batch / compute_pck_3d_batch / compute_auc_3d_batch   <- pose_estimation/utils/loss.py:9-48
  batch_compute_similarity_transform_numpy                            <- models/utils/pose_metric.py:104-116
  evaluate_pose(pred_pose, gt_pose, prefix)            <- pl_wrappers/egoposeformer/pose_3d_mvf_ex.py:317-333
  evaluate_heatmap(pred_heatmap, gt_heatmap, prefix)   <- pl_wrappers/egoposeformer/pose_3d_mvf_ex.py:335-360
  evaluate(pred_heatmap, gt_heatmap, prefix)           <- pl_wrappers/egoposeformer/heatmap_mvf_ex.py:263-299

The reference loops over the batch in Python (one numpy SVD per sample, one `.cpu()` sync per sample and view for the
masked L1); here each call is one or three kernel launches through libegorear_b200.so and ONE device->host copy.
There is no CPU path.
"""
import ctypes
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from .ops import _need_cuda, _ptr, _stream

CM2MM = 10                                   # pose_3d_mvf_ex.py:96
_AUC_THRESHOLDS = np.linspace(0, 150, 31).tolist()    # utils/loss.py:40
_thr_cache = {}


def _thresholds(thresholds, device):
    key = (tuple(float(t) for t in thresholds), device)
    t = _thr_cache.get(key)
    if t is None:
        t = torch.tensor(key[0], dtype=torch.float32, device=device)
        _thr_cache[key] = t
    return t


def _poses(pred, gt, name):
    _need_cuda(pred, name)
    _need_cuda(gt, name)
    if pred.ndim != 3 or pred.shape[-1] != 3 or pred.shape != gt.shape:
        raise RuntimeError("egorear_b200.%s: expected two [B, J, 3] tensors of the same shape" % name)
    return pred.detach().float().contiguous(), gt.detach().float().contiguous()


def _pose_metrics(pred, gt, unit_scale, pck_threshold, thresholds, want_s1_hat, name):
    p, g = _poses(pred, gt, name)
    B, J, _ = p.shape
    thr = _thresholds(thresholds, p.device)
    metrics = torch.empty((B, 4), dtype=torch.float64, device=p.device)
    s1_hat = torch.empty((B, J, 3), dtype=torch.float64, device=p.device) if want_s1_hat else None
    _lib.check(_lib.load().egr_eval_pose(_ptr(p), _ptr(g), B, J, float(unit_scale), float(pck_threshold), _ptr(thr),
                                         int(thr.numel()), _ptr(metrics), _ptr(s1_hat), _stream()))
    return metrics, s1_hat


def compute_mpjpe_batch(pred_keypoints, gt_keypoints):
    """utils/loss.py:9-12 -> [B] float32"""
    m, _ = _pose_metrics(pred_keypoints, gt_keypoints, 1.0, 150.0, (), False, "compute_mpjpe_batch")
    return m[:, 0].float()


def compute_pck_3d_batch(pred_keypoints, gt_keypoints, threshold=150):
    """utils/loss.py:32-37 -> [B] float32 in [0, 1]"""
    m, _ = _pose_metrics(pred_keypoints, gt_keypoints, 1.0, threshold, (), False, "compute_pck_3d_batch")
    return m[:, 2].float() / 100.0


def compute_auc_3d_batch(pred_keypoints, gt_keypoints, thresholds=_AUC_THRESHOLDS):
    """utils/loss.py:40-48 -> [B] float32 in [0, 1]"""
    m, _ = _pose_metrics(pred_keypoints, gt_keypoints, 1.0, 150.0, thresholds, False, "compute_auc_3d_batch")
    return m[:, 3].float() / 100.0


def batch_compute_similarity_transform_numpy(S1, S2):
    """models/utils/pose_metric.py:104-116: similarity-aligned S1 -> [B, J, 3] float64 on S1's device (the reference
    returns the float64 numpy result moved back to the device)."""
    _, s1_hat = _pose_metrics(S1, S2, 1.0, 150.0, (), True, "batch_compute_similarity_transform_numpy")
    return s1_hat


def evaluate_pose(pred_pose, gt_pose, prefix, cm2mm=CM2MM):
    """pose_3d_mvf_ex.py:317-333: OrderedDict of per-sample numpy arrays (mm / percent), one kernel + one D2H copy."""
    m, _ = _pose_metrics(pred_pose, gt_pose, cm2mm, 150.0, _AUC_THRESHOLDS, False, "evaluate_pose")
    host = m.cpu().numpy()
    metrics = OrderedDict()
    metrics[prefix + "_mpjpe"] = host[:, 0].astype(np.float32)
    metrics[prefix + "_pa_mpjpe"] = host[:, 1].copy()
    metrics[prefix + "_pck_3d"] = host[:, 2].astype(np.float32)
    metrics[prefix + "_auc_3d"] = host[:, 3].astype(np.float32)
    return metrics


# ------------------------------------------------------------------------------------------------
# heatmaps
# ------------------------------------------------------------------------------------------------
_ws_cache = {}


def _frames(t, name):
    """[B,V,C,H,W] fp32 with a contiguous [V,C,H,W] block per frame -> (tensor, batch stride); view slices pass as is"""
    _need_cuda(t, name)
    if t.ndim != 5:
        raise RuntimeError("egorear_b200.%s: expected [B, V, C, H, W] heatmaps" % name)
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    B, V, C, H, W = t.shape
    inner_ok = t.stride()[1:] == (C * H * W, H * W, W, 1)
    bs = t.stride(0) if B > 1 else V * C * H * W
    if not inner_ok or bs < V * C * H * W or bs % 4 or t.data_ptr() % 16:
        t = t.contiguous()
        bs = V * C * H * W
    return t, bs


def _heatmap_metrics(pred_heatmap, gt_heatmap, threshold, name):
    p, p_bs = _frames(pred_heatmap, name)
    g, g_bs = _frames(gt_heatmap, name)
    if p.shape != g.shape:
        raise RuntimeError("egorear_b200.%s: prediction and target shapes differ" % name)
    B, V, C, H, W = p.shape
    lib = _lib.load()
    need = int(lib.egr_eval_heatmap_workspace_bytes(B, V, C))
    key = (p.device, torch.cuda.current_stream().cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(max(need, 1 << 20), dtype=torch.uint8, device=p.device)
        _ws_cache[key] = ws
    per_frame = torch.empty((2, B), dtype=torch.float32, device=p.device)
    scalars = torch.full((2,), float("nan"), dtype=torch.float32, device=p.device) if B == 0 else \
        torch.empty((2,), dtype=torch.float32, device=p.device)          # nn.MSELoss of nothing is NaN
    _lib.check(lib.egr_eval_heatmap(_ptr(p), p_bs, _ptr(g), g_bs, B, V, C, H, W, float(threshold), _ptr(per_frame[0]),
                                    _ptr(per_frame[1]), _ptr(scalars), _ptr(ws), ctypes.c_size_t(ws.numel()), _stream()))
    return per_frame, scalars


def evaluate_heatmap(pred_heatmap, gt_heatmap, prefix):
    """pose_3d_mvf_ex.py:335-360: {prefix_l1_error_heatmap, prefix_pos_l1_error_heatmap}, [B] CPU tensors."""
    per_frame, _ = _heatmap_metrics(pred_heatmap, gt_heatmap, 1.0, "evaluate_heatmap")
    host = per_frame.cpu()
    metrics = OrderedDict()
    metrics[prefix + "_l1_error_heatmap"] = host[0]
    metrics[prefix + "_pos_l1_error_heatmap"] = host[1]
    return metrics


def evaluate(pred_heatmap, gt_heatmap, prefix, threshold=1.0):
    """heatmap_mvf_ex.py:263-299: the two per-sample L1 errors ([B] CPU tensors) plus mse_heatmap and mse_pts2d
    (0-dim tensors on the device, like the reference's nn.MSELoss results)."""
    per_frame, scalars = _heatmap_metrics(pred_heatmap, gt_heatmap, threshold, "evaluate")
    host = per_frame.cpu()
    metrics = OrderedDict()
    metrics[prefix + "_l1_error_heatmap"] = host[0]
    metrics[prefix + "_pos_l1_error_heatmap"] = host[1]
    metrics[prefix + "_mse_heatmap"] = scalars[0]
    metrics[prefix + "_mse_pts2d"] = scalars[1]
    return metrics
