"""Training-side operators: what `run.py fit` needs from the extension.

  MultiScaleDeformableAttnFunction   autograd.Function with mmcv's call signature — forward `egr_msda_forward`, backward
                                     `egr_msda_backward` (grad of value / sampling locations / attention weights): the
                                     as-executed MSDeformAttn module (modules.py) trains through it
                                     (reference call site: pose_estimation/models/utils/deform_attn.py:155-162)
  mse_loss / MSELoss                 nn.MSELoss(reduction="mean") of the heatmap wrappers (heatmap_mvf_ex.py:258-261)
  mpjpe_loss / MpjpeLoss             pose_estimation/models/utils/pose_metric.py:10-16

The conv / Linear stacks of the hot path train through PyTorch's own autograd when the reference classes are kept for
`fit` (egorear_b200.patch() swaps inference classes only): these operators cover the one op PyTorch
does not have (the deformable attention) and the two losses.  Everything is fp32 and deterministic except the atomic
scatter into grad_value (order of fp32 additions), like mmcv's own backward.
"""
import ctypes

import torch

from . import _lib
from .ops import _on, _ptr, _stream


class MultiScaleDeformableAttnFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights, im2col_step):
        if not value.is_cuda:
            raise RuntimeError("egorear_b200.train: expected CUDA tensors (there is no CPU fallback)")
        B, L, nh, hd = value.shape
        _, Q, _, n_levels, P, _ = sampling_locations.shape
        if n_levels != 1:
            raise NotImplementedError("egorear_b200: single-level deformable attention only (n_levels=%d)" % n_levels)
        H, W = int(value_spatial_shapes[0][0]), int(value_spatial_shapes[0][1])
        assert H * W == L
        v = value.detach().float().contiguous()
        loc = sampling_locations.detach().float().contiguous()
        aw = attention_weights.detach().float().contiguous()
        out = torch.empty((B, Q, nh * hd), dtype=torch.float32, device=v.device)
        with _on(v):
            _lib.check(_lib.load().egr_msda_forward(_ptr(v), B, H, W, nh, hd, _ptr(loc), _ptr(aw), Q, P, _ptr(out), _stream(v)))
        ctx.save_for_backward(v, loc, aw)
        ctx.geom = (B, H, W, nh, hd, Q, P)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        v, loc, aw = ctx.saved_tensors
        B, H, W, nh, hd, Q, P = ctx.geom
        go = grad_output.detach().float().contiguous()
        gv, gl, ga = torch.empty_like(v), torch.empty_like(loc), torch.empty_like(aw)
        with _on(v):
            _lib.check(_lib.load().egr_msda_backward(_ptr(v), B, H, W, nh, hd, _ptr(loc), _ptr(aw), Q, P, _ptr(go), _ptr(gv),
                                                     _ptr(gl), _ptr(ga), _stream(v)))
        return gv, None, None, gl, ga, None


def _ws(device):
    n = int(_lib.load().egr_loss_workspace_bytes())
    return torch.empty(n, dtype=torch.uint8, device=device)


class _MSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target):
        p, t = pred.detach().float().contiguous(), target.detach().float().contiguous()
        loss = torch.empty((), dtype=torch.float32, device=p.device)
        ws = _ws(p.device)
        with _on(p):
            _lib.check(_lib.load().egr_mse_loss_forward(_ptr(p), _ptr(t), p.numel(), _ptr(loss), _ptr(ws), _stream(p)))
        ctx.save_for_backward(p, t)
        ctx.shape = pred.shape
        return loss

    @staticmethod
    def backward(ctx, g):
        p, t = ctx.saved_tensors
        gp = torch.empty_like(p)
        gg = g.detach().float().reshape(1).contiguous()
        with _on(p):
            _lib.check(_lib.load().egr_mse_loss_backward(_ptr(p), _ptr(t), _ptr(gg), p.numel(), _ptr(gp), _stream(p)))
        return gp.view(ctx.shape), None


class _MPJPE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, gt):
        p, t = pred.detach().float().contiguous(), gt.detach().float().contiguous()
        D = p.shape[-1]
        loss = torch.empty((), dtype=torch.float32, device=p.device)
        ws = _ws(p.device)
        with _on(p):
            _lib.check(_lib.load().egr_mpjpe_loss_forward(_ptr(p), _ptr(t), p.numel() // D, D, _ptr(loss), _ptr(ws), _stream(p)))
        ctx.save_for_backward(p, t)
        ctx.shape = pred.shape
        return loss

    @staticmethod
    def backward(ctx, g):
        p, t = ctx.saved_tensors
        D = p.shape[-1]
        gp = torch.empty_like(p)
        gg = g.detach().float().reshape(1).contiguous()
        with _on(p):
            _lib.check(_lib.load().egr_mpjpe_loss_backward(_ptr(p), _ptr(t), _ptr(gg), p.numel() // D, D, _ptr(gp), _stream(p)))
        return gp.view(ctx.shape), None


def mse_loss(pred, target):
    """nn.MSELoss()(pred, target) — mean of squared differences (the wrappers' `criteria` for heatmaps)"""
    if pred.shape != target.shape:
        raise RuntimeError("mse_loss: shapes differ (%s vs %s)" % (tuple(pred.shape), tuple(target.shape)))
    return _MSE.apply(pred, target)


def mpjpe_loss(pred_pose, gt_pose):
    """MpjpeLoss.forward: mean over all joints of ||gt - pred||_2 (models/utils/pose_metric.py:10-16)"""
    if pred_pose.shape != gt_pose.shape:
        raise RuntimeError("mpjpe_loss: shapes differ (%s vs %s)" % (tuple(pred_pose.shape), tuple(gt_pose.shape)))
    return _MPJPE.apply(pred_pose, gt_pose)


class MSELoss(torch.nn.Module):
    def forward(self, pred, target):
        return mse_loss(pred, target)


class MpjpeLoss(torch.nn.Module):
    def forward(self, pred_pose, gt_pose):
        return mpjpe_loss(pred_pose, gt_pose)
