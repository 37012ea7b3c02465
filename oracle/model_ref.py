"""ORACLE (test infrastructure, NOT product code).

Plain PyTorch fp32, state-dict driven, functional restatement of the floating-point part of the
EgoRear hot path: the mvfex/jqa heatmap refinement, the heatmap heads, the argmax decode and
the pose3d lifting.  Nothing here is imported by the shipped package (egorear_b200/); only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.

Every function cites the reference lines it restates (paths relative to /root/reference).
The restatement is *as executed by the reference* (dense value_proj, dense memory projection,
per-view Python loop) — the CUDA path folds some of those maps, this file deliberately does not.

Parity pinning: the reference has no tests or golden vectors.  The oracle is pinned
against the reference modules themselves, imported live in the build container by
tests/golden/make_golden.py (same deterministic state-dict, same inputs); the resulting vectors are
committed under tests/golden/ and checked by tests/test_oracle_model.py.  The deformable-attention
op (mmcv==2.2.0 `ms_deform_attn_forward`, un-vendored: README.md:134, call site
pose_estimation/models/utils/deform_attn.py:155-162) is restated from its published per-thread
arithmetic in `ms_deform_attn`.
"""
import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

VIEW_NAMES_4 = ["front_left", "front_right", "back_left", "back_right"]


# ----------------------------------------------------------------------------------------------
# small layer helpers
# ----------------------------------------------------------------------------------------------
def _lin(sd: SD, p: str, x: Tensor) -> Tensor:
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def _conv(sd: SD, p: str, x: Tensor, stride: int = 1, pad: int = 0) -> Tensor:
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride, padding=pad)


def _ln(sd: SD, p: str, x: Tensor) -> Tensor:
    # nn.LayerNorm default eps 1e-5, biased variance
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def _up2(x: Tensor) -> Tensor:
    # nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)


# ----------------------------------------------------------------------------------------------
# D1: decode                                                       utils/loss.py:122-142
# ----------------------------------------------------------------------------------------------
def get_max_preds(heatmaps: Tensor, threshold: float = 0.5, normalize: bool = False):
    assert heatmaps.ndim == 4
    B, J, H, W = heatmaps.shape
    flat = heatmaps.reshape(B, J, -1)
    maxvals, idx = torch.max(flat, dim=2, keepdim=True)          # first index on ties
    preds = torch.tile(idx, (1, 1, 2)).float()
    preds[:, :, 0] = preds[:, :, 0] % W
    preds[:, :, 1] = preds[:, :, 1] // W
    if normalize:
        preds[:, :, 0] = preds[:, :, 0] / W
        preds[:, :, 1] = preds[:, :, 1] / H
    valid = maxvals >= threshold
    return preds, maxvals.squeeze(), valid.squeeze()


def get_max_preds_soft_pytorch(batch_heatmaps: Tensor, normalize: bool = False):
    """soft-argmax decoder, utils/loss.py:145-177: softmax over H*W, marginals, expected column / row index"""
    assert batch_heatmaps.ndim == 4
    B, J, H, W = batch_heatmaps.shape
    flat = batch_heatmaps.reshape(B, J, -1)
    maxvals = flat.max(dim=2)[0].view(B, J, 1)
    p = F.softmax(flat, dim=2).view(B, J, H, W)
    x = (p.sum(dim=2) * torch.arange(W, dtype=torch.float32).view(1, 1, W)).sum(dim=2, keepdim=True)
    y = (p.sum(dim=3) * torch.arange(H, dtype=torch.float32).view(1, 1, H)).sum(dim=2, keepdim=True)
    if normalize:
        x, y = x / W, y / H
    return torch.cat([x, y], dim=2), maxvals


def integrate_tensor_2d(heatmaps: Tensor, softmax: bool = True, multiplier: float = 100.0):
    """integral decoder, utils/util.py:80-109: softmax (or relu) of hm * multiplier, marginal masses times the pixel
    index; the relu variant divides by the total mass.  Returns (coordinates [B,J,2], normalised heatmaps)."""
    B, J, H, W = heatmaps.shape
    v = (heatmaps * multiplier).reshape(B, J, -1)
    p = (F.softmax(v, dim=2) if softmax else F.relu(v)).reshape(B, J, H, W)
    mass_x, mass_y = p.sum(dim=2), p.sum(dim=3)
    x = (mass_x * torch.arange(W, dtype=torch.float32)).sum(dim=2, keepdim=True)
    y = (mass_y * torch.arange(H, dtype=torch.float32)).sum(dim=2, keepdim=True)
    if not softmax:
        x = x / mass_x.sum(dim=2, keepdim=True)
        y = y / mass_y.sum(dim=2, keepdim=True)
    return torch.cat((x, y), dim=2).reshape(B, J, 2), p


# ----------------------------------------------------------------------------------------------
# mmcv ms_deform_attn_forward, single level            models/utils/deform_attn.py:155-162
# ----------------------------------------------------------------------------------------------
def ms_deform_attn(value: Tensor, H: int, W: int, loc: Tensor, aw: Tensor) -> Tensor:
    """value [B, H*W, nh, hd]; loc [B, Q, nh, 1, P, 2] (x, y) in [0,1]; aw [B, Q, nh, 1, P].

    Literal restatement of mmcv's `ms_deformable_im2col_gpu_kernel` arithmetic (mmcv 2.2.0):
    h_im = loc_y*H - 0.5, w_im = loc_x*W - 0.5; a point contributes only if
    h_im > -1 and w_im > -1 and h_im < H and w_im < W; each of the 4 bilinear corners is
    dropped when it falls outside [0,H)x[0,W).  Output [B, Q, nh*hd], channel = head*hd + c.
    """
    B, L, nh, hd = value.shape
    Q, P = loc.shape[1], loc.shape[4]
    assert L == H * W and loc.shape[3] == 1
    x = loc[:, :, :, 0, :, 0] * W - 0.5                       # [B,Q,nh,P]
    y = loc[:, :, :, 0, :, 1] * H - 0.5
    x0 = torch.floor(x)
    y0 = torch.floor(y)
    lx, ly = x - x0, y - y0
    x0 = x0.long()
    y0 = y0.long()
    inside = (y > -1) & (x > -1) & (y < H) & (x < W)
    out = value.new_zeros(B, Q, nh, hd)
    vflat = value.permute(0, 2, 1, 3)                         # [B,nh,L,hd]
    a = aw[:, :, :, 0, :]                                     # [B,Q,nh,P]
    for dy, dx, wgt in ((0, 0, (1 - ly) * (1 - lx)), (0, 1, (1 - ly) * lx),
                        (1, 0, ly * (1 - lx)), (1, 1, ly * lx)):
        yy, xx = y0 + dy, x0 + dx
        ok = inside & (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        lin = (yy.clamp(0, H - 1) * W + xx.clamp(0, W - 1))   # [B,Q,nh,P]
        lin_h = lin.permute(0, 2, 1, 3).reshape(B, nh, Q * P)
        g = torch.gather(vflat, 2, lin_h[..., None].expand(B, nh, Q * P, hd))
        g = g.reshape(B, nh, Q, P, hd).permute(0, 2, 1, 3, 4)   # [B,Q,nh,P,hd]
        coef = (wgt * a * ok.to(value.dtype))[..., None]
        out = out + (g * coef).sum(dim=3)
    return out.reshape(B, Q, nh * hd)


def msda_module(sd: SD, p: str, query: Tensor, ref_pts: Tensor, mem: Tensor, H: int, W: int,
                n_heads: int, n_points: int = 16) -> Tensor:
    """MSDeformAttn.forward with one level                models/utils/deform_attn.py:90-168."""
    N, Q, C = query.shape
    value = _lin(sd, p + ".value_proj", mem).view(N, H * W, n_heads, C // n_heads)       # :116-120
    off = _lin(sd, p + ".sampling_offsets", query).view(N, Q, n_heads, 1, n_points, 2)   # :122-124
    aw = _lin(sd, p + ".attention_weights", query).view(N, Q, n_heads, n_points)         # :125-127
    aw = F.softmax(aw, -1).view(N, Q, n_heads, 1, n_points)                              # :128-130
    normalizer = torch.tensor([W, H], dtype=query.dtype, device=query.device)            # :133-135 (x: W, y: H)
    loc = ref_pts[:, :, None, :, None, :] + off / normalizer[None, None, None, None, None, :]  # :136-139
    out = ms_deform_attn(value.float(), H, W, loc, aw)                                   # :155-162
    return _lin(sd, p + ".output_proj", out)                                             # :164


def mha(sd: SD, p: str, x: Tensor, n_heads: int) -> Tensor:
    """SpatialMHA / EgoformerSpatialMHA forward (bias=None)
    estimator/egoposeformer_heatmap_mvf_ex.py:799-817, estimator/egoposeformer_mvf_ex.py:481-498."""
    B, J, C = x.shape
    hd = C // n_heads
    q = _lin(sd, p + ".q_proj", x).reshape(B, J, n_heads, hd).permute(0, 2, 1, 3)
    k = _lin(sd, p + ".k_proj", x).reshape(B, J, n_heads, hd).permute(0, 2, 1, 3)
    v = _lin(sd, p + ".v_proj", x).reshape(B, J, n_heads, hd).permute(0, 2, 1, 3)
    attn = (q @ k.transpose(-2, -1)) * (hd ** -0.5)
    attn = attn.softmax(dim=-1)
    o = (attn @ v).permute(0, 2, 1, 3).reshape(B, J, C)
    return _lin(sd, p + ".out_proj", o)


def transformer_layer(sd: SD, p: str, x: Tensor, mem: Tensor, anchors_2d: Tensor, anchors_valid: Tensor,
                      H: int, W: int, n_heads: int) -> Tensor:
    """MultiViewTransformerLayer.forward / EgoPoseFormerTransformerLayer.forward
    estimator/egoposeformer_heatmap_mvf_ex.py:874-935, estimator/egoposeformer_mvf_ex.py:546-588.
    mem [B,V,HW,C]; anchors_2d [B,V,J,2]; anchors_valid [B,V,J] bool."""
    V = mem.shape[1]
    identity = x
    per_view = []
    for i in range(V):
        r = msda_module(sd, p + ".cross_attn", x, anchors_2d[:, i].reshape(x.shape[0], -1, 1, 2),
                        mem[:, i], H, W, n_heads)
        r = r.masked_fill(~anchors_valid[:, i][..., None].expand_as(r), 0.0)     # :910 / :563
        per_view.append(r)
    x = identity + _lin(sd, p + ".fuse_mlp", torch.cat(per_view, dim=-1))       # :913-915
    x = _ln(sd, p + ".norm_cross", x)
    x = _ln(sd, p + ".norm_spatial", x + mha(sd, p + ".spatial_attn", x, n_heads))   # :861-871
    # FFN models/utils/transformer.py:8-33: Linear, GELU(erf), Dropout(0), Linear, Dropout(0)
    h = F.gelu(_lin(sd, p + ".ffn.layers.0.0", x))
    h = _lin(sd, p + ".ffn.layers.1", h)
    x = _ln(sd, p + ".norm_ffn", x + h)                                          # :919-922
    return x


# ----------------------------------------------------------------------------------------------
# heatmap heads
# ----------------------------------------------------------------------------------------------
def head5(sd: SD, p: str, x: Tensor) -> Tensor:
    """conv_heatmap_layers_stereo_{front,back}   egoposeformer_heatmap_mvf_ex.py:101-126 (H1)."""
    x = F.relu(_conv(sd, p + ".0", x))
    x = F.relu(_conv(sd, p + ".2", x, 2, 1))
    x = F.relu(_conv(sd, p + ".4", x))
    x = _up2(x)
    x = F.relu(_conv(sd, p + ".7", x))
    return _conv(sd, p + ".9", x)


def head4(sd: SD, p: str, x: Tensor) -> Tensor:
    """HeatmapMVF.conv_heatmap_layers[idx]       egoposeformer_heatmap_mvf_ex.py:571-584 (H2)."""
    x = F.relu(_conv(sd, p + ".0", x, 2, 1))
    x = F.relu(_conv(sd, p + ".2", x))
    x = _up2(x)
    x = F.relu(_conv(sd, p + ".5", x))
    return _conv(sd, p + ".7", x)


def init_heatmaps(sd: SD, feat: Tensor) -> Tensor:
    """forward_heatmap_estimation_from_feat       egoposeformer_heatmap_mvf_ex.py:177-210."""
    B, V, C, H, W = feat.shape
    if V == 2:
        return head5(sd, "conv_heatmap_layers_stereo_front", feat.reshape(B * 2, C, H, W)).view(B, 2, -1, H, W)
    front = head5(sd, "conv_heatmap_layers_stereo_front", feat[:, 0:2].reshape(B * 2, C, H, W))
    back = head5(sd, "conv_heatmap_layers_stereo_back", feat[:, 2:].reshape(B * (V - 2), C, H, W))
    return torch.cat((front.view(B, 2, -1, H, W), back.view(B, V - 2, -1, H, W)), dim=1)


# ----------------------------------------------------------------------------------------------
# HeatmapMVF.forward (jqa variant)                egoposeformer_heatmap_mvf_ex.py:652-731
# ----------------------------------------------------------------------------------------------
def heatmap_mvf(sd: SD, p: str, heatmap: Tensor, frame_feat: Tensor, feat_mv: Tensor, anchors_2d: Tensor,
                anchors_valid: Tensor, bfb: Tensor, n_heads: int = 4, stages: Optional[dict] = None):
    B, V, C, H, W = feat_mv.shape
    # Q1 jqa query :655-665
    hm_embed = _lin(sd, p + ".heatmap_proj.2", F.relu(_lin(sd, p + ".heatmap_proj.0", heatmap.reshape(B, -1, H * W))))
    g = _lin(sd, p + ".fc_bfb", F.adaptive_avg_pool2d(bfb, (1, 1)).view(B, -1)).unsqueeze(1)
    x = F.relu(_lin(sd, p + ".fc_query.0", sd[p + ".joint_query_embed.weight"].unsqueeze(0) + g + hm_embed))
    if stages is not None:
        stages["query"] = x
    # M1 memory :690-693
    mem = _conv(sd, p + ".frame_feat_multi_view_proj", feat_mv.reshape(B * V, C, H, W))
    E = mem.shape[1]
    mem = mem.view(B, V, E, H * W).permute(0, 1, 3, 2) + sd[p + ".frame_feat_multi_view_pos_embed"]
    # F1 :695, :525-532
    ff = F.relu(_conv(sd, p + ".frame_feat_proj_layers.0", frame_feat))
    ff = F.relu(_conv(sd, p + ".frame_feat_proj_layers.2", ff, 2, 1))
    ff = F.relu(_conv(sd, p + ".frame_feat_proj_layers.4", ff))
    if stages is not None:
        stages["frame_feat_sq"] = ff
    # A1-A3 (num_former_layers == 1 in every shipped config; loop kept for generality) :699-705
    n_layers = 0
    while (p + ".transformer_layers.%d.fuse_mlp.weight" % n_layers) in sd:
        n_layers += 1
    hms, feats = [], []
    for li in range(n_layers):
        x = transformer_layer(sd, p + ".transformer_layers.%d" % li, x, mem, anchors_2d, anchors_valid, H, W, n_heads)
        if stages is not None:
            stages["tokens%d" % li] = x
        _x = _ln(sd, p + ".post_norm.%d" % li, x)                                    # :707
        s = int(math.sqrt(_x.shape[-1]))
        _x = _x.view(B, -1, s, s)                                                    # :709-711
        # T1 TransformerHeadLayer :947-954
        o = F.relu(_conv(sd, p + ".head_layers.%d.head.0" % li, _x))
        o = F.relu(_conv(sd, p + ".head_layers.%d.head.3" % li, _up2(o)))
        # R1 :715, :553-563
        r = F.relu(_conv(sd, p + ".frame_feat_refined_proj_layers.%d.0" % li, o + ff))
        r = F.relu(_conv(sd, p + ".frame_feat_refined_proj_layers.%d.3" % li, _up2(r)))
        # H2 :726
        hms.append(head4(sd, p + ".conv_heatmap_layers.%d" % li, r))
        feats.append(r)
    return hms, feats


def mvfex_hot_path(sd: SD, feat: Tensor, bfb: Tensor, heatmap_for_anchor: Optional[Tensor] = None,
                   heatmap_threshold: float = 0.5, n_heads: int = 4, stages: Optional[dict] = None):
    """EgoPoseFormerHeatmapMVFEX.forward after the backbones (use_1by1_conv=False branch)
    egoposeformer_heatmap_mvf_ex.py:262-437.  feat [B,V,128,64,64], bfb [B,V,512,8,8].
    Returns (list_heatmap_pred [init, refined...], list_frame_feat [init, refined...], anchors_2d, anchors_valid)."""
    B, V = feat.shape[:2]
    hm_init = init_heatmaps(sd, feat)
    src = heatmap_for_anchor if isinstance(heatmap_for_anchor, torch.Tensor) else hm_init      # :293-296
    J = src.shape[2]
    pts, maxv, valid = get_max_preds(src.reshape(B * V, J, *src.shape[-2:]), heatmap_threshold, True)
    anchors_2d = pts.view(B, V, J, 2)
    anchors_valid = valid.reshape(B, V, J)
    names = VIEW_NAMES_4[:V] if V != 3 else ["front_left", "front_right", "back"]
    per_view_h, per_view_f = [], []
    for v, nm in enumerate(names):
        st = {} if stages is not None else None
        h, f = heatmap_mvf(sd, "heatmap_refiner_" + nm, hm_init[:, v], feat[:, v], feat, anchors_2d, anchors_valid,
                           bfb[:, v], n_heads, st)
        if stages is not None:
            stages[nm] = st
        per_view_h.append(h)
        per_view_f.append(f)
    list_h, list_f = [hm_init], [feat]
    for li in range(len(per_view_h[0])):
        list_h.append(torch.stack([h[li] for h in per_view_h], dim=1))
        list_f.append(torch.stack([f[li] for f in per_view_f], dim=1))
    return list_h, list_f, anchors_2d, anchors_valid


# ----------------------------------------------------------------------------------------------
# P3: fisheye reprojection                       utils/camera_models.py:53-104
# ----------------------------------------------------------------------------------------------
def world2camera(pts3d: Tensor, cam: dict, camera_model: str, camera_name: str,
                 coord_trans_mat: Optional[Tensor] = None):
    """cam = {"size":[h,w], "image_center":[cx,cy], "polynomialW2C":[...]}.
    NOTE the syn branch mutates `pts3d` IN PLACE exactly like the reference (:57-63)."""
    if camera_model.startswith("ego4view_rw"):
        B, J = pts3d.shape[:2]
        hom = torch.cat([pts3d * 0.01, torch.ones((B, J, 1), dtype=pts3d.dtype, device=pts3d.device)], dim=2)
        M = coord_trans_mat.unsqueeze(1).expand(-1, J, -1, -1)
        p = torch.matmul(M, hom.unsqueeze(3)).squeeze(3)[:, :, :3] * 100.0                  # :55, :187-215
    else:
        offs = {"camera_front_left": [6.0, 0.0, 0.0], "camera_front_right": [-6.0, 0.0, 0.0],
                "camera_back_left": [-6.0, 37.0, 0.0], "camera_back_right": [6.0, 37.0, 0.0]}[camera_name]
        if camera_name in ("camera_back_left", "camera_back_right"):
            pts3d[..., 0:2] *= -1                                                           # :58-61
        pts3d += torch.tensor(offs, dtype=pts3d.dtype, device=pts3d.device)                 # :62
        p = pts3d
    p = p[:, None]
    x, y, z = p[..., 0], p[..., 1], p[..., 2]
    norm = torch.sqrt(x * x + y * y)
    theta = torch.atan(-z / norm)
    poly = torch.tensor(cam["polynomialW2C"], dtype=torch.float32, device=p.device)
    rho = sum(a * theta ** i for i, a in enumerate(poly))                                   # :85 (powers, not Horner)
    cx, cy = torch.tensor(cam["image_center"], dtype=torch.float32, device=p.device)
    size = torch.tensor(cam["size"], device=p.device)
    u = (x / norm * rho + cx) / size[1]
    v = (y / norm * rho + cy) / size[0]
    pt = torch.stack((u, v), dim=-1)
    in_fov = (pt[..., 0] > 0) & (pt[..., 1] > 0) & (pt[..., 0] < 1) & (pt[..., 1] < 1)
    return pt.clamp(min=0.0, max=1.0), in_fov


def reproject(init_anchors_3d: Tensor, cams: Dict[str, dict], camera_model: str,
              coord_trans_mat: Optional[Tensor] = None):
    """_reproject_3d_to_2d                             egoposeformer_mvf_ex.py:340-382."""
    if camera_model in ("ego4view_syn", "ego4view_rw"):
        names = ["camera_front_left", "camera_front_right", "camera_back_left", "camera_back_right"]
    elif camera_model.endswith("stereo_front"):
        names = ["camera_front_left", "camera_front_right"]
    else:
        names = ["camera_back_left", "camera_back_right"]
    a2, av = [], []
    for i, nm in enumerate(names):
        ctm = coord_trans_mat[:, i] if camera_model.startswith("ego4view_rw") else None
        p, f = world2camera(init_anchors_3d, cams[nm], camera_model, nm, ctm)
        a2.append(p)
        av.append(f)
    return torch.cat(a2, dim=1), torch.cat(av, dim=1)


# ----------------------------------------------------------------------------------------------
# EgoPoseFormerPose3D.forward                      egoposeformer_mvf_ex.py:422-452
# ----------------------------------------------------------------------------------------------
def pose3d_forward(sd: SD, feats_init: Tensor, feats_final: Tensor, cams: Dict[str, dict], camera_model: str,
                   coord_trans_mat: Optional[Tensor] = None, use_pred_heatmap_init: bool = True,
                   n_heads: int = 4, stages: Optional[dict] = None) -> List[Tensor]:
    feats = feats_init if use_pred_heatmap_init else feats_final                              # :424-427
    B, V, C, H, W = feats.shape
    mem = _conv(sd, "feat_proj", feats.reshape(B * V, C, H, W))                               # P1 :431
    E = mem.shape[1]
    mem = mem.view(B, V, E, H * W).permute(0, 1, 3, 2)                                        # :394
    # P2 _forward_mlp_conv :309-322, conv_frame_feat :229-239
    x = feats_final.reshape(B * V, C, H, W)
    x = F.relu(_conv(sd, "conv_frame_feat.0", x))
    x = F.relu(_conv(sd, "conv_frame_feat.2", x, 2, 1))
    x = F.max_pool2d(x, 2)
    x = F.relu(_conv(sd, "conv_frame_feat.5", x))
    x = F.relu(_conv(sd, "conv_frame_feat.7", x, 2, 1))
    if stages is not None:
        stages["conv_feat"] = x
    x = x.reshape(B, -1)                                                                      # "(b v) c h w -> b (v c h w)"
    x = F.gelu(_lin(sd, "mlp_pred.0.0", x))
    x = F.gelu(_lin(sd, "mlp_pred.1.0", x))
    mlp_pred_3d = _lin(sd, "mlp_pred.2", x).reshape(B, -1, 3)
    J = mlp_pred_3d.shape[1]
    init_anchors_3d = mlp_pred_3d.clone().detach()                                            # :441
    anchors_2d, anchors_valid = reproject(init_anchors_3d, cams, camera_model, coord_trans_mat)   # P3 (mutates syn)
    if stages is not None:
        stages["anchors_2d"], stages["anchors_valid"] = anchors_2d, anchors_valid
        stages["anchors_3d_after"] = init_anchors_3d.clone()
    joint_inds = (torch.arange(1, J + 1, dtype=mem.dtype, device=mem.device).reshape(1, J, 1).repeat(B, 1, 1)) / float(J)
    q = torch.cat((joint_inds, init_anchors_3d), dim=-1)                                      # :400-406
    q = F.relu(_lin(sd, "query_gen_mlp.0", q))
    q = F.relu(_lin(sd, "query_gen_mlp.2", q))
    q = _lin(sd, "query_gen_mlp.4", q)
    preds = [mlp_pred_3d]
    li = 0
    while ("layers.%d.fuse_mlp.weight" % li) in sd:
        q = transformer_layer(sd, "layers.%d" % li, q, mem, anchors_2d, anchors_valid, H, W, n_heads)
        _x = _ln(sd, "post_norm.%d" % li, q)
        off = _lin(sd, "reg_mlp.%d.2" % li, F.gelu(_lin(sd, "reg_mlp.%d.0" % li, _x)))      # :255-262
        preds.append(off + init_anchors_3d)                                                   # :418
        li += 1
    return preds


# ----------------------------------------------------------------------------------------------
# H1': EgoPoseFormerHeatmap 1x1 head               egoposeformer_heatmap.py:23,34-39
# ----------------------------------------------------------------------------------------------
def heatmap_1x1(sd: SD, feats: Tensor, p: str = "conv_heatmap") -> Tensor:
    B, V = feats.shape[:2]
    h = _conv(sd, p, feats.reshape(B * V, *feats.shape[2:]))
    return h.view(B, V, *h.shape[1:])
