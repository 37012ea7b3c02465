"""Host-side mirror of the reference's module interface for the hot path.

Same class names, constructor signatures (they are splatted from the YAML `model_cfg` dicts), parameter /
buffer names and shapes (so `load_state_dict(strict=True)` of reference checkpoints works), and
the same forward signatures / return structures.  Only the forwards differ: they call the C-ABI engines
(libegorear_b200.so) instead of PyTorch ops.  The ResNet18+FPN backbone stays plain PyTorch (out of scope,
SURVEY §2 #10) and produces the engines' inputs.

Reference files mirrored (paths relative to the EgoRear checkout):
  pose_estimation/models/estimator/egoposeformer_heatmap.py         EgoPoseFormerHeatmap
  pose_estimation/models/estimator/egoposeformer_heatmap_mvf_ex.py  EgoPoseFormerHeatmapMVFEX, HeatmapMVF,
        MultiViewTransformerLayer, DeformMultiViewAttn, SpatialMHA, TransformerHeadLayer
  pose_estimation/models/estimator/egoposeformer_mvf_ex.py          EgoPoseFormerMVFEX, EgoPoseFormerPose3D,
        EgoPoseFormerTransformerLayer, DeformStereoAttn, EgoformerSpatialMHA
  pose_estimation/models/utils/deform_attn.py                       MSDeformAttn
  pose_estimation/models/utils/transformer.py                       FFN, CustomMultiheadAttention
  pose_estimation/models/backbones/resnet.py                        ResnetBackbone (PyTorch, unchanged role)
"""
import copy
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops, torch_ops
from .calib import load_calibration
from .engine import BackboneEngine, MvfexEngine, Pose3DEngine

_VIEWS4 = ("front_left", "front_right", "back_left", "back_right")


class _EngineOwner:
    """Mixin of the three modules that own a C engine (HeatmapMVF standalone, EgoPoseFormerHeatmapMVFEX,
    EgoPoseFormerPose3D): keeps the pointers the engine registered in step with the module's parameters.
    .to() / .cuda() / .float() re-create the tensors -> re-register; load_state_dict writes in place -> re-pack;
    forward under training mode with grad enabled raises instead of failing later with an opaque autograd error
    (the engines are inference-only: they detach everything)."""
    _engine = None

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        # key under which the egr::* operators find this module (weak registry): a plain int, so that a traced forward
        # (torch.compile) reads it as a constant
        self._egr_key = torch_ops.register_module(self)

    def __getstate__(self):
        st = dict(self.__dict__)
        st["_engine"] = None               # a C handle: rebuilt lazily by the copy
        if "_bb_engine" in st:
            st["_bb_engine"] = None
        return st

    def __setstate__(self, state):
        super().__setstate__(state)
        self._engine = None
        self._egr_hooked = False
        self._egr_key = torch_ops.register_module(self)      # copy.deepcopy / pickle: the copy gets its own key and engine

    def _engine_params(self):
        raise NotImplementedError

    def _make_engine(self):
        raise NotImplementedError

    def engine(self):
        if self._engine is None:
            self._engine = self._make_engine()
            self._engine.set_params(self._engine_params())
            if not getattr(self, "_egr_hooked", False):
                def _invalidate(module, incompatible_keys):
                    for e in (module._engine, getattr(module, "_bb_engine", None)):
                        if e is not None:
                            e.invalidate()
                self.register_load_state_dict_post_hook(_invalidate)
                self._egr_hooked = True
        return self._engine

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        if self._engine is not None:       # tensors were re-created (.cuda(), .float(), ...): re-register
            st = self._engine_params()
            if all(v.is_cuda for v in st.values()):
                self._engine.set_params(st)
            else:
                self._engine = None        # moved off the GPU: the engine is rebuilt when it comes back
        return out

    def _check_inference(self):
        if self.training and torch.is_grad_enabled():
            raise RuntimeError(
                "egorear_b200.%s.forward is inference-only (the CUDA engine detaches its inputs and parameters): call "
                ".eval() or run under torch.no_grad().  For `run.py fit` keep the reference classes for this module "
                "(egorear_b200.patch(modules=[...]) without it); the training-side kernels live in egorear_b200.train."
                % type(self).__name__)


def _seq_conv(spec):
    """spec: list of ("conv", cin, cout, k, s, p) | "relu" | "up" | "pool" -> nn.Sequential with the reference's indices."""
    layers = []
    for s in spec:
        if s == "relu":
            layers.append(nn.ReLU(inplace=False))
        elif s == "up":
            layers.append(nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True))
        elif s == "pool":
            layers.append(nn.MaxPool2d(2))
        else:
            _, cin, cout, k, st, p = s
            layers.append(nn.Conv2d(cin, cout, k, st, p))
    return nn.Sequential(*layers)


# ------------------------------------------------------------------------------------------------------
# backbone (PyTorch; produces the hot path's inputs)                  models/backbones/resnet.py:6-152
# ------------------------------------------------------------------------------------------------------
class ResNetTorchvision(nn.Module):
    def __init__(self, model_name, use_imagenet_pretrain, out_stride):
        super().__init__()
        import torchvision
        if model_name != "resnet18":
            raise NotImplementedError("model type [%s] is invalid", model_name)
        net = torchvision.models.resnet18(weights="DEFAULT" if use_imagenet_pretrain else None)
        ch = list(net.children())
        self.layer_s2 = nn.Sequential(*ch[:3])
        self.layer_s4 = nn.Sequential(*ch[3:5])
        self.layer_s8, self.layer_s16, self.layer_s32 = ch[5], ch[6], ch[7]
        self.out_stride = out_stride

    def forward(self, x):
        if x.ndim == 4:
            B, V, H, W = x.shape
            x = x.reshape(B * V, 1, H, W).repeat(1, 3, 1, 1)
        else:
            B, V, C, H, W = x.shape
            x = x.reshape(B * V, C, H, W)
        s2 = self.layer_s2(x)
        s4 = self.layer_s4(s2)
        s8 = self.layer_s8(s4)
        s16 = self.layer_s16(s8)
        s32 = self.layer_s32(s16)
        outs = {4: [s4, s8, s16, s32], 8: [s8, s16, s32], 16: [s16, s32], 32: [s32]}.get(self.out_stride)
        if outs is None:
            raise NotImplementedError
        return [o.reshape(B, V, *o.shape[1:]) for o in outs]


class EfficientFPN(nn.Module):
    def __init__(self, in_channels, out_channels, with_relu=True):
        super().__init__()
        assert isinstance(in_channels, list)
        self.in_channels, self.out_channels, self.with_relu = in_channels, out_channels, with_relu
        self.updample = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)   # (sic) reference attribute name
        act = (lambda: [nn.ReLU(inplace=False)]) if with_relu else (lambda: [])
        self.lateral_convs, self.fuse_convs, self.fpn_convs = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        for i, cin in enumerate(in_channels):
            self.lateral_convs.append(nn.Sequential(nn.Conv2d(cin, out_channels, 1), *act()))
            if i != 0:
                self.fuse_convs.append(nn.Sequential(nn.Conv2d(out_channels * 2, out_channels, 1), *act()))
                self.fpn_convs.append(nn.Sequential(nn.Conv2d(out_channels, out_channels, 3, padding=1), *act()))

    def forward(self, inputs):
        B, V = inputs[0].shape[:2]
        lat = [conv(x.flatten(0, 1)) for conv, x in zip(self.lateral_convs, inputs)]
        for i in range(len(lat) - 1, 0, -1):
            lat[i - 1] = self.fpn_convs[i - 1](self.fuse_convs[i - 1](torch.cat((lat[i - 1], self.updample(lat[i])), 1)))
        return lat[0].reshape(B, V, *lat[0].shape[1:])


class ResnetBackbone(nn.Module):
    def __init__(self, resnet_cfg, neck_cfg):
        super().__init__()
        self.backbone = ResNetTorchvision(**resnet_cfg)
        self.neck = EfficientFPN(**neck_cfg)

    def get_output_channel(self):
        return self.neck.out_channels

    def forward(self, image):
        feats = self.backbone(image)
        return self.neck(feats), feats


# ------------------------------------------------------------------------------------------------------
# H1'  EgoPoseFormerHeatmap                                 estimator/egoposeformer_heatmap.py:8-44
# ------------------------------------------------------------------------------------------------------
class EgoPoseFormerHeatmap(nn.Module):
    def __init__(self, encoder_cfg, num_heatmap, detach_heatmap_feat_init=False, **kwargs):
        super().__init__()
        self.num_heatmap = num_heatmap
        self.detach_heatmap_feat_init = detach_heatmap_feat_init
        self.encoder = ResnetBackbone(**encoder_cfg)
        self.conv_heatmap = nn.Conv2d(self.encoder.get_output_channel(), num_heatmap, 1)

    def forward_backbone(self, img, return_feat=False):
        return self.encoder(img)

    def forward(self, img, return_feat=False):
        B, V = img.shape[:2]
        feats, backbone_feats = self.forward_backbone(img)
        hm = ops.heatmap_head_1x1(feats.reshape(B * V, *feats.shape[2:]), self.conv_heatmap.weight, self.conv_heatmap.bias)
        hm = hm.view(B, V, *hm.shape[1:])
        return (hm, feats, backbone_feats) if return_feat else hm


# ------------------------------------------------------------------------------------------------------
# attention building blocks (parameter containers + as-executed forwards through the CUDA MSDA op)
# ------------------------------------------------------------------------------------------------------
class MSDeformAttn(nn.Module):
    """models/utils/deform_attn.py:25-168.  forward = as executed by the reference, with the mmcv op replaced by
    ops.ms_deform_attn.  (The engines do NOT use this forward: they fold value_proj into the sampling.)"""

    def __init__(self, d_model=256, n_levels=1, n_heads=8, n_points=16):
        super().__init__()
        if d_model % n_heads != 0:
            raise ValueError("d_model must be divisible by n_heads, but got {} and {}".format(d_model, n_heads))
        self.im2col_step = 32
        self.d_model, self.n_levels, self.n_heads, self.n_points = d_model, n_levels, n_heads, n_points
        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)
        self._reset_parameters()

    def _reset_parameters(self):
        # stock init :67-88: zero offset weights, ring-pattern offset bias, uniform attention, xavier projections
        with torch.no_grad():
            self.sampling_offsets.weight.zero_()
            th = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
            g = torch.stack([th.cos(), th.sin()], -1)
            g = (g / g.abs().max(-1, keepdim=True)[0]).view(self.n_heads, 1, 1, 2).repeat(1, self.n_levels, self.n_points, 1)
            g = g * torch.arange(1, self.n_points + 1, dtype=torch.float32).view(1, 1, -1, 1)
            self.sampling_offsets.bias.copy_(g.reshape(-1))
            self.attention_weights.weight.zero_()
            self.attention_weights.bias.zero_()
            nn.init.xavier_uniform_(self.value_proj.weight)
            self.value_proj.bias.zero_()
            nn.init.xavier_uniform_(self.output_proj.weight)
            self.output_proj.bias.zero_()

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index,
                input_padding_mask=None, return_sampled_points=False):
        N, Lq, _ = query.shape
        N, Lin, _ = input_flatten.shape
        assert (input_spatial_shapes[:, 0] * input_spatial_shapes[:, 1]).sum() == Lin
        value = self.value_proj(input_flatten)
        if input_padding_mask is not None:
            value = value.masked_fill(input_padding_mask[..., None], float(0))
        value = value.view(N, Lin, self.n_heads, self.d_model // self.n_heads)
        off = self.sampling_offsets(query).view(N, Lq, self.n_heads, self.n_levels, self.n_points, 2)
        aw = self.attention_weights(query).view(N, Lq, self.n_heads, self.n_levels * self.n_points)
        aw = F.softmax(aw, -1).view(N, Lq, self.n_heads, self.n_levels, self.n_points)
        if reference_points.shape[-1] == 2:
            norm = torch.stack([input_spatial_shapes[..., 1], input_spatial_shapes[..., 0]], -1)
            loc = reference_points[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
        elif reference_points.shape[-1] == 4:
            loc = reference_points[:, :, None, :, None, :2] + off / self.n_points * reference_points[:, :, None, :, None, 2:] * 0.5
        else:
            raise ValueError("Last dim of reference_points must be 2 or 4, but get {} instead.".format(reference_points.shape[-1]))
        if torch.is_grad_enabled() and (value.requires_grad or loc.requires_grad or aw.requires_grad):
            from .train import MultiScaleDeformableAttnFunction as _TrainMSDA     # forward + backward kernels (run.py fit)
            out = _TrainMSDA.apply(value.to(torch.float32), input_spatial_shapes, input_level_start_index, loc, aw, self.im2col_step)
        else:
            out = ops.ms_deform_attn(value.to(torch.float32), input_spatial_shapes, input_level_start_index, loc, aw, self.im2col_step)
        out = self.output_proj(out)
        return (out, loc) if return_sampled_points else out


class _DeformViewAttn(MSDeformAttn):
    def __init__(self, feat_shape, **kwargs):
        super().__init__(d_model=kwargs.pop("embed_dim"), n_heads=kwargs.pop("num_heads"), n_points=16)
        self.register_buffer("spatial_shapes", torch.tensor([[feat_shape[0], feat_shape[1]]], dtype=torch.long))
        self.register_buffer("start_index", torch.tensor([0], dtype=torch.long))

    def forward(self, query, img_feat, anchors_2d):
        B, J, C = query.shape
        out = super().forward(query.reshape(B, J, C), anchors_2d.detach().reshape(B, J, 1, 2), img_feat.reshape(B, -1, C),
                              self.spatial_shapes, self.start_index)
        return out.reshape(B, J, C)


class DeformMultiViewAttn(_DeformViewAttn):      # egoposeformer_heatmap_mvf_ex.py:767-796
    pass


class DeformStereoAttn(_DeformViewAttn):         # egoposeformer_mvf_ex.py:455-478
    pass


class FFN(nn.Module):                            # models/utils/transformer.py:8-33
    def __init__(self, embed_dims, feedforward_dims, num_fcs, ffn_drop):
        super().__init__()
        self.embed_dims, self.feedforward_channels, self.num_fcs = embed_dims, feedforward_dims, num_fcs
        layers, cin = [], embed_dims
        for _ in range(num_fcs - 1):
            layers.append(nn.Sequential(nn.Linear(cin, feedforward_dims), nn.GELU(), nn.Dropout(ffn_drop)))
            cin = feedforward_dims
        layers += [nn.Linear(feedforward_dims, embed_dims), nn.Dropout(ffn_drop)]
        self.layers = nn.Sequential(*layers)

    def forward(self, x):
        return self.layers(x)


class CustomMultiheadAttention(nn.Module):       # models/utils/transformer.py:36-107
    def __init__(self, embed_dim, num_heads, dropout=0.0, bias=True, add_bias_kv=False, add_zero_attn=False, kdim=None,
                 vdim=None, batch_first=False, with_output_proj=True, device=None, dtype=None):
        super().__init__()
        assert dropout == 0.0 and kdim is None and vdim is None and not add_bias_kv and not add_zero_attn and batch_first
        self.input_dims = self.embed_dims = embed_dim
        self.num_heads, self.with_bias = num_heads, bias
        self.head_dims = embed_dim // num_heads
        self.scale = self.head_dims ** -0.5
        self.q_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
        self.k_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
        self.v_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=bias) if with_output_proj else None
        for m in (self.q_proj, self.k_proj, self.v_proj):
            nn.init.xavier_uniform_(m.weight)
            if bias:
                nn.init.constant_(m.bias, 0.0)
        if bias and self.out_proj is not None:
            nn.init.constant_(self.out_proj.bias, 0.0)

    def _attend(self, q, k, v, bias=None):
        B = q.shape[0]
        sp = lambda t: t.reshape(B, -1, self.num_heads, self.head_dims).permute(0, 2, 1, 3)
        _q, _k, _v = sp(self.q_proj(q)), sp(self.k_proj(k)), sp(self.v_proj(v))
        attn = (_q @ _k.transpose(-2, -1)) * self.scale
        if bias is not None:
            attn = attn + bias
        attn = attn.softmax(dim=-1)
        x = (attn @ _v).permute(0, 2, 1, 3).reshape(B, -1, self.embed_dims)
        return (self.out_proj(x) if self.out_proj is not None else x), attn

    def forward(self, q, k, v):
        return self._attend(q, k, v)


class SpatialMHA(CustomMultiheadAttention):           # egoposeformer_heatmap_mvf_ex.py:799-817
    def forward(self, q, k, v, bias):
        return self._attend(q, k, v, bias)[0]


class EgoformerSpatialMHA(SpatialMHA):                # egoposeformer_mvf_ex.py:481-498
    pass


class _ViewTransformerLayer(nn.Module):
    _attn_cls = DeformMultiViewAttn
    _mha_cls = SpatialMHA

    def __init__(self, num_views, embed_dims, cross_attn_cfg, spatial_attn_cfg, ffn_cfg, feat_shape, **extra):
        super().__init__()
        if extra.pop("use_normal_cross_attn", False):
            raise NotImplementedError("egorear_b200: use_normal_cross_attn=True is not used by any shipped config")
        self.embed_dims = embed_dims
        ca = copy.deepcopy(cross_attn_cfg)
        ca.update({"embed_dim": embed_dims, "feat_shape": feat_shape})
        ca.pop("batch_first", None)
        self.cross_attn = self._attn_cls(**ca)
        self.fuse_mlp = nn.Linear(num_views * embed_dims, embed_dims)
        self.norm_cross = nn.LayerNorm(embed_dims)
        sa = copy.deepcopy(spatial_attn_cfg)
        sa.update({"embed_dim": embed_dims})
        self.spatial_attn = self._mha_cls(**sa)
        self.norm_spatial = nn.LayerNorm(embed_dims)
        fc = copy.deepcopy(ffn_cfg)
        fc.update({"embed_dims": embed_dims})
        self.ffn = FFN(**fc)
        self.norm_ffn = nn.LayerNorm(embed_dims)

    def forward(self, feat_query, frame_feat_multi_view, anchors_2d, anchors_valid):
        """As-executed layer on an already projected memory [B,V,HW,C] (module-level drop-in; the engines use the
        folded kernel instead).  :874-935 / :546-588."""
        V = frame_feat_multi_view.shape[1]
        x = feat_query
        per_view = []
        for i in range(V):
            r = self.cross_attn(x, frame_feat_multi_view[:, i], anchors_2d[:, i])
            per_view.append(r.masked_fill(~anchors_valid[:, i][..., None].expand_as(r), 0.0))
        x = self.norm_cross(x + self.fuse_mlp(torch.cat(per_view, dim=-1)))
        x = self.norm_spatial(x + self.spatial_attn(x, x, x, bias=None))
        return self.norm_ffn(x + self.ffn(x))


class MultiViewTransformerLayer(_ViewTransformerLayer):         # egoposeformer_heatmap_mvf_ex.py:820-935
    pass


class EgoPoseFormerTransformerLayer(_ViewTransformerLayer):     # egoposeformer_mvf_ex.py:501-588
    _attn_cls = DeformStereoAttn
    _mha_cls = EgoformerSpatialMHA


class TransformerHeadLayer(nn.Module):                           # egoposeformer_heatmap_mvf_ex.py:938-970
    def __init__(self, input_dims, output_dims):
        super().__init__()
        if output_dims != 128:
            raise NotImplementedError("egorear_b200: TransformerHeadLayer output_dims=%d (shipped configs use 128)" % output_dims)
        self.head = _seq_conv([("conv", input_dims, output_dims // 2, 1, 1, 0), "relu", "up",
                               ("conv", output_dims // 2, output_dims, 1, 1, 0), "relu"])

    def forward(self, x):
        return self.head(x)


# ------------------------------------------------------------------------------------------------------
# HeatmapMVF                                                   egoposeformer_heatmap_mvf_ex.py:442-731
# ------------------------------------------------------------------------------------------------------
class HeatmapMVF(_EngineOwner, nn.Module):
    def __init__(self, input_dims, embed_dims, num_former_layers, image_size, feat_down_stride, detach_heatmap_feat,
                 mvf_transformer_cfg, heatmap_threshold, num_views, num_heatmap, joint_query_adaptation=False,
                 joint_query_adaptation_multi_view=False, joint_query_only=False, use_1by1_conv=False,
                 precision="fp16"):
        super().__init__()
        if not joint_query_adaptation or joint_query_adaptation_multi_view or joint_query_only or use_1by1_conv:
            raise NotImplementedError("egorear_b200: only the shipped jqa configuration (joint_query_adaptation=True, "
                                      "use_1by1_conv=False) is implemented")
        if num_former_layers != 1 or input_dims != 128 or embed_dims != 256:
            raise NotImplementedError("egorear_b200: mvf_cfg must be input_dims=128, embed_dims=256, num_former_layers=1")
        self.num_heatmap, self.num_views = num_heatmap, num_views
        self.heatmap_threshold, self.detach_heatmap_feat = heatmap_threshold, detach_heatmap_feat
        self.joint_query_adaptation, self.use_1by1_conv = True, False
        self.joint_query_adaptation_multi_view = self.joint_query_only = False
        self.feat_shape = (image_size[0] // feat_down_stride, image_size[1] // feat_down_stride)
        hw = self.feat_shape[0] * self.feat_shape[1]
        self.heatmap_proj = nn.Sequential(nn.Linear(hw, embed_dims), nn.ReLU(inplace=False), nn.Linear(embed_dims, embed_dims))
        self.fc_bfb = nn.Linear(512, embed_dims)
        self.fc_query = nn.Sequential(nn.Linear(embed_dims, embed_dims), nn.ReLU(inplace=False))
        self.joint_query_embed = nn.Embedding(num_heatmap, embed_dims)
        self.frame_feat_multi_view_proj = nn.Conv2d(input_dims, embed_dims, 1, 1, 0)
        self.frame_feat_multi_view_pos_embed = nn.Parameter(torch.zeros(1, num_views, hw, embed_dims))
        c = input_dims
        self.frame_feat_proj_layers = _seq_conv([("conv", c, 2 * c, 1, 1, 0), "relu", ("conv", 2 * c, 4 * c, 3, 2, 1), "relu",
                                                 ("conv", 4 * c, c, 1, 1, 0), "relu"])
        self.transformer_layers = nn.ModuleList()
        for _ in range(num_former_layers):
            cfg = copy.deepcopy(mvf_transformer_cfg)
            cfg.update({"num_views": num_views, "embed_dims": embed_dims, "feat_shape": self.feat_shape})
            self.transformer_layers.append(MultiViewTransformerLayer(**cfg))
        self.post_norm = nn.ModuleList([nn.LayerNorm(embed_dims) for _ in range(num_former_layers)])
        self.head_layers = nn.ModuleList([TransformerHeadLayer(num_heatmap, input_dims) for _ in range(num_former_layers)])
        self.frame_feat_refined_proj_layers = nn.ModuleList([
            _seq_conv([("conv", c, c, 1, 1, 0), "relu", "up", ("conv", c, c, 1, 1, 0), "relu"]) for _ in range(num_former_layers)])
        self.conv_heatmap_layers = nn.ModuleList([
            _seq_conv([("conv", c, 2 * c, 3, 2, 1), "relu", ("conv", 2 * c, 2 * c, 1, 1, 0), "relu", "up",
                       ("conv", 2 * c, c, 1, 1, 0), "relu", ("conv", c, num_heatmap, 1, 1, 0)]) for _ in range(num_former_layers)])
        self._precision = precision
        self._engine = None   # standalone use (not owned by an EgoPoseFormerHeatmapMVFEX)

    def _engine_params(self):
        return {"heatmap_refiner_front_left." + k: v for k, v in self.state_dict(keep_vars=True).items() if v.is_floating_point()}

    def _make_engine(self):
        return MvfexEngine(self.num_views, self.num_heatmap, self.heatmap_threshold, self._precision)

    def forward(self, heatmap, frame_feat, frame_feat_multi_view, anchors_2d, anchors_valid, backbone_feat_bottom,
                backbone_feat_bottom_multi_view):
        self._check_inference()
        hm, ft = torch.ops.egr.mvfex_refiner_forward(self._egr_key, heatmap, frame_feat, frame_feat_multi_view, anchors_2d,
                                                     anchors_valid, backbone_feat_bottom)
        return [hm], [ft]


# ------------------------------------------------------------------------------------------------------
# EgoPoseFormerHeatmapMVFEX                                     egoposeformer_heatmap_mvf_ex.py:27-437
# ------------------------------------------------------------------------------------------------------
class EgoPoseFormerHeatmapMVFEX(_EngineOwner, nn.Module):
    def __init__(self, num_views, image_size, num_heatmap, feat_down_stride, heatmap_threshold, encoder_cfg, mvf_cfg,
                 camera_model, full_training=False, detach_heatmap_feat=False, detach_heatmap_feat_init=False,
                 use_pred_heatmap_init=False, no_detach_feat_init=False, precision="fp16", build_backbone=True,
                 backbone_impl="torch", **kwargs):
        """backbone_impl (not in the reference): "torch" = the PyTorch ResNet18+FPN modules run as in the reference (the
        features exist as NCHW fp32 tensors and are returned in list_frame_feat[0]); "egr" = the backbone engine
        (tcgen05 conv stages, 16-bit precisions): the FPN map goes to the hot path as its staged channels-last copy and
        list_frame_feat[0] is None."""
        super().__init__()
        if backbone_impl not in ("torch", "egr"):
            raise ValueError("backbone_impl must be 'torch' or 'egr'")
        self.backbone_impl = backbone_impl
        self._bb_engine = None
        if num_views not in (2, 4):
            raise NotImplementedError("egorear_b200: num_views must be 4 or 2 (no shipped config uses 3)")
        self.num_views, self.num_heatmap = num_views, num_heatmap
        self.heatmap_threshold, self.camera_model = heatmap_threshold, camera_model
        self.full_training, self.detach_heatmap_feat = full_training, detach_heatmap_feat
        self.detach_heatmap_feat_init, self.use_pred_heatmap_init = detach_heatmap_feat_init, use_pred_heatmap_init
        self.no_detach_feat_init = no_detach_feat_init
        cfg = copy.deepcopy(mvf_cfg)
        cfg.update({"num_views": num_views, "num_heatmap": num_heatmap, "heatmap_threshold": heatmap_threshold,
                    "image_size": image_size, "feat_down_stride": feat_down_stride, "detach_heatmap_feat": detach_heatmap_feat})
        if build_backbone:
            self.heatmap_estimator_stereo_front = EgoPoseFormerHeatmap(encoder_cfg, num_heatmap, detach_heatmap_feat_init)
            if num_views == 4:
                self.heatmap_estimator_stereo_back = EgoPoseFormerHeatmap(encoder_cfg, num_heatmap, detach_heatmap_feat_init)
        self._has_backbone = build_backbone
        for name in _VIEWS4[:num_views]:
            setattr(self, "heatmap_refiner_" + name, HeatmapMVF(precision=precision, **cfg))
        self.use_1by1_conv = False
        c = 128
        head = lambda: _seq_conv([("conv", c, c, 1, 1, 0), "relu", ("conv", c, 2 * c, 3, 2, 1), "relu",
                                  ("conv", 2 * c, 2 * c, 1, 1, 0), "relu", "up", ("conv", 2 * c, c, 1, 1, 0), "relu",
                                  ("conv", c, num_heatmap, 1, 1, 0)])
        self.conv_heatmap_layers_stereo_front = head()
        if num_views == 4:
            self.conv_heatmap_layers_stereo_back = head()
        self._precision = precision
        self._engine = None

    # -- engine plumbing --
    def hot_path_state(self):
        """state_dict entries the C engine consumes (everything except the two backbone estimators)."""
        return {k: v for k, v in self.state_dict(keep_vars=True).items()
                if not k.startswith("heatmap_estimator_") and v.is_floating_point()}

    _engine_params = hot_path_state

    def _make_engine(self):
        return MvfexEngine(self.num_views, self.num_heatmap, self.heatmap_threshold, self._precision)

    # -- backbone engine --
    def backbone_state(self):
        """state_dict entries the backbone engine consumes: conv weights / biases and the BN affine + running statistics of
        both encoders (not num_batches_tracked, not the estimators' own unused 1x1 heads)"""
        return {k: v for k, v in self.state_dict(keep_vars=True).items()
                if k.startswith("heatmap_estimator_") and ".encoder." in k and v.is_floating_point()}

    def backbone_engine(self):
        if not self._has_backbone:
            raise RuntimeError("built with build_backbone=False: there are no backbone parameters")
        if self._bb_engine is None:
            self._bb_engine = BackboneEngine(self.num_views, self._precision)
            self._bb_engine.set_params(self.backbone_state())
            self.engine()                   # installs the load_state_dict hook, which invalidates both engines
        return self._bb_engine

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        if self._bb_engine is not None:
            st = self.backbone_state()
            if all(v.is_cuda for v in st.values()):
                self._bb_engine.set_params(st)
            else:
                self._bb_engine = None
        return out

    def forward_backbone_staged(self, img, lane=0):
        """img [B,V,3,256,256] -> (feat_staged [V,B,64,64,128] 16-bit channels-last view-major, bfb [B,V,512,8,8] fp32)
        through the backbone engine (eval-mode BatchNorm folded into the convs)"""
        if self.heatmap_estimator_stereo_front.training:
            raise RuntimeError("egorear_b200: the backbone engine folds eval-mode BatchNorm; call .eval() first")
        return torch.ops.egr.backbone_forward(self._egr_key, img, int(lane))

    def get_anchors_2d_from_hm(self, heatmap):           # :128-143
        with torch.no_grad():
            B, V, C, H, W = heatmap.shape
            pts2d, maxvals, valid = ops.get_max_preds(heatmap.reshape(B * V, C, H, W), threshold=self.heatmap_threshold,
                                                      normalize=True)
            return (pts2d.view(B, V, self.num_heatmap, 2), maxvals.view(B, V, self.num_heatmap),
                    valid.view(B, V, self.num_heatmap))

    def forward_heatmap_feat_estimation(self, img):      # :212-234 (PyTorch backbones)
        if self.num_views == 4:
            ff, bf = self.heatmap_estimator_stereo_front.forward_backbone(img[:, 0:2], return_feat=True)
            fb, bb = self.heatmap_estimator_stereo_back.forward_backbone(img[:, 2:], return_feat=True)
            return torch.cat((ff, fb), dim=1), [torch.cat((a, b), dim=1) for a, b in zip(bf, bb)]
        return self.heatmap_estimator_stereo_front.forward_backbone(img, return_feat=True)

    def forward_from_feats(self, frame_feat_multi_view, backbone_feat_bottom_multi_view, heatmap_for_anchor=None,
                           want_feat_refined=True, feat_staged=None, lane=0):
        """The hot path proper: backbone features in, (list_heatmap_pred, list_frame_feat) out (:284-437).
        want_feat_refined=False (chained forward with exported channels-last copies only): list_frame_feat[1] is None.
        feat_staged: view-major channels-last bf16 features [V,B,64,64,128] straight from a channels-last bf16 backbone
        (frame_feat_multi_view may then be None; see MvfexEngine.forward)."""
        self._check_inference()
        hfa = heatmap_for_anchor if isinstance(heatmap_for_anchor, torch.Tensor) else None
        hm_init, hm_refined, feat_refined, a2, av = torch.ops.egr.mvfex_forward(
            self._egr_key, None if feat_staged is not None else frame_feat_multi_view, backbone_feat_bottom_multi_view, hfa,
            feat_staged, bool(want_feat_refined), int(lane))
        if not torch.compiler.is_compiling():          # debugging / test conveniences, not part of the traced graph
            self.last_anchors = (a2, av)
        return [hm_init, hm_refined], [None if feat_staged is not None else frame_feat_multi_view,
                                       feat_refined if want_feat_refined else None]

    def forward(self, img, heatmap_for_anchor=None):
        if not self._has_backbone:
            raise RuntimeError("built with build_backbone=False: call forward_from_feats(feat, bfb)")
        if self.backbone_impl == "egr":
            xh, bfb = self.forward_backbone_staged(img)
            return self.forward_from_feats(None, bfb, heatmap_for_anchor, feat_staged=xh)
        if self.full_training:
            feat, bb = self.forward_heatmap_feat_estimation(img)
        else:
            with torch.no_grad():
                feat, bb = self.forward_heatmap_feat_estimation(img)
        return self.forward_from_feats(feat, bb[-1], heatmap_for_anchor)


# ------------------------------------------------------------------------------------------------------
# EgoPoseFormerPose3D / EgoPoseFormerMVFEX                         egoposeformer_mvf_ex.py:22-452
# ------------------------------------------------------------------------------------------------------
class EgoPoseFormerPose3D(_EngineOwner, nn.Module):
    def __init__(self, num_views, image_size, use_pred_heatmap_init, num_joints, input_dims, embed_dims, mlp_dims,
                 mlp_dropout, num_mlp_layers, transformer_cfg, num_former_layers, num_pred_mlp_layers, camera_model,
                 feat_down_stride, coor_norm_max, coor_norm_min, conv_heatmap_dim_init, norm_mlp_pred=False,
                 use_mlp_avgpool=True, use_mlp_heatmap=False, camera_calib_file_dir_path=None, precision="fp16", **kwargs):
        super().__init__()
        if use_mlp_avgpool or use_mlp_heatmap:
            raise NotImplementedError("egorear_b200: only the conv-MLP proposal branch (use_mlp_avgpool=False, "
                                      "use_mlp_heatmap=False) of the shipped configs is implemented")
        if input_dims != 128 or embed_dims != 128 or num_mlp_layers != 2 or num_pred_mlp_layers != 2:
            raise NotImplementedError("egorear_b200: pose3d_cfg must match the shipped configs (128/128, 2 MLP layers)")
        from .calib import cameras_for
        cameras_for(camera_model)                      # raises ValueError('Unknown camera model !') like the reference
        self.invalid_pad = 1e10
        self.num_views, self.num_joints, self.embed_dims = num_views, num_joints, embed_dims
        self.feat_down_stride, self.image_size, self.camera_model = feat_down_stride, image_size, camera_model
        self.feat_shape = (image_size[0] // feat_down_stride, image_size[1] // feat_down_stride)
        self.use_pred_heatmap_init = use_pred_heatmap_init
        self.use_mlp_avgpool, self.use_mlp_heatmap, self.norm_mlp_pred = False, False, norm_mlp_pred
        self._calib = load_calibration(camera_calib_file_dir_path)
        self.feat_proj = nn.Conv2d(input_dims, embed_dims, 1, 1, 0)
        self.layers = nn.ModuleList()
        for _ in range(num_former_layers):
            cfg = copy.deepcopy(transformer_cfg)
            cfg.update({"num_views": num_views, "embed_dims": embed_dims, "feat_shape": self.feat_shape})
            self.layers.append(EgoPoseFormerTransformerLayer(**cfg))
        self.query_gen_mlp = nn.Sequential(nn.Linear(4, embed_dims), nn.ReLU(inplace=True), nn.Linear(embed_dims, embed_dims),
                                           nn.ReLU(inplace=True), nn.Linear(embed_dims, embed_dims))
        c = input_dims
        self.conv_frame_feat = _seq_conv([("conv", c, c // 2, 1, 1, 0), "relu", ("conv", c // 2, c, 3, 2, 1), "relu", "pool",
                                          ("conv", c, c // 2, 1, 1, 0), "relu", ("conv", c // 2, c, 3, 2, 1), "relu"])
        mlp, cin = [], num_views * 128 * 8 * 8
        for _ in range(num_mlp_layers):
            mlp.append(nn.Sequential(nn.Linear(cin, cin // 16), nn.GELU(), nn.Dropout(mlp_dropout)))
            cin //= 16
        mlp.append(nn.Linear(cin, 3 * num_joints))
        self.mlp_pred = nn.Sequential(*mlp)
        self.reg_mlp = nn.ModuleList([nn.Sequential(nn.Linear(embed_dims, embed_dims), nn.GELU(), nn.Linear(embed_dims, 3))
                                      for _ in range(num_former_layers)])
        self.post_norm = nn.ModuleList([nn.LayerNorm(embed_dims) for _ in range(num_former_layers)])
        if norm_mlp_pred:
            self.register_buffer("coor_min", torch.tensor(coor_norm_min))
            self.register_buffer("coor_max", torch.tensor(coor_norm_max))
        self._precision = precision
        self._engine = None

    def hot_path_state(self):
        return {k: v for k, v in self.state_dict(keep_vars=True).items()
                if v.is_floating_point() and not k.startswith("coor_")}

    _engine_params = hot_path_state

    def _make_engine(self):
        return Pose3DEngine(self.num_views, self.num_joints, len(self.layers), self.camera_model,
                            self.use_pred_heatmap_init, self._precision, self._calib)

    def forward(self, frame_feats_init, frame_feats_final, heatmap, coord_trans_mat=None, origin_3d=None, staged=None, lane=0):
        # `heatmap` and `origin_3d` are accepted and unused, as in the reference's shipped configuration (:434-439)
        # `staged` (not in the reference): the EgoPoseFormerHeatmapMVFEX whose forward produced the inputs - its
        # channels-last copies are then lifted directly (they are matched against the tensors passed here by address)
        self._check_inference()
        chain = staged._egr_key if staged is not None else 0
        preds = torch.ops.egr.pose3d_forward(self._egr_key, chain, frame_feats_init, frame_feats_final, heatmap,
                                             coord_trans_mat, int(lane))
        return [preds[i] for i in range(preds.shape[0])]


class EgoPoseFormerMVFEX(nn.Module):
    def __init__(self, num_views, image_size, camera_model, heatmap_mvf_cfg, pose3d_cfg, precision="fp16", **kwargs):
        super().__init__()
        h = dict(heatmap_mvf_cfg)
        h.update({"num_views": num_views, "image_size": image_size, "camera_model": camera_model, "precision": precision})
        self.heatmap_estimator = EgoPoseFormerHeatmapMVFEX(**h)
        self.use_pred_heatmap_init = self.heatmap_estimator.use_pred_heatmap_init
        p = dict(pose3d_cfg)
        p.update({"num_views": num_views, "image_size": image_size, "use_pred_heatmap_init": self.use_pred_heatmap_init,
                  "camera_model": camera_model, "precision": precision})
        self.pose3d_estimator = EgoPoseFormerPose3D(**p)
        self._chain = True      # hand the heatmap engine's channels-last copies to the pose3d engine
        self.heatmap_estimator.chain_use_init = bool(self.use_pred_heatmap_init)

    def forward_from_feats(self, feat, bfb, coord_trans_mat=None, origin_3d=None):
        if self._chain and not torch.compiler.is_compiling():
            pd = self.pose3d_estimator.engine().proposal_dtype()
            hp = ("f16_only" if self.use_pred_heatmap_init else "f16") if pd == "f16" else "tf32" if pd == "tf32" else None
            self.heatmap_estimator.engine().export_staged(True, hp=hp)
        # the chained model returns (poses, heatmaps) only (:50-58): the refined features stay channels-last, internal
        list_hm, list_ff = self.heatmap_estimator.forward_from_feats(feat, bfb, want_feat_refined=not self._chain)
        return self.pose3d_estimator(list_ff[0], list_ff[-1], list_hm[-1], coord_trans_mat, origin_3d,
                                     staged=self.heatmap_estimator), list_hm

    def forward(self, img, coord_trans_mat=None, origin_3d=None):
        list_hm, list_ff = self.heatmap_estimator(img)
        return self.pose3d_estimator(list_ff[0], list_ff[-1], list_hm[-1], coord_trans_mat, origin_3d), list_hm
