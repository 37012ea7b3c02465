"""ORACLE helper (test infrastructure): import the LIVE reference from /root/reference on a CUDA-less host.

In the build container this is the checkout; the GPU box has no /root/reference, but it receives the git-ignored file
copy oracle/_ref/ (oracle/make_ref.py), which bench.py's CPU arm times there (`kind: "reference"`).  Used by
tests/golden/make_golden.py to generate the committed fixtures and by tests/test_oracle_*.py (skipped
when the checkout is absent) to pin oracle/model_ref.py and oracle/gt_decode.c against the reference.

Shims, none of which touches the reference's arithmetic on the hot path:
  * mmcv (un-vendored, no CPU kernel): MultiScaleDeformableAttnFunction -> the standard per-level
    F.grid_sample(align_corners=False, padding_mode="zeros") restatement of ms_deform_attn_forward;
  * timm.models.layers.to_2tuple (only used by dead PatchEmbed), natsort.natsorted (generate_heatmap import);
  * utils/camera_models.py hard-codes device="cuda": torch.tensor is wrapped to drop that kwarg on CPU hosts.
"""
import os
import sys
import types

import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
# the checkout itself in the build container; on the GPU box the git-ignored file copy oracle/_ref/ that
# oracle/make_ref.py (run by __graft_entry__.build()) leaves behind and gpurun ships with the snapshot
_CANDIDATES = [os.environ.get("EGOREAR_REFERENCE"), "/root/reference", os.path.join(_HERE, "_ref")]
REF = next((c for c in _CANDIDATES if c and os.path.isdir(os.path.join(c, "pose_estimation"))), "/root/reference")
FORCE_CPU = False      # bench.py's CPU arm on the GPU box: keep the camera model's tensors on the host


def available():
    return os.path.isdir(os.path.join(REF, "pose_estimation"))


def is_copy():
    """True when the reference in use is the shipped file copy (oracle/_ref) rather than the checkout"""
    return os.path.abspath(REF) == os.path.abspath(os.path.join(_HERE, "_ref"))


def _stub(name, **attrs):
    import importlib.machinery
    m = types.ModuleType(name)
    # a real spec: other packages probe optional dependencies with importlib.util.find_spec(name), which raises on a module
    # whose __spec__ is None (transformers does that for timm)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _MSDAGridSample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, value, spatial_shapes, level_start_index, sampling_locations, attention_weights, im2col_step):
        B, L, nh, hd = value.shape
        _, Q, _, nl, P, _ = sampling_locations.shape
        assert nl == 1
        H, W = int(spatial_shapes[0, 0]), int(spatial_shapes[0, 1])
        v = value.permute(0, 2, 3, 1).reshape(B * nh, hd, H, W)
        grid = (2 * sampling_locations[:, :, :, 0] - 1).permute(0, 2, 1, 3, 4).reshape(B * nh, Q, P, 2)
        s = F.grid_sample(v, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
        w = attention_weights[:, :, :, 0].permute(0, 2, 1, 3).reshape(B * nh, 1, Q, P)
        return (s * w).sum(-1).view(B, nh * hd, Q).transpose(1, 2).contiguous()


_done = False


def install():
    global _done
    if _done:
        return
    if not available():
        raise RuntimeError("reference checkout not found at %s" % REF)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    if "mmcv" not in sys.modules:
        _stub("mmcv")
        _stub("mmcv.ops")
        _stub("mmcv.ops.multi_scale_deform_attn", MultiScaleDeformableAttnFunction=_MSDAGridSample)
    if "timm" not in sys.modules:
        _stub("timm")
        _stub("timm.models")
        _stub("timm.models.layers", to_2tuple=lambda x: tuple(x) if isinstance(x, (tuple, list)) else (x, x))
    if "natsort" not in sys.modules:
        _stub("natsort", natsorted=sorted)
    _done = True


def import_estimators():
    install()
    from pose_estimation.models.estimator import EgoPoseFormerHeatmap, EgoPoseFormerHeatmapMVFEX, EgoPoseFormerMVFEX
    from pose_estimation.models.estimator.egoposeformer_mvf_ex import EgoPoseFormerPose3D
    from pose_estimation.models.estimator.egoposeformer_heatmap_mvf_ex import HeatmapMVF
    if FORCE_CPU or not torch.cuda.is_available():
        import pose_estimation.utils.camera_models as cm
        if not getattr(cm, "_egr_cpu_shim", False):
            real = torch.tensor

            class _TorchProxy:
                def __getattr__(self, k):
                    return getattr(torch, k)

                @staticmethod
                def tensor(*a, **kw):
                    kw.pop("device", None)
                    return real(*a, **kw)
            cm.torch = _TorchProxy()
            cm._egr_cpu_shim = True
    return dict(EgoPoseFormerHeatmap=EgoPoseFormerHeatmap, EgoPoseFormerHeatmapMVFEX=EgoPoseFormerHeatmapMVFEX,
                EgoPoseFormerMVFEX=EgoPoseFormerMVFEX, EgoPoseFormerPose3D=EgoPoseFormerPose3D, HeatmapMVF=HeatmapMVF)


def import_functions():
    install()
    from pose_estimation.utils import loss as ref_loss
    import generate_heatmap

    def soft(batch_heatmaps, normalize=False):
        # utils/loss.py:160-161 builds its index vectors with torch.cuda.comm.broadcast / torch.cuda.FloatTensor;
        # on a CUDA-less host both are shimmed to their CPU equivalents for the duration of the call
        import torch
        if torch.cuda.is_available():
            return ref_loss.get_max_preds_soft_pytorch(batch_heatmaps, normalize)
        import torch.cuda.comm as comm          # not imported by `import torch` itself
        old = (comm.broadcast, torch.cuda.FloatTensor)
        comm.broadcast = lambda t, devices=None: [t]
        torch.cuda.FloatTensor = torch.FloatTensor
        try:
            return ref_loss.get_max_preds_soft_pytorch(batch_heatmaps, normalize)
        finally:
            comm.broadcast, torch.cuda.FloatTensor = old
    return dict(get_max_preds=ref_loss.get_max_preds, generate_target=generate_heatmap.generate_target,
                get_max_preds_soft_pytorch=soft)


def import_wrappers():
    """The Lightning wrappers whose eval-time metric methods (evaluate / evaluate_pose / evaluate_heatmap) the golden
    generator calls unbound.  pytorch_lightning is not installed: LightningModule -> nn.Module, ParallelStrategy -> object
    (neither is touched by those methods)."""
    install()
    import torch.nn as nn
    if "pytorch_lightning" not in sys.modules:
        _stub("pytorch_lightning", LightningModule=nn.Module)
        _stub("pytorch_lightning.strategies", ParallelStrategy=object)
    from pose_estimation.pl_wrappers.egoposeformer.heatmap_mvf_ex import PoseHeatmapMVFEXLightningModel
    from pose_estimation.pl_wrappers.egoposeformer.pose_3d_mvf_ex import Pose3DMVFEXLightningModel
    from pose_estimation.models.utils import pose_metric
    from pose_estimation.utils import loss as ref_loss
    return dict(PoseHeatmapMVFEXLightningModel=PoseHeatmapMVFEXLightningModel,
                Pose3DMVFEXLightningModel=Pose3DMVFEXLightningModel, pose_metric=pose_metric, loss=ref_loss)


def load_model_cfg(name):
    import yaml
    cfg = yaml.safe_load(open(os.path.join(REF, "configs", name)))["model"]["init_args"]["model_cfg"]

    def _no_pretrain(d):
        if isinstance(d, dict):
            for k, v in d.items():
                if k == "use_imagenet_pretrain":
                    d[k] = False
                else:
                    _no_pretrain(v)
    _no_pretrain(cfg)
    if "pose3d_cfg" in cfg:
        cfg["pose3d_cfg"]["camera_calib_file_dir_path"] = os.path.join(REF, "pose_estimation/utils/camera_calib_file/ego4view")
    return cfg
