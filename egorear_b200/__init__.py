"""egorear_b200 — B200-native (sm_100a) hot path of EgoRear behind the reference's module interfaces.

    import egorear_b200
    egorear_b200.patch(precision="fp16")      # swap the hot path under an imported EgoRear checkout

Everything computes through egorear_b200/libegorear_b200.so (C-ABI, include/egorear_b200.h); there is no CPU fallback.
"""
import functools
import importlib

__all__ = ["patch", "PATCH_TABLE", "METHOD_PATCH_TABLE"]

# reference module -> names rebound by patch() (reference file:line in INTEGRATION.md)
PATCH_TABLE = {
    "pose_estimation.utils.loss": ["get_max_preds", "get_max_preds_soft_pytorch"],
    "pose_estimation.utils.util": ["integrate_tensor_2d"],
    "generate_heatmap": ["generate_target"],
    "pose_estimation.models.utils.deform_attn": ["MSDeformAttn"],
    "pose_estimation.models.estimator.egoposeformer_heatmap": ["EgoPoseFormerHeatmap"],
    "pose_estimation.models.estimator.egoposeformer_heatmap_mvf_ex": [
        "EgoPoseFormerHeatmapMVFEX", "HeatmapMVF", "MultiViewTransformerLayer", "DeformMultiViewAttn", "SpatialMHA",
        "TransformerHeadLayer", "EgoPoseFormerHeatmap", "get_max_preds"],
    "pose_estimation.models.estimator.egoposeformer_mvf_ex": [
        "EgoPoseFormerMVFEX", "EgoPoseFormerPose3D", "EgoPoseFormerTransformerLayer", "DeformStereoAttn",
        "EgoformerSpatialMHA", "EgoPoseFormerHeatmapMVFEX"],
    "pose_estimation.models.estimator": ["EgoPoseFormerHeatmap", "EgoPoseFormerHeatmapMVFEX", "EgoPoseFormerMVFEX"],
    "pose_estimation.pl_wrappers.egoposeformer.heatmap": ["EgoPoseFormerHeatmap", "get_max_preds"],
    "pose_estimation.pl_wrappers.egoposeformer.heatmap_mvf_ex": ["EgoPoseFormerHeatmapMVFEX", "get_max_preds"],
    "pose_estimation.pl_wrappers.egoposeformer.pose_3d_mvf_ex": ["EgoPoseFormerMVFEX"],
}

# reference module -> {class: {method: egorear_b200.metrics function}}: the wrappers' eval-time metric methods
# (per-sample Python loops with .cpu() syncs / numpy SVDs in the reference)
METHOD_PATCH_TABLE = {
    "pose_estimation.pl_wrappers.egoposeformer.heatmap": {"PoseHeatmapLightningModel": {"evaluate": "evaluate"}},
    "pose_estimation.pl_wrappers.egoposeformer.heatmap_mvf_ex": {"PoseHeatmapMVFEXLightningModel": {"evaluate": "evaluate"}},
    "pose_estimation.pl_wrappers.egoposeformer.pose_3d_mvf_ex": {
        "Pose3DMVFEXLightningModel": {"evaluate_pose": "evaluate_pose", "evaluate_heatmap": "evaluate_heatmap"}},
}

_PRECISION_CLASSES = ("EgoPoseFormerHeatmapMVFEX", "HeatmapMVF", "EgoPoseFormerPose3D", "EgoPoseFormerMVFEX")


def _with_precision(cls, precision):
    """subclass whose constructor defaults `precision` (the one kwarg the reference's YAML does not carry)"""
    @functools.wraps(cls, updated=())
    class _P(cls):
        def __init__(self, *a, **k):
            k.setdefault("precision", precision)
            super().__init__(*a, **k)
    _P.__name__, _P.__qualname__ = cls.__name__, cls.__qualname__
    return _P


def _as_method(fn):
    @functools.wraps(fn)
    def method(self, *a, **k):
        return fn(*a, **k)
    return method


def patch(precision="fp16", modules=None, strict=False, datasets=False):
    """Rebind the hot-path names inside the reference's modules to the B200 implementations.

    precision: "fp16" (tensor cores, inside the fp32 parity bound), "bf16" (tensor cores, stated looser bound) or "fp32"
    (SIMT, reference-grade).  datasets=True additionally moves the datasets' resize + normalise and heatmap-target loading
    to the GPU (egorear_b200/datasets.py): the dataset modules' call sites are substituted and the three wrapper classes
    get an `on_after_batch_transfer` hook that finishes the batch on the device.

    Only modules that can be imported (or are already in sys.modules) are touched; with strict=True a module of
    PATCH_TABLE that fails to import raises.  Returns {module name: [names rebound]}.
    Fails loudly (ImportError) when libegorear_b200.so has not been built.
    """
    if precision not in ("bf16", "fp16", "fp32"):
        raise ValueError("precision must be 'bf16', 'fp16' or 'fp32'")
    from . import _lib, modules as M, ops
    _lib.load()
    repl = {"get_max_preds": ops.get_max_preds, "generate_target": ops.generate_target,
            "get_max_preds_soft_pytorch": ops.get_max_preds_soft_pytorch, "integrate_tensor_2d": ops.integrate_tensor_2d}
    for name in set(n for names in PATCH_TABLE.values() for n in names):
        if name in repl:
            continue
        cls = getattr(M, name)
        repl[name] = _with_precision(cls, precision) if name in _PRECISION_CLASSES else cls
    done = {}
    for modname, names in PATCH_TABLE.items():
        if modules is not None and modname not in modules:
            continue
        try:
            mod = importlib.import_module(modname)
        except Exception:
            if strict:
                raise
            continue
        for n in names:
            if hasattr(mod, n) or strict:
                setattr(mod, n, repl[n])
                done.setdefault(modname, []).append(n)
    from . import metrics
    for modname, classes in METHOD_PATCH_TABLE.items():
        if modules is not None and modname not in modules:
            continue
        try:
            mod = importlib.import_module(modname)
        except Exception:
            if strict:
                raise
            continue
        for cname, methods in classes.items():
            cls = getattr(mod, cname, None)
            if cls is None:
                continue
            for mname, fname in methods.items():
                setattr(cls, mname, _as_method(getattr(metrics, fname)))
                done.setdefault(modname, []).append("%s.%s" % (cname, mname))
            if datasets:
                from . import datasets as egd
                cls.on_after_batch_transfer = egd.on_after_batch_transfer
                done.setdefault(modname, []).append("%s.on_after_batch_transfer" % cname)
    if datasets:
        from . import datasets as egd
        for name in egd.patch_datasets(strict=strict):
            done.setdefault(name, []).extend(["Image", "transforms", "np"])
    return done
