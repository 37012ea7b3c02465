"""The C-ABI entry points as `torch.library` custom operators (namespace `egr`), with fake (meta) implementations.

Why: the reference wraps `model.network` in `torch.compile(...)` whenever `model.compile` is set — and every shipped
config sets it (run.py:7-9, configs/*.yaml:60).  A module whose forward calls ctypes + `data_ptr()` cannot be traced;
a custom op can: Dynamo records one opaque node per op, AOT-autograd / Inductor see its output shapes through the
fake implementation, and the real implementation (the ctypes call into libegorear_b200.so) runs untouched at run time.
The mirror modules (`modules.py`, `ops.py`) call these ops, so `torch.compile(network)` captures the hot path as a
handful of `egr::*` nodes with no graph break inside it (tests/test_gpu_compile.py, tests/test_torch_ops.py).

Engine-backed ops (`egr::mvfex_forward`, `egr::pose3d_forward`, `egr::mvfex_refiner_forward`) take the owning module's
integer key instead of a handle: the module registers itself in a weak registry at construction, the op body looks it
up at run time and calls its engine (creating / re-packing it lazily, exactly like the eager path).  The channels-last
hand-over between a chained heatmap estimator and pose3d stays inside the two op bodies (keyed by the producer's key);
the data dependency between the two nodes is the heatmap tensor the reference passes along (egoposeformer_mvf_ex.py:57).

All ops are inference ops (no autograd formula): outputs never require grad.
"""
import weakref
from typing import List, Optional, Tuple

import torch
from torch import Tensor

_MODULES = weakref.WeakValueDictionary()
_NEXT_KEY = [1]
_STAGED = {}          # producer module key -> dict left by its last forward (device pointers into its workspace)


def register_module(module):
    """-> integer key under which the engine-backed ops find `module` (weak reference)"""
    key = _NEXT_KEY[0]
    _NEXT_KEY[0] += 1
    _MODULES[key] = module
    return key


def _module(key):
    m = _MODULES.get(int(key))
    if m is None:
        raise RuntimeError("egr op: module key %d is not alive (the module was garbage-collected)" % int(key))
    return m


def _none_if_empty(t):
    return None if t is None or t.numel() == 0 else t


# ------------------------------------------------------------------------------------------------------
# D1 / D1s / D1i
# ------------------------------------------------------------------------------------------------------
@torch.library.custom_op("egr::decode_argmax", mutates_args=())
def decode_argmax(heatmaps: Tensor, threshold: float, normalize: bool) -> Tuple[Tensor, Tensor, Tensor]:
    from . import ops
    p, m, v = ops._decode_argmax_impl(heatmaps, threshold, normalize)
    return p, m, v


@decode_argmax.register_fake
def _(heatmaps, threshold, normalize):
    B, J = heatmaps.shape[:2]
    return (heatmaps.new_empty((B, J, 2), dtype=torch.float32), heatmaps.new_empty((B, J, 1), dtype=torch.float32),
            heatmaps.new_empty((B, J, 1), dtype=torch.bool))


@torch.library.custom_op("egr::decode_soft_argmax", mutates_args=())
def decode_soft_argmax(heatmaps: Tensor, normalize: bool) -> Tuple[Tensor, Tensor]:
    from . import ops
    return ops._decode_soft_argmax_impl(heatmaps, normalize)


@decode_soft_argmax.register_fake
def _(heatmaps, normalize):
    B, J = heatmaps.shape[:2]
    return heatmaps.new_empty((B, J, 2), dtype=torch.float32), heatmaps.new_empty((B, J, 1), dtype=torch.float32)


@torch.library.custom_op("egr::integrate_tensor_2d", mutates_args=())
def integrate_tensor_2d(heatmaps: Tensor, softmax: bool, multiplier: float) -> Tuple[Tensor, Tensor]:
    from . import ops
    return ops._integrate_tensor_2d_impl(heatmaps, softmax, multiplier)


@integrate_tensor_2d.register_fake
def _(heatmaps, softmax, multiplier):
    B, J = heatmaps.shape[:2]
    return heatmaps.new_empty((B, J, 2), dtype=torch.float32), heatmaps.new_empty(tuple(heatmaps.shape), dtype=torch.float32)


# ------------------------------------------------------------------------------------------------------
# G1, MSDA, H1', P3, pack
# ------------------------------------------------------------------------------------------------------
@torch.library.custom_op("egr::generate_target", mutates_args=())
def generate_target(joints: Tensor, image_size: float, heatmap_size: int, sigma: float) -> Tensor:
    from . import ops
    return ops._generate_target_impl(joints, image_size, heatmap_size, sigma, None)


@generate_target.register_fake
def _(joints, image_size, heatmap_size, sigma):
    return joints.new_empty(tuple(joints.shape[:-1]) + (heatmap_size, heatmap_size), dtype=torch.float32)


@torch.library.custom_op("egr::msda_forward", mutates_args=())
def msda_forward(value: Tensor, H: int, W: int, sampling_locations: Tensor, attention_weights: Tensor) -> Tensor:
    from . import ops
    return ops._msda_impl(value, H, W, sampling_locations, attention_weights)


@msda_forward.register_fake
def _(value, H, W, sampling_locations, attention_weights):
    B, L, nh, hd = value.shape
    return value.new_empty((B, sampling_locations.shape[1], nh * hd), dtype=torch.float32)


@torch.library.custom_op("egr::heatmap_head_1x1", mutates_args=())
def heatmap_head_1x1(feat: Tensor, weight: Tensor, bias: Tensor) -> Tensor:
    from . import ops
    return ops._heatmap_head_1x1_impl(feat, weight, bias)


@heatmap_head_1x1.register_fake
def _(feat, weight, bias):
    N, C, H, W = feat.shape
    return feat.new_empty((N, weight.shape[0], H, W), dtype=torch.float32)


@torch.library.custom_op("egr::pack_joints", mutates_args=())
def pack_joints(preds2d: Tensor, pose3d: Tensor) -> Tensor:
    from . import ops
    return ops._pack_joints_impl(preds2d, pose3d)


@pack_joints.register_fake
def _(preds2d, pose3d):
    B = preds2d.shape[0]
    return preds2d.new_empty((B, preds2d.numel() // max(B, 1) + pose3d.numel() // max(B, 1)), dtype=torch.float32)


# ------------------------------------------------------------------------------------------------------
# engines
# ------------------------------------------------------------------------------------------------------
@torch.library.custom_op("egr::mvfex_forward", mutates_args=())
def mvfex_forward(module_key: int, feat: Optional[Tensor], bfb: Tensor, heatmap_for_anchor: Optional[Tensor],
                  feat_staged: Optional[Tensor], want_feat_refined: bool, lane: int) -> List[Tensor]:
    """-> [hm_init, hm_refined, feat_refined (numel 0 when not wanted), anchors_2d, anchors_valid]"""
    m = _module(module_key)
    eng = m.engine()
    if not want_feat_refined and not getattr(eng, "_export", False):
        # chained model: keep the channels-last copies for the consumer (the eager pipeline sets this up itself, from the
        # pose3d engine's actual proposal dtype; here the default fp16 proposal branch of the 16-bit precisions is assumed)
        use_init = bool(getattr(m, "chain_use_init", True))
        eng.export_staged(True, hp=(("f16_only" if use_init else "f16") if eng.precision in ("bf16", "fp16") else None))
    out = eng.forward(feat, bfb, heatmap_for_anchor, want_feat_refined=want_feat_refined, feat_staged=feat_staged, lane=lane)
    _STAGED[int(module_key)] = out.get("staged")
    fr = out["feat_refined"]
    if fr is None:
        fr = bfb.new_empty((0,), dtype=torch.float32)
    return [out["hm_init"], out["hm_refined"], fr, out["anchors_2d"], out["anchors_valid"]]


@mvfex_forward.register_fake
def _(module_key, feat, bfb, heatmap_for_anchor, feat_staged, want_feat_refined, lane):
    B, V = bfb.shape[:2]
    J = 15
    f32 = dict(dtype=torch.float32)
    return [bfb.new_empty((B, V, J, 64, 64), **f32), bfb.new_empty((B, V, J, 64, 64), **f32),
            bfb.new_empty((B, V, 128, 64, 64), **f32) if want_feat_refined else bfb.new_empty((0,), **f32),
            bfb.new_empty((B, V, J, 2), **f32), bfb.new_empty((B, V, J), dtype=torch.bool)]


@torch.library.custom_op("egr::mvfex_refiner_forward", mutates_args=())
def mvfex_refiner_forward(module_key: int, heatmap: Tensor, frame_feat: Tensor, feat_mv: Tensor, anchors_2d: Tensor,
                          anchors_valid: Tensor, bfb: Tensor) -> Tuple[Tensor, Tensor]:
    m = _module(module_key)
    return m.engine().refiner_forward(0, heatmap, frame_feat, feat_mv, anchors_2d, anchors_valid, bfb)


@mvfex_refiner_forward.register_fake
def _(module_key, heatmap, frame_feat, feat_mv, anchors_2d, anchors_valid, bfb):
    return (heatmap.new_empty(tuple(heatmap.shape), dtype=torch.float32),
            frame_feat.new_empty(tuple(frame_feat.shape), dtype=torch.float32))


@torch.library.custom_op("egr::backbone_forward", mutates_args=())
def backbone_forward(module_key: int, img: Tensor, lane: int) -> Tuple[Tensor, Tensor]:
    """ResnetBackbone.forward of both stereo estimators -> (feat_staged [V,B,64,64,128] 16-bit, bfb [B,V,512,8,8])"""
    return _module(module_key).backbone_engine().forward(img, lane=lane)


@backbone_forward.register_fake
def _(module_key, img, lane):
    B, V = img.shape[:2]
    m = _MODULES.get(int(module_key))
    dt = torch.float16 if (m is not None and getattr(m, "_precision", "fp16") != "bf16") else torch.bfloat16
    return img.new_empty((V, B, 64, 64, 128), dtype=dt), img.new_empty((B, V, 512, 8, 8), dtype=torch.float32)


@torch.library.custom_op("egr::pose3d_forward", mutates_args=())
def pose3d_forward(module_key: int, chain_key: int, feats_init: Optional[Tensor], feats_final: Optional[Tensor],
                   heatmap: Optional[Tensor], coord_trans_mat: Optional[Tensor], lane: int) -> Tensor:
    """-> preds [L+1, B, 16, 3].  chain_key: key of the heatmap estimator whose forward produced the inputs (its
    channels-last copies are then lifted directly), or 0.  `heatmap` is unused by the computation (as in the reference's
    shipped configuration, egoposeformer_mvf_ex.py:434-439) but orders this node after the producer's."""
    m = _module(module_key)
    staged = _STAGED.get(int(chain_key)) if chain_key else None
    return m.engine().forward(_none_if_empty(feats_init), _none_if_empty(feats_final), coord_trans_mat, staged=staged,
                              use_init=m.use_pred_heatmap_init, lane=lane)


@pose3d_forward.register_fake
def _(module_key, chain_key, feats_init, feats_final, heatmap, coord_trans_mat, lane):
    ref = next(t for t in (heatmap, feats_final, feats_init) if t is not None)
    B = ref.shape[0]
    m = _MODULES.get(int(module_key))
    L = len(m.layers) if m is not None else 3
    J = m.num_joints if m is not None else 16
    return ref.new_empty((L + 1, B, J, 3), dtype=torch.float32)
