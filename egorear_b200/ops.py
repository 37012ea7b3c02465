"""Python entry points over the C-ABI: the functions of the reference that the extension replaces.

Same names, argument meaning and error behaviour as the reference functions:

  get_max_preds     <- pose_estimation/utils/loss.py:122-142
  integrate_tensor_2d <- pose_estimation/utils/util.py:80-109
  generate_target   <- generate_heatmap.py:10-48
  ms_deform_attn    <- mmcv MultiScaleDeformableAttnFunction (models/utils/deform_attn.py:155-162)
  reproject_fisheye <- EgoPoseFormerPose3D._reproject_3d_to_2d (estimator/egoposeformer_mvf_ex.py:340-382)

Everything runs on the tensors' CUDA device through libegorear_b200.so; there is no CPU path.  The public functions go
through the `egr::*` torch.library operators of torch_ops.py (so `torch.compile` of a model that calls them records one
opaque node each, run.py:7-9); the `_*_impl` functions below are the operator bodies: the ctypes calls.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .calib import CAMERA_NAMES, cameras_for
from . import torch_ops  # noqa: F401  (registers the egr::* operators)


def _stream(t=None):
    """the current stream of t's device (of the current device without t)"""
    return ctypes.c_void_p(torch.cuda.current_stream(t.device if t is not None else None).cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _need_cuda(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError("egorear_b200.%s: expected a CUDA tensor (there is no CPU fallback)" % name)


def _on(t):
    """context: make t's device current for the launch (a tensor on cuda:1 while cuda:0 is current must not launch there)"""
    return torch.cuda.device_of(t)


def _aligned16(t):
    """the 128-bit kernels need 16-byte aligned bases: a contiguous view at an odd storage offset is copied"""
    return t if t.data_ptr() % 16 == 0 else t.clone()


def _same_device(name, ref, *others):
    for o in others:
        if o is not None and isinstance(o, torch.Tensor) and o.device != ref.device:
            raise RuntimeError("egorear_b200.%s: tensors live on different devices (%s vs %s)" % (name, ref.device, o.device))


# ------------------------------------------------------------------------------------------------
# D1
# ------------------------------------------------------------------------------------------------
def get_max_preds(heatmaps, threshold=0.5, normalize=False, return_index=False):
    """Drop-in for pose_estimation.utils.loss.get_max_preds (same asserts, same squeeze quirk)."""
    assert isinstance(heatmaps, torch.Tensor), 'heatmaps should be a torch.Tensor'
    assert heatmaps.ndim == 4, 'heatmaps should be 4-ndim'
    _need_cuda(heatmaps, "get_max_preds")
    if return_index:
        preds, maxvals, valid, idx = _decode_argmax_impl(heatmaps, threshold, normalize, True)
        return preds, maxvals.squeeze(), valid.squeeze(), idx
    preds, maxvals, valid = torch.ops.egr.decode_argmax(heatmaps, float(threshold), bool(normalize))
    return preds, maxvals.squeeze(), valid.squeeze()      # loss.py:142 squeezes size-1 dims


def _decode_argmax_impl(heatmaps, threshold, normalize, return_index=False):
    B, J, H, W = heatmaps.shape
    hm = heatmaps.detach()
    if hm.dtype != torch.float32 or not hm.is_contiguous():
        hm = hm.float().contiguous()
    hm = _aligned16(hm)
    preds = torch.empty((B, J, 2), dtype=torch.float32, device=hm.device)
    maxvals = torch.empty((B, J, 1), dtype=torch.float32, device=hm.device)
    valid = torch.empty((B, J, 1), dtype=torch.bool, device=hm.device)
    idx = torch.empty((B, J), dtype=torch.int32, device=hm.device) if return_index else None
    lib = _lib.load()
    with _on(hm):
        _lib.check(lib.egr_decode_argmax(_ptr(hm), B, J, H, W, float(threshold), int(bool(normalize)), _ptr(preds),
                                         _ptr(maxvals), _ptr(valid), _ptr(idx), _stream(hm)))
    return (preds, maxvals, valid, idx) if return_index else (preds, maxvals, valid)


def get_max_preds_soft_pytorch(batch_heatmaps, normalize=False):
    """Drop-in for pose_estimation.utils.loss.get_max_preds_soft_pytorch (:145-177): soft-argmax joint decoding.
    -> preds [B, J, 2] (expected x, y under softmax over H*W), maxvals [B, J, 1]."""
    assert len(batch_heatmaps.shape) == 4, 'batch_images should be 4-ndim (B, J, H, W)'
    _need_cuda(batch_heatmaps, "get_max_preds_soft_pytorch")
    return torch.ops.egr.decode_soft_argmax(batch_heatmaps, bool(normalize))


def _decode_soft_argmax_impl(batch_heatmaps, normalize):
    B, J, H, W = batch_heatmaps.shape
    hm = batch_heatmaps.detach()
    if hm.dtype != torch.float32 or not hm.is_contiguous():
        hm = hm.float().contiguous()
    hm = _aligned16(hm)
    preds = torch.empty((B, J, 2), dtype=torch.float32, device=hm.device)
    maxvals = torch.empty((B, J, 1), dtype=torch.float32, device=hm.device)
    with _on(hm):
        _lib.check(_lib.load().egr_decode_soft_argmax(_ptr(hm), B, J, H, W, int(bool(normalize)), _ptr(preds), _ptr(maxvals),
                                                      _stream(hm)))
    return preds, maxvals


def integrate_tensor_2d(heatmaps, softmax=True, multiplier=100.0):
    """Drop-in for pose_estimation.utils.util.integrate_tensor_2d (:80-109): integral (soft-argmax) decoding.
    -> coordinates [B, J, 2] (x, y) and the normalised heatmaps [B, J, H, W] (softmax(hm * multiplier) or relu)."""
    _need_cuda(heatmaps, "integrate_tensor_2d")
    return torch.ops.egr.integrate_tensor_2d(heatmaps, bool(softmax), float(multiplier))


def _integrate_tensor_2d_impl(heatmaps, softmax, multiplier):
    batch_size, n_heatmaps, h, w = heatmaps.shape
    hm = heatmaps.detach()
    if hm.dtype != torch.float32 or not hm.is_contiguous():
        hm = hm.float().contiguous()
    hm = _aligned16(hm)
    coords = torch.empty((batch_size, n_heatmaps, 2), dtype=torch.float32, device=hm.device)
    out = torch.empty_like(hm)
    with _on(hm):
        _lib.check(_lib.load().egr_integrate_tensor_2d(_ptr(hm), batch_size, n_heatmaps, h, w, int(bool(softmax)), float(multiplier),
                                                       _ptr(coords), _ptr(out), _stream(hm)))
    return coords, out


# ------------------------------------------------------------------------------------------------
# I1  dataset preprocessing
# ------------------------------------------------------------------------------------------------
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def preprocess_images(images, size=(256, 256), mean=IMAGENET_MEAN, std=IMAGENET_STD, return_resized=False):
    """What every reference dataset does per view (e.g. datasets/ego4view_rw/ego4view_rw_heatmap_mvf.py:40-41,96-99):
    `Normalize(mean, std)(ToTensor()(img.resize(size, Image.BICUBIC)))`, batched and bit-exact with PIL + torchvision.

    images: uint8 CUDA tensor [..., H, W, 3] of decoded RGB images (any leading dims, e.g. [B, V]).
    -> float32 [..., 3, size[1], size[0]]; with return_resized also the resized uint8 images [..., size[1], size[0], 3].
    `size` is (width, height) like PIL's."""
    _need_cuda(images, "preprocess_images")
    if images.dtype != torch.uint8 or images.ndim < 3 or images.shape[-1] != 3:
        raise RuntimeError("egorear_b200.preprocess_images: expected uint8 [..., H, W, 3]")
    lead = tuple(images.shape[:-3])
    Hin, Win = int(images.shape[-3]), int(images.shape[-2])
    Wout, Hout = int(size[0]), int(size[1])
    img = images.detach().contiguous().view(-1, Hin, Win, 3)
    N = img.shape[0]
    out = torch.empty((N, 3, Hout, Wout), dtype=torch.float32, device=img.device)
    res = torch.empty((N, Hout, Wout, 3), dtype=torch.uint8, device=img.device) if return_resized else None
    m = (ctypes.c_float * 3)(*[float(v) for v in mean])
    s = (ctypes.c_float * 3)(*[float(v) for v in std])
    lib = _lib.load()
    for i in range(0, max(N, 1), 65535):
        n = min(65535, N - i)
        with _on(img):
            _lib.check(lib.egr_preprocess_images(_ptr(img[i:]) if n > 0 else None, n, Hin, Win, Hout, Wout, m, s,
                                                 _ptr(out[i:]) if n > 0 else None, _ptr(res[i:]) if res is not None and n > 0 else None,
                                                 _stream(img)))
    out = out.view(lead + (3, Hout, Wout))
    return (out, res.view(lead + (Hout, Wout, 3))) if return_resized else out


# ------------------------------------------------------------------------------------------------
# G1
# ------------------------------------------------------------------------------------------------
def _numpy_patch(sigma):
    # the caller-side Gaussian patch, evaluated by numpy in float32 exactly like generate_heatmap.py:32-36
    tmp_size = sigma * 3
    size = 2 * tmp_size + 1
    x = np.arange(0, size, 1, np.float32)
    y = x[:, np.newaxis]
    x0 = y0 = size // 2
    return np.ascontiguousarray(np.exp(-((x - x0) ** 2 + (y - y0) ** 2) / (2 * sigma ** 2)), dtype=np.float32)


def generate_target_batch(joints, image_size=872, heatmap_size=64, sigma=1, out=None):
    """Batched on-GPU Gaussian heatmap synthesis.

    joints: float64 [..., J, 2] (x, y) pixels — CUDA tensor, CPU tensor or numpy array.
    Returns a CUDA float32 tensor [..., J, heatmap_size, heatmap_size] (or writes into `out`).
    """
    if isinstance(joints, np.ndarray):
        joints = torch.from_numpy(np.ascontiguousarray(joints, dtype=np.float64))
    joints = joints.to(dtype=torch.float64)
    if not joints.is_cuda:
        joints = joints.cuda(non_blocking=True)
    joints = joints.contiguous()
    assert joints.shape[-1] == 2 and joints.ndim >= 2
    if out is None:
        return torch.ops.egr.generate_target(joints, float(image_size), int(heatmap_size), float(sigma))
    return _generate_target_impl(joints, image_size, heatmap_size, sigma, out)


def _generate_target_impl(joints, image_size, heatmap_size, sigma, out):
    J = joints.shape[-2]
    n_maps = joints.numel() // (J * 2)
    shape = tuple(joints.shape[:-1]) + (heatmap_size, heatmap_size)
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=joints.device)
    else:
        assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.numel() == n_maps * J * heatmap_size ** 2
        _same_device("generate_target_batch", joints, out)
    patch = _numpy_patch(sigma)
    lib = _lib.load()
    with _on(joints):
        _lib.check(lib.egr_generate_target(_ptr(joints), _ptr(out), n_maps, J, float(image_size), int(heatmap_size),
                                           float(sigma), ctypes.c_void_p(patch.ctypes.data), _stream(joints)))
    return out


def generate_target(joints, image_size=872, heatmap_size=64, num_joints=15, sigma=1):
    """Drop-in for generate_heatmap.generate_target: numpy in, numpy float32 [num_joints, hs, hs] out."""
    j = np.asarray([[joints[i][0], joints[i][1]] for i in range(num_joints)], dtype=np.float64)
    return generate_target_batch(j, image_size, heatmap_size, sigma).cpu().numpy()


# ------------------------------------------------------------------------------------------------
# MSDA (mmcv op replacement)
# ------------------------------------------------------------------------------------------------
def ms_deform_attn(value, spatial_shapes, level_start_index, sampling_locations, attention_weights, im2col_step=None):
    """Forward of mmcv's MultiScaleDeformableAttnFunction for one feature level (the only case the
    reference uses: DeformMultiViewAttn / DeformStereoAttn register spatial_shapes of shape (1, 2))."""
    _need_cuda(value, "ms_deform_attn")
    B, L, nh, hd = value.shape
    _, Q, _, n_levels, P, _ = sampling_locations.shape
    if n_levels != 1:
        raise NotImplementedError("egorear_b200.ms_deform_attn: single-level only (n_levels=%d)" % n_levels)
    if isinstance(spatial_shapes, torch.Tensor) and not torch.compiler.is_compiling():
        H, W = int(spatial_shapes[0, 0]), int(spatial_shapes[0, 1])
    elif isinstance(spatial_shapes, torch.Tensor):
        # under torch.compile reading the buffer would be a graph break: the reference only uses square single-level maps
        H = W = int(round(L ** 0.5))
    else:
        H, W = int(spatial_shapes[0][0]), int(spatial_shapes[0][1])
    assert H * W == L
    return torch.ops.egr.msda_forward(value, H, W, sampling_locations, attention_weights)


def _msda_impl(value, H, W, sampling_locations, attention_weights):
    B, L, nh, hd = value.shape
    _, Q, _, n_levels, P, _ = sampling_locations.shape
    value = value.float().contiguous()
    loc = sampling_locations.float().contiguous()
    aw = attention_weights.float().contiguous()
    _same_device("ms_deform_attn", value, loc, aw)
    out = torch.empty((B, Q, nh * hd), dtype=torch.float32, device=value.device)
    lib = _lib.load()
    with _on(value):
        _lib.check(lib.egr_msda_forward(_ptr(value), B, H, W, nh, hd, _ptr(loc), _ptr(aw), Q, P, _ptr(out), _stream(value)))
    return out


class MultiScaleDeformableAttnFunction:
    """Inference stand-in with mmcv's call signature: `MultiScaleDeformableAttnFunction.apply(...)`."""

    @staticmethod
    def apply(value, spatial_shapes, level_start_index, sampling_locations, attention_weights, im2col_step):
        return ms_deform_attn(value, spatial_shapes, level_start_index, sampling_locations, attention_weights,
                              im2col_step)


# ------------------------------------------------------------------------------------------------
# P3
# ------------------------------------------------------------------------------------------------
def calib_table(calib):
    """{camera_name: {...}} -> float32 [4,16] host table for the C-ABI."""
    t = np.zeros((4, 16), dtype=np.float32)
    for i, name in enumerate(CAMERA_NAMES):
        c = calib[name]
        poly = c["polynomialW2C"]
        assert len(poly) <= 11
        t[i, 0:2] = c["image_center"]
        t[i, 2:4] = c["size"]
        t[i, 4] = len(poly)
        t[i, 5:5 + len(poly)] = poly
    return t


def reproject_fisheye(pts3d, camera_model, coord_trans_mat=None, calib=None):
    """anchors_2d [B,V,J,2], anchors_valid [B,V,J] bool.  `pts3d` [B,J,3] fp32 is mutated in place for the
    synthetic rigs, reproducing utils/camera_models.py:57-63."""
    _need_cuda(pts3d, "reproject_fisheye")
    assert pts3d.dtype == torch.float32 and pts3d.is_contiguous() and pts3d.ndim == 3
    names = cameras_for(camera_model)
    ids = (ctypes.c_int * len(names))(*[CAMERA_NAMES.index(n) for n in names])
    is_rw = camera_model.startswith("ego4view_rw")
    B, J = pts3d.shape[:2]
    if is_rw:
        assert coord_trans_mat is not None
        if coord_trans_mat.dtype != torch.float32:
            # utils/camera_models.py:210 matmuls the matrix against fp32 points: a float64 matrix is a dtype error there
            raise RuntimeError("expected m1 and m2 to have the same dtype, but got: double != float")
        if not coord_trans_mat.is_cuda or coord_trans_mat.device != pts3d.device:
            raise RuntimeError("egorear_b200.reproject_fisheye: coord_trans_mat must live on pts3d's CUDA device")
        if tuple(coord_trans_mat.shape) != (B, len(names), 4, 4):
            # the kernel strides by the model's camera count (the reference indexes [:, v])
            raise RuntimeError("egorear_b200.reproject_fisheye: coord_trans_mat must be [B=%d, n_cams=%d, 4, 4], got %s"
                               % (B, len(names), tuple(coord_trans_mat.shape)))
        coord_trans_mat = coord_trans_mat.contiguous()
    a2 = torch.empty((B, len(names), J, 2), dtype=torch.float32, device=pts3d.device)
    av = torch.empty((B, len(names), J), dtype=torch.bool, device=pts3d.device)
    tab = calib_table(calib) if calib is not None else None
    lib = _lib.load()
    with _on(pts3d):
        _lib.check(lib.egr_reproject_fisheye(_ptr(pts3d), B, J, ids, len(names), int(is_rw),
                                             _ptr(coord_trans_mat) if is_rw else None,
                                             ctypes.c_void_p(tab.ctypes.data) if tab is not None else None,
                                             _ptr(a2), _ptr(av), _stream(pts3d)))
    return a2, av


def heatmap_head_1x1(feat, weight, bias):
    """1x1 conv head of EgoPoseFormerHeatmap on [N,C,H,W] fp32 (estimator/egoposeformer_heatmap.py:34-39)."""
    _need_cuda(feat, "heatmap_head_1x1")
    return torch.ops.egr.heatmap_head_1x1(feat, weight, bias)


def _heatmap_head_1x1_impl(feat, weight, bias):
    N, C, H, W = feat.shape
    J = weight.shape[0]
    feat = _aligned16(feat.float().contiguous())
    w = weight.detach().reshape(J, C).float().contiguous()
    b = bias.detach().float().contiguous()
    _same_device("heatmap_head_1x1", feat, w, b)
    out = torch.empty((N, J, H, W), dtype=torch.float32, device=feat.device)
    lib = _lib.load()
    with _on(feat):
        _lib.check(lib.egr_heatmap_head_1x1(_ptr(feat), _ptr(w), _ptr(b), N, C, H * W, J, _ptr(out), _stream(feat)))
    return out


def pack_joints(preds2d, pose3d):
    """[B, ...] 2D joints + [B, J3, 3] pose -> packed fp32 [B, n2d+n3d] row per frame (all-gather payload)."""
    return torch.ops.egr.pack_joints(preds2d, pose3d)


def _pack_joints_impl(preds2d, pose3d):
    B = preds2d.shape[0]
    p2 = preds2d.reshape(B, -1).float().contiguous()
    p3 = pose3d.reshape(B, -1).float().contiguous()
    _same_device("pack_joints", p2, p3)
    out = torch.empty((B, p2.shape[1] + p3.shape[1]), dtype=torch.float32, device=p2.device)
    lib = _lib.load()
    with _on(p2):
        _lib.check(lib.egr_pack_joints(_ptr(p2), _ptr(p3), B, p2.shape[1], p3.shape[1], _ptr(out), _stream(p2)))
    return out
