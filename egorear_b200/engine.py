"""Handle wrappers around the two C-ABI engines (egr_mvfex_*, egr_pose3d_*).

The wrappers own nothing but (a) references to the parameter tensors they registered (so the device
pointers stay alive), (b) a cached workspace tensor.  Derived weights live in the C library; they are
rebuilt (`prepack`) lazily whenever a registered parameter changed (`load_state_dict`, optimizer step),
detected through the tensors' version counters and data pointers.
"""
import ctypes
import os

import torch

from . import _lib
from .ops import _ptr, _stream, calib_table

PREC = {"fp32": 0, "bf16": 1, "fp16": 3}      # EGR_PREC_* (include/egorear_b200.h)
ACT_DTYPE = {"bf16": torch.bfloat16, "fp16": torch.float16}
CAMERA_MODEL_ID = {"ego4view_syn": 0, "ego4view_rw": 1, "ego4view_syn_stereo_front": 2, "ego4view_rw_stereo_front": 3,
                   "ego4view_syn_stereo_back": 4, "ego4view_rw_stereo_back": 5}


def set_option(key, value):
    _lib.check(_lib.load().egr_set_option(key.encode(), int(value)))


if os.environ.get("EGR_TC") == "0":      # debugging: bf16 activations through the SIMT GEMM instead of tcgen05
    set_option("tc", 0)


class _EngineBase:
    _prefix = None

    def __init__(self):
        self._lib = _lib.load()
        self._h = ctypes.c_void_p()
        self._tensors = {}
        self._sig = None
        self._ws_lanes = {}          # lane -> cached workspace tensor (concurrent forwards on different streams)
        self.lane = 0                # lane of the LAST forward (debug_buffer reads its workspace); forwards take `lane=` explicitly
        self.frozen = False
        self._calls = 0

    def _fn(self, name):
        return getattr(self._lib, "egr_%s_%s" % (self._prefix, name))

    def close(self):
        if self._h:
            self._fn("destroy")(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- parameters ----
    def _signature(self):
        return tuple((t.data_ptr(), t._version) for t in self._tensors.values())

    def set_params(self, named_tensors):
        """named_tensors: {state_dict key: fp32 CUDA tensor}.  Non-contiguous / non-fp32 tensors are copied."""
        self._tensors = {}
        for k, t in named_tensors.items():
            if not t.is_cuda:
                raise RuntimeError("egorear_b200: parameter '%s' is not on a CUDA device (no CPU fallback)" % k)
            self._tensors[k] = t
        self._sig = None

    def invalidate(self):
        """parameters were changed behind the module's back (in-place edits outside load_state_dict / .to()): re-pack"""
        self._sig = None

    def _sync_params(self):
        """Re-register + re-pack when a parameter changed.  The owning module invalidates on load_state_dict / .to() /
        .cuda() / .float() (`_EngineOwner`); on top of that the version counters of all tensors (~570 for mvfex) are
        compared on the first call and then every 32nd one, so that an in-place edit is not missed for long while the
        common eval loop does not walk the list per forward.  frozen: never re-checked."""
        self._calls += 1
        if self._sig is not None and (self.frozen or self._calls % 32 != 1):
            return
        sig = self._signature()
        if sig == self._sig:
            return
        keep = []
        for k, t in self._tensors.items():
            d = t.detach()
            if d.dtype != torch.float32 or not d.is_contiguous():
                d = d.float().contiguous()
                keep.append(d)
            _lib.check(self._fn("set_param")(self._h, k.encode(), _ptr(d), d.numel()))
        self._keep = keep
        any_t = next(iter(self._tensors.values()))
        with torch.cuda.device_of(any_t):              # the parameters' device, not whatever is current
            _lib.check(self._fn("prepack")(self._h, _stream(any_t)))
        self._sig = sig

    def _workspace(self, B, device, lane=0):
        need = int(self._fn("workspace_bytes")(self._h, B))
        self.lane = lane
        ws = self._ws_lanes.get(lane)
        if ws is None or ws.numel() < need or ws.device != device:
            ws = self._ws_lanes[lane] = torch.empty(need, dtype=torch.uint8, device=device)
        return ws

    @property
    def _ws(self):
        return self._ws_lanes.get(self.lane)

    def debug_buffer(self, name, dtype, shape):
        """Copy of a named intermediate of the last forward (tests only)."""
        p = ctypes.c_void_p()
        n = ctypes.c_int64()
        _lib.check(self._fn("debug_buffer")(self._h, name.encode(), ctypes.byref(p), ctypes.byref(n)))
        off = p.value - self._ws.data_ptr()
        raw = self._ws[off:off + n.value]
        return raw.view(dtype)[: int(torch.tensor(shape).prod())].view(*shape).clone()


class MvfexEngine(_EngineBase):
    """Everything EgoPoseFormerHeatmapMVFEX.forward does after the backbones."""
    _prefix = "mvfex"

    def __init__(self, num_views=4, num_heatmap=15, heatmap_threshold=0.5, precision="fp16"):
        super().__init__()
        self.V, self.J, self.precision = num_views, num_heatmap, precision
        _lib.check(self._lib.egr_mvfex_create(num_views, num_heatmap, float(heatmap_threshold), PREC[precision],
                                              ctypes.byref(self._h)))

    def export_staged(self, enable=True, hp="tf32"):
        """keep channels-last copies of the input / refined features for a chained Pose3DEngine.forward(staged=...);
        hp: high-precision copy of the refined features for pose3d's proposal branch: "tf32" (fp32), "f16" or None"""
        mode = {"tf32": 1, None: 2, "f16": 3, "f16_only": 4}[hp] if enable else 0
        _lib.check(self._lib.egr_mvfex_export_staged(self._h, mode))
        self._export, self._export_hp = bool(enable), hp

    def forward(self, feat, bfb, heatmap_for_anchor=None, want_feat_refined=True, feat_staged=None, lane=0):
        """feat [B,V,128,64,64], bfb [B,V,512,8,8] fp32 CUDA ->
        dict(hm_init, hm_refined [B,V,15,64,64], feat_refined [B,V,128,64,64], anchors_2d [B,V,15,2], anchors_valid).
        want_feat_refined=False (needs export_staged): the NCHW fp32 refined features are not materialised (None).
        feat_staged (bf16 precision only): the features as a channels-last bf16 producer leaves them, view-major
        [V,B,64,64,128] contiguous; `feat` may then be None and the staging pass is skipped (egr_mvfex_use_staged_input)."""
        self._sync_params()
        feat_arg = feat
        if feat_staged is not None:
            V, B = feat_staged.shape[:2]
            if self.precision not in ACT_DTYPE:
                raise RuntimeError("feat_staged needs the bf16 precision (bf16 copy) or the fp16 precision (fp16 copy)")
            assert feat_staged.dtype == ACT_DTYPE[self.precision] and feat_staged.is_contiguous() and tuple(feat_staged.shape[2:]) == (64, 64, 128)
            feat = None
        else:
            B, V = feat.shape[:2]
            assert tuple(feat.shape[2:]) == (128, 64, 64)
            feat = feat.detach().float().contiguous()
        assert V == self.V and tuple(bfb.shape) == (B, V, 512, 8, 8)
        bfb = bfb.detach().float().contiguous()
        hfa = heatmap_for_anchor.detach().float().contiguous() if isinstance(heatmap_for_anchor, torch.Tensor) else None
        dev = bfb.device
        out = {
            "hm_init": torch.empty((B, V, self.J, 64, 64), dtype=torch.float32, device=dev),
            "hm_refined": torch.empty((B, V, self.J, 64, 64), dtype=torch.float32, device=dev),
            "feat_refined": torch.empty((B, V, 128, 64, 64), dtype=torch.float32, device=dev) if want_feat_refined else None,
            "anchors_2d": torch.empty((B, V, self.J, 2), dtype=torch.float32, device=dev),
            "anchors_valid": torch.empty((B, V, self.J), dtype=torch.bool, device=dev),
        }
        ws = self._workspace(B, dev, lane)
        if feat_staged is not None:
            _lib.check(self._lib.egr_mvfex_use_staged_input(self._h, _ptr(feat_staged)))
            out["_keepalive"] = feat_staged          # the staged copies exported to a chained pose3d point into it
        with torch.cuda.device_of(bfb):
            _lib.check(self._lib.egr_mvfex_forward(self._h, B, _ptr(feat), _ptr(bfb), _ptr(hfa), _ptr(out["hm_init"]),
                                                   _ptr(out["hm_refined"]), _ptr(out["feat_refined"]),
                                                   _ptr(out["anchors_2d"]), _ptr(out["anchors_valid"]), _ptr(ws),
                                                   ws.numel(), _stream(bfb)))
        if getattr(self, "_export", False):
            pi, pr, pt, bf = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int()
            _lib.check(self._lib.egr_mvfex_staged(self._h, ctypes.byref(pi), ctypes.byref(pr), ctypes.byref(pt), ctypes.byref(bf)))
            # device pointers into this engine's workspace: valid until its next forward
            hp = getattr(self, "_export_hp", "tf32")
            out["staged"] = {"init": pi.value, "refined": pr.value, "refined_tf32": pt.value if hp == "tf32" else None,
                             "refined_f16": pt.value if hp in ("f16", "f16_only") else None, "bf16": bf.value,
                             "feat": None if feat_staged is not None else feat_arg, "feat_refined": out["feat_refined"],
                             "B": B, "device": dev, "keepalive": feat_staged}
        return out

    def refiner_forward(self, r, heatmap, frame_feat, feat_mv, anchors_2d, anchors_valid, bfb):
        """One HeatmapMVF.forward: returns (hm_refined [B,15,64,64], feat_refined [B,128,64,64])."""
        self._sync_params()
        B = heatmap.shape[0]
        dev = heatmap.device
        heatmap = heatmap.detach().float().contiguous()
        frame_feat = frame_feat.detach().float().contiguous()
        feat_mv = feat_mv.detach().float().contiguous()
        a2 = anchors_2d.detach().float().contiguous()
        av = anchors_valid.detach().to(torch.bool).contiguous()
        bfb = bfb.detach().float().contiguous()
        hm = torch.empty((B, self.J, 64, 64), dtype=torch.float32, device=dev)
        ft = torch.empty((B, 128, 64, 64), dtype=torch.float32, device=dev)
        ws = self._workspace(B, dev)
        with torch.cuda.device_of(heatmap):
            _lib.check(self._lib.egr_mvfex_refiner_forward(self._h, r, B, _ptr(heatmap), _ptr(frame_feat), _ptr(feat_mv),
                                                           _ptr(a2), _ptr(av), _ptr(bfb), _ptr(hm), _ptr(ft), _ptr(ws),
                                                           ws.numel(), _stream(heatmap)))
        return hm, ft


class BackboneEngine(_EngineBase):
    """ResnetBackbone.forward of the stereo estimators on the tcgen05 conv stages: images in, the staged
    16-bit channels-last FPN map + the fp32 stride-32 map out (egr_backbone_*)."""
    _prefix = "backbone"

    def __init__(self, num_views=4, precision="fp16"):
        super().__init__()
        if precision not in ACT_DTYPE:
            raise RuntimeError("egorear_b200: the backbone engine runs in the tensor-core precisions (bf16 / fp16); fp32 keeps the PyTorch backbone")
        self.V, self.precision = num_views, precision
        _lib.check(self._lib.egr_backbone_create(num_views, PREC[precision], ctypes.byref(self._h)))

    def forward(self, img, lane=0):
        """img [B,V,3,256,256] fp32 CUDA -> (feat_staged [V,B,64,64,128] 16-bit, bfb [B,V,512,8,8] fp32)"""
        self._sync_params()
        B, V = img.shape[:2]
        assert V == self.V and tuple(img.shape[2:]) == (3, 256, 256), "backbone engine: img must be [B, %d, 3, 256, 256]" % self.V
        img = img.detach().float().contiguous()
        dev = img.device
        feat = torch.empty((V, B, 64, 64, 128), dtype=ACT_DTYPE[self.precision], device=dev)
        bfb = torch.empty((B, V, 512, 8, 8), dtype=torch.float32, device=dev)
        ws = self._workspace(B, dev, lane)
        with torch.cuda.device_of(img):
            _lib.check(self._lib.egr_backbone_forward(self._h, B, _ptr(img), _ptr(feat), _ptr(bfb), _ptr(ws), ws.numel(), _stream(img)))
        return feat, bfb


class Pose3DEngine(_EngineBase):
    """EgoPoseFormerPose3D.forward."""
    _prefix = "pose3d"

    def __init__(self, num_views=4, num_joints=16, num_layers=3, camera_model="ego4view_syn", use_pred_heatmap_init=True,
                 precision="fp16", calib=None):
        super().__init__()
        if camera_model not in CAMERA_MODEL_ID:
            raise ValueError('Unknown camera model !')
        self.V, self.J, self.L = num_views, num_joints, num_layers
        self.camera_model, self.precision = camera_model, precision
        tab = calib_table(calib) if calib is not None else None
        _lib.check(self._lib.egr_pose3d_create(num_views, num_joints, num_layers, CAMERA_MODEL_ID[camera_model],
                                               int(bool(use_pred_heatmap_init)), PREC[precision],
                                               ctypes.c_void_p(tab.ctypes.data) if tab is not None else None,
                                               ctypes.byref(self._h)))

    def proposal_dtype(self):
        """operand type of the proposal branch: "fp32", "bf16", "tf32" or "f16" """
        return ("fp32", "bf16", "tf32", "f16")[int(self._lib.egr_pose3d_proposal_dtype(self._h))]

    def forward(self, feats_init, feats_final, coord_trans_mat=None, staged=None, use_init=True, lane=0):
        """-> preds [L+1, B, 16, 3] fp32 (cm): preds[0] MLP proposal, preds[1:] transformer layers.
        staged: MvfexEngine.forward(...)["staged"] of the SAME tensors (chained forward): skips the re-staging passes."""
        self._sync_params()
        ref = feats_final if feats_final is not None else feats_init
        if ref is None:                    # fully chained: both feature maps only exist as the staged channels-last copies
            assert staged is not None, "pose3d: feats_init / feats_final may both be None only with staged copies"
            B, V, dev = staged["B"], self.V, staged["device"]
        else:
            B, V, dev = ref.shape[0], ref.shape[1], ref.device
        assert V == self.V
        fi = feats_init.detach().float().contiguous() if feats_init is not None else None
        ff = feats_final.detach().float().contiguous() if feats_final is not None else None
        ctm = None
        if self.camera_model.startswith("ego4view_rw"):
            if coord_trans_mat is None:
                raise RuntimeError("ego4view_rw camera model needs coord_trans_mat")
            if coord_trans_mat.dtype != torch.float32:
                # the reference matmuls this against fp32 points (utils/camera_models.py:210): dtype error there too
                raise RuntimeError("expected m1 and m2 to have the same dtype, but got: double != float")
            if not coord_trans_mat.is_cuda or coord_trans_mat.device != dev:
                raise RuntimeError("coord_trans_mat must live on the features' CUDA device (%s)" % dev)
            if tuple(coord_trans_mat.shape) != (B, V, 4, 4):
                # the kernel strides by the model's camera count; the reference indexes [:, v] and would fail differently
                raise RuntimeError("coord_trans_mat must be [B=%d, n_cams=%d, 4, 4], got %s" % (B, V, tuple(coord_trans_mat.shape)))
            ctm = coord_trans_mat.contiguous()
        preds = torch.empty((self.L + 1, B, self.J, 3), dtype=torch.float32, device=dev)
        ws = self._workspace(B, dev, lane)
        def same(a, b):      # the tensors cross an operator boundary: compare storage, not Python identity
            if a is None or b is None:
                return a is None and b is None
            return a.data_ptr() == b.data_ptr() and a.shape == b.shape
        if staged is not None and staged["B"] == B and same(staged["feat_refined"], feats_final) and \
                (same(staged["feat"], feats_init) or not use_init):
            sampled = staged["init"] if use_init else staged["refined"]
            _lib.check(self._lib.egr_pose3d_use_staged(self._h, ctypes.c_void_p(sampled), int(staged["bf16"]),
                                                       ctypes.c_void_p(staged["refined_tf32"]) if staged["refined_tf32"] else None))
            if staged.get("refined_f16"):
                _lib.check(self._lib.egr_pose3d_use_staged_final_f16(self._h, ctypes.c_void_p(staged["refined_f16"])))
        with torch.cuda.device_of(preds):
            _lib.check(self._lib.egr_pose3d_forward(self._h, B, _ptr(fi), _ptr(ff), _ptr(ctm), _ptr(preds), _ptr(ws),
                                                    ws.numel(), _stream(preds)))
        return preds
