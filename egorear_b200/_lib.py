"""ctypes binding of libegorear_b200.so (the C-ABI declared in include/egorear_b200.h).

This is the stub a maintainer of the reference would add.  It fails loudly:
a missing library or a non-B200 device raises, there is no CPU / eager fallback anywhere in the
package.
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_uint32, c_void_p, POINTER

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libegorear_b200.so")

_lib = None


class DenseDesc(ctypes.Structure):
    """struct egr_dense_desc (include/egorear_b200.h)"""
    _fields_ = [("A", c_void_p), ("W", c_void_p), ("bias", c_void_p), ("D", c_void_p), ("aux", c_void_p),
                ("M", ctypes.c_int32), ("N", ctypes.c_int32), ("K", ctypes.c_int32),
                ("lda", c_int64), ("ldd", c_int64),
                ("amode", ctypes.c_int32), ("epi", ctypes.c_int32),
                ("kblk", ctypes.c_int32), ("kblk_stride", c_int64),
                ("Hin", ctypes.c_int32), ("Win", ctypes.c_int32), ("Cin", ctypes.c_int32),
                ("Hout", ctypes.c_int32), ("Wout", ctypes.c_int32),
                ("groups", ctypes.c_int32),
                ("a_gs", c_int64), ("w_gs", c_int64), ("b_gs", c_int64), ("d_gs", c_int64), ("aux_gs", c_int64),
                ("a_is_bf16", ctypes.c_int32), ("d_is_bf16", ctypes.c_int32), ("use_tc", ctypes.c_int32),
                ("ka", ctypes.c_int32),
                ("out_pair", ctypes.c_int32)]


class EgrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("egorear_b200 error %d: %s" % (code, msg))
        self.code = code


# name -> (restype, argtypes); kept in one table so tests can check it against the header
SIGNATURES = {
    "egr_last_error": (c_char_p, []),
    "egr_version": (c_int, []),
    "egr_profile_enable": (c_int, [c_int]),
    "egr_profile_read": (c_int, [ctypes.c_char_p, c_int]),
    "egr_device_check": (c_int, [POINTER(c_int), POINTER(c_int)]),
    "egr_launch_count": (c_int64, []),
    "egr_generate_target": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_double, c_int, c_double, c_void_p, c_void_p]),
    "egr_decode_soft_argmax": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "egr_integrate_tensor_2d": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "egr_decode_argmax": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_float, c_int, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p]),
    "egr_msda_forward": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int,
                                 c_void_p, c_void_p]),
    "egr_msda_backward": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p]),
    "egr_loss_workspace_bytes": (c_int64, []),
    "egr_mse_loss_forward": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "egr_mse_loss_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "egr_mpjpe_loss_forward": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "egr_mpjpe_loss_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "egr_reproject_fisheye": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p]),
    "egr_heatmap_head_1x1": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p]),
    "egr_dense_stage": (c_int, [POINTER(DenseDesc), c_void_p]),
    "egr_up2_relu_stage": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p]),
    "egr_head_tail_stage": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int64, c_int64,
                                    c_void_p, c_int, c_int, c_void_p]),
    "egr_mvfex_create": (c_int, [c_int, c_int, c_float, c_int, POINTER(c_void_p)]),
    "egr_mvfex_destroy": (c_int, [c_void_p]),
    "egr_mvfex_set_param": (c_int, [c_void_p, c_char_p, c_void_p, c_int64]),
    "egr_mvfex_prepack": (c_int, [c_void_p, c_void_p]),
    "egr_mvfex_workspace_bytes": (c_int64, [c_void_p, c_int]),
    "egr_mvfex_forward": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "egr_mvfex_refiner_forward": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "egr_mvfex_export_staged": (c_int, [c_void_p, c_int]),
    "egr_mvfex_use_staged_input": (c_int, [c_void_p, c_void_p]),
    "egr_mvfex_staged": (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), POINTER(c_int)]),
    "egr_pose3d_use_staged": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "egr_pose3d_use_staged_final_f16": (c_int, [c_void_p, c_void_p]),
    "egr_pose3d_proposal_dtype": (c_int, [c_void_p]),
    "egr_mvfex_debug_buffer": (c_int, [c_void_p, c_char_p, POINTER(c_void_p), POINTER(c_int64)]),
    "egr_pose3d_debug_buffer": (c_int, [c_void_p, c_char_p, POINTER(c_void_p), POINTER(c_int64)]),
    "egr_set_option": (c_int, [c_char_p, c_int]),
    "egr_pose3d_create": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, POINTER(c_void_p)]),
    "egr_pose3d_destroy": (c_int, [c_void_p]),
    "egr_pose3d_set_param": (c_int, [c_void_p, c_char_p, c_void_p, c_int64]),
    "egr_pose3d_prepack": (c_int, [c_void_p, c_void_p]),
    "egr_pose3d_workspace_bytes": (c_int64, [c_void_p, c_int]),
    "egr_pose3d_forward": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                   c_void_p]),
    "egr_pack_joints": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "egr_backbone_create": (c_int, [c_int, c_int, POINTER(c_void_p)]),
    "egr_backbone_destroy": (c_int, [c_void_p]),
    "egr_backbone_set_param": (c_int, [c_void_p, c_char_p, c_void_p, c_int64]),
    "egr_backbone_prepack": (c_int, [c_void_p, c_void_p]),
    "egr_backbone_workspace_bytes": (c_int64, [c_void_p, c_int]),
    "egr_backbone_forward": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "egr_resample_coeffs": (c_int, [c_int, c_int, POINTER(c_int), c_void_p, c_void_p]),
    "egr_resample_digits": (c_int, [c_int, c_int, POINTER(c_int), c_void_p]),
    "egr_preprocess_images": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_int, POINTER(c_float), POINTER(c_float),
                                      c_void_p, c_void_p, c_void_p]),
    "egr_eval_heatmap_workspace_bytes": (ctypes.c_size_t, [c_int64, c_int, c_int]),
    "egr_eval_heatmap": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_int, c_float,
                                 c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_size_t, c_void_p]),
    "egr_eval_pose": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_float, c_float, c_void_p, c_int, c_void_p, c_void_p,
                              c_void_p]),
}


def load():
    """Load the shared library once; raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "egorear_b200: %s not found. Build it with `python -m egorear_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the .so is stale: loud on purpose
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise EgrError(rc, load().egr_last_error().decode("utf-8", "replace"))


def launch_count():
    return int(load().egr_launch_count())
