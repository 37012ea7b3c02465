import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden():
    return {name: np.load(os.path.join(GOLDEN, name + ".npz")) for name in ("generate_target", "get_max_preds", "get_max_preds_soft", "models", "eval_metrics", "integrate_tensor_2d", "preprocess", "pose3d_stereo")}


def soft_inputs():
    """the heatmaps of tests/golden/make_golden.py::soft_inputs (kept in sync by test_oracle.py)"""
    import torch
    g = torch.Generator().manual_seed(12)
    hm = torch.randn((4, 15, 64, 64), generator=g)
    hm[1] *= 8.0
    hm[2] *= 0.05
    hm[3, 0] = 0.25
    hm[3, 1] = 0.0; hm[3, 1, 5, 60] = 30.0
    hm[3, 2] = -50.0; hm[3, 2, 63, 0] = 40.0; hm[3, 2, 0, 63] = 40.0
    return hm


PREPROCESS_CASES = ((872, 872, 256, 256), (480, 640, 256, 256), (100, 75, 256, 256), (256, 256, 256, 256), (301, 257, 64, 96))
INTEGRATE_CASES = (("sm100", True, 100.0), ("sm1", True, 1.0), ("relu", False, 100.0))
INTEGRATE_MAPS = ((0, 0), (1, 5), (2, 7), (3, 1), (3, 2), (3, 3))


def integrate_inputs():
    """tests/golden/make_golden.py::integrate_inputs"""
    hm = soft_inputs()
    hm[3, 3] = -1.0
    return hm


@pytest.fixture(scope="session")
def oracle_lib():
    """The C oracle (oracle/gt_decode.c), built on demand.  Test infrastructure only."""
    import ctypes
    import subprocess
    so = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
    src = os.path.join(ROOT, "oracle", "gt_decode.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    lib = ctypes.CDLL(so)
    lib.orc_gaussian_patch.argtypes = [ctypes.c_double, ctypes.c_int, ctypes.c_void_p]
    lib.orc_generate_target.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_double,
                                        ctypes.c_int, ctypes.c_double, ctypes.c_void_p]
    lib.orc_get_max_preds.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_float, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_void_p]
    return lib


def orc_generate_target(lib, joints, image_size=872.0, hs=64, sigma=1.0):
    joints = np.ascontiguousarray(joints, dtype=np.float64)
    J = joints.shape[-2]
    n = joints.size // (2 * J)
    size = 2 * int(3 * sigma) + 1
    patch = np.zeros((size, size), dtype=np.float32)
    lib.orc_gaussian_patch(float(sigma), size, patch.ctypes.data)
    out = np.empty(joints.shape[:-1] + (hs, hs), dtype=np.float32)
    lib.orc_generate_target(joints.ctypes.data, out.ctypes.data, n, J, float(image_size), hs, float(sigma), patch.ctypes.data)
    return out


def orc_get_max_preds(lib, hm, threshold, normalize):
    hm = np.ascontiguousarray(hm, dtype=np.float32)
    N, J, H, W = hm.shape
    preds = np.empty((N, J, 2), np.float32)
    maxv = np.empty((N, J), np.float32)
    valid = np.empty((N, J), np.uint8)
    idx = np.empty((N, J), np.int32)
    lib.orc_get_max_preds(hm.ctypes.data, N, J, H, W, float(threshold), int(normalize), preds.ctypes.data,
                          maxv.ctypes.data, valid.ctypes.data, idx.ctypes.data)
    return preds, maxv, valid.astype(bool), idx


def dense_from_sparse(idx, val, shape):
    out = np.zeros(int(np.prod(shape)), dtype=np.float32)
    out[idx] = val
    return out.reshape(tuple(shape))


def rel_err(a, b):
    """max |a-b| / max |b|  — the 'relative' of the parity bounds (BASELINE.json: 1e-3 in fp32)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
