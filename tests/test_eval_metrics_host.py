"""CPU: eval-time metrics.

  * oracle/metrics_ref.py is pinned to tests/golden/eval_metrics.npz (produced by the live reference's wrapper methods);
  * the per-sample arithmetic the CUDA kernel runs (egorear_b200/csrc/eval_pose_math.cuh: float32 Jacobi 3x3 SVD, Procrustes,
    MPJPE / PCK / AUC) is compiled for the host (tests/host/eval_pose_host.cpp, test infrastructure) and checked against
    the oracle and the golden vectors, including rank-deficient inputs;
  * patch() rebinds the wrappers' metric methods.
"""
import ctypes
import os
import subprocess
import sys
import types

import numpy as np
import pytest

from egorear_b200 import synth
from oracle import metrics_ref as mr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
POSE_KEYS = ("mpjpe", "pa_mpjpe", "pck_3d", "auc_3d")


def test_pose_oracle_matches_golden(golden):
    g = golden["eval_metrics"]
    pred, gt = synth.synth_eval_poses(64, 16, seed=0)
    o = mr.evaluate_pose(pred, gt)
    assert np.allclose(o["mpjpe"], g["pose_final_mpjpe"], rtol=1e-6, atol=1e-5)
    assert np.allclose(o["pa_mpjpe"], g["pose_final_pa_mpjpe"], rtol=1e-6, atol=1e-9)      # same numpy code path
    assert np.array_equal(o["pck_3d"], g["pose_final_pck_3d"])
    assert np.allclose(o["auc_3d"], g["pose_final_auc_3d"], rtol=1e-6)
    assert np.allclose(o["s1_hat"], g["pose_s1_hat"], rtol=0, atol=1e-9)
    # known answers: exact match and pure similarity align to ~0; the metrics are in mm / percent
    assert g["pose_final_mpjpe"][0] == 0 and g["pose_final_pa_mpjpe"][0] < 1e-3 and g["pose_final_pa_mpjpe"][3] < 1e-3
    assert g["pose_final_pck_3d"][0] == 100 and g["pose_final_auc_3d"][0] == 100


def test_heatmap_oracle_matches_golden(golden):
    g = golden["eval_metrics"]
    pred, gt = synth.synth_eval_heatmaps(6, 4, 15, seed=0)
    for tag, sl in (("front", slice(0, 2)), ("back", slice(2, 4)), ("all", slice(0, 4))):
        o = mr.evaluate_heatmap(pred[:, sl], gt[:, sl])
        for k, v in o.items():
            assert np.allclose(v, g["hm_%s_%s" % (tag, k)], rtol=2e-6), (tag, k)
    o = mr.evaluate_heatmap(pred, gt)
    assert np.allclose(o["l1_error_heatmap"], g["hm_p3d_l1_error_heatmap"], rtol=2e-6)
    assert np.allclose(o["pos_l1_error_heatmap"], g["hm_p3d_pos_l1_error_heatmap"], rtol=2e-6)
    # the inputs exercise the validity mask: some targets are all-zero / attenuated below the wrappers' threshold 1.0
    mx = gt.reshape(-1, 64 * 64).max(1)
    assert (mx == 0).any() and ((mx > 0) & (mx < 1)).any() and (mx >= 1).any()


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("host") / "eval_pose_host.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "host", "eval_pose_host.cpp")])
    lib = ctypes.CDLL(so)
    lib.eval_pose_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float,
                                   ctypes.c_float, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    lib.eval_pose_host.restype = None
    return lib


def run_host(lib, pred, gt, unit_scale=10.0, pck=150.0, thr=mr.AUC_THRESHOLDS):
    pred = np.ascontiguousarray(pred, np.float32)
    gt = np.ascontiguousarray(gt, np.float32)
    B, J, _ = pred.shape
    thr = np.asarray(thr, np.float32)
    m = np.zeros((B, 4), np.float64)
    s = np.zeros((B, J, 3), np.float64)
    lib.eval_pose_host(pred.ctypes.data, gt.ctypes.data, B, J, unit_scale, pck, thr.ctypes.data, len(thr), m.ctypes.data,
                       s.ctypes.data)
    return m, s


def test_kernel_arithmetic_matches_golden_and_oracle(golden, host_lib):
    g = golden["eval_metrics"]
    pred, gt = synth.synth_eval_poses(64, 16, seed=0)
    m, s = run_host(host_lib, pred, gt)
    assert np.allclose(m[:, 0], g["pose_final_mpjpe"], rtol=1e-6, atol=1e-5)
    # float32 SVD on both sides (LAPACK there, Jacobi here): agreement to float32 accuracy of the aligned points (cm)
    assert np.allclose(s, g["pose_s1_hat"], rtol=0, atol=2e-4)
    assert np.allclose(m[:, 1], g["pose_final_pa_mpjpe"], rtol=1e-5, atol=2e-3)
    assert np.array_equal(m[:, 2].astype(np.float32), g["pose_final_pck_3d"])
    assert np.allclose(m[:, 3], g["pose_final_auc_3d"], rtol=1e-6)
    for seed, J in ((1, 16), (2, 15), (3, 5), (4, 32)):
        pred, gt = synth.synth_eval_poses(200, J, seed=seed)
        m, s = run_host(host_lib, pred, gt)
        o = mr.evaluate_pose(pred, gt)
        assert np.allclose(s, o["s1_hat"], rtol=0, atol=5e-4)
        for i, k in enumerate(POSE_KEYS):
            assert np.allclose(m[:, i], o[k], rtol=1e-5, atol=5e-3), (seed, k)


def test_kernel_arithmetic_rank_deficient_and_reflection(host_lib):
    """coplanar / collinear sources and targets, mirror images (det(U V^T) = -1), against a float64 numpy Procrustes"""
    rng = np.random.default_rng(9)

    def ref64(S1, S2):
        return mr.compute_similarity_transform(S1.astype(np.float64), S2.astype(np.float64))

    gt = rng.uniform(-50, 50, (6, 16, 3)).astype(np.float32)
    pred = gt + rng.normal(0, 2, gt.shape).astype(np.float32)
    pred[0, :, 2] = 3.0                                    # coplanar source
    gt[1, :, 0] = -7.0                                     # coplanar target
    pred[2] = gt[2] * np.array([1, 1, -1], np.float32)     # mirror image: the best ROTATION is not the reflection
    pred[3] = (gt[3] @ np.diag([-1.0, 1.0, 1.0])).astype(np.float32) * 0.5 + 3
    pred[4, :, 1:] = 0.0                                   # collinear source: R is not unique, the residual is
    _, s = run_host(host_lib, pred, gt)
    for b in range(6):
        want = ref64(pred[b], gt[b])
        if b == 4:
            assert abs(np.linalg.norm(s[b] - gt[b], axis=-1).mean() - np.linalg.norm(want - gt[b], axis=-1).mean()) < 5e-5
        else:
            assert np.allclose(s[b], want, rtol=0, atol=5e-5), b     # float32 SVD (like the reference's), cm
    # the mirror cases really take the det < 0 branch: a reflection would fit exactly, the rotation cannot
    assert np.linalg.norm(s[2] - gt[2], axis=-1).mean() > 1.0


def test_thresholds_are_inclusive(host_lib):
    """distances equal to a threshold count as correct (`<=`, utils/loss.py:34)"""
    gt = np.zeros((1, 4, 3), np.float32)
    pred = np.zeros((1, 4, 3), np.float32)
    pred[0, :, 0] = [0.0, 0.5, 15.0, 15.5]                 # cm -> 0, 5, 150, 155 mm
    m, _ = run_host(host_lib, pred, gt)
    assert m[0, 2] == 75.0                                  # 3 of 4 within 150 mm
    want = mr.compute_auc_3d_batch(pred * np.float32(10), gt * np.float32(10))[0] * 100
    assert abs(m[0, 3] - want) < 1e-4
    assert abs(m[0, 3] - 100.0 * (31 + 30 + 1 + 0) / (31 * 4)) < 1e-4    # thresholds 0, 5, ..., 150


def test_patch_rebinds_metric_methods():
    import egorear_b200
    from egorear_b200 import metrics
    name = "pose_estimation.pl_wrappers.egoposeformer.pose_3d_mvf_ex"
    fake = types.ModuleType(name)

    class Pose3DMVFEXLightningModel:
        def evaluate_pose(self, pred_pose, gt_pose, prefix):
            return "reference"

        def evaluate_heatmap(self, pred_heatmap, gt_heatmap, prefix):
            return "reference"
    fake.Pose3DMVFEXLightningModel = Pose3DMVFEXLightningModel
    parents = ["pose_estimation", "pose_estimation.pl_wrappers", "pose_estimation.pl_wrappers.egoposeformer"]
    saved = {k: sys.modules.get(k) for k in parents + [name]}
    try:
        for k in parents:
            sys.modules[k] = types.ModuleType(k)
        sys.modules[name] = fake
        done = egorear_b200.patch(precision="bf16", modules=[name])
        assert "Pose3DMVFEXLightningModel.evaluate_pose" in done[name]
        assert Pose3DMVFEXLightningModel.evaluate_pose.__wrapped__ is metrics.evaluate_pose
        assert Pose3DMVFEXLightningModel.evaluate_heatmap.__wrapped__ is metrics.evaluate_heatmap
        with pytest.raises(RuntimeError, match="CUDA tensor"):          # no CPU fallback behind the patched method
            import torch
            Pose3DMVFEXLightningModel().evaluate_pose(torch.zeros(2, 16, 3), torch.zeros(2, 16, 3), "final")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
