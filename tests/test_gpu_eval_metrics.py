"""GPU: eval-time metrics through the C-ABI vs the golden vectors of the live reference and the oracle."""
import numpy as np
import pytest
import torch

from egorear_b200 import synth
from oracle import metrics_ref as mr

pytestmark = pytest.mark.gpu

POSE_KEYS = ("mpjpe", "pa_mpjpe", "pck_3d", "auc_3d")


@pytest.fixture(scope="module")
def metrics():
    from egorear_b200 import metrics as m
    return m


def test_evaluate_pose_golden(metrics, golden):
    g = golden["eval_metrics"]
    pred, gt = synth.synth_eval_poses(64, 16, seed=0)
    p, t = torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda()
    m = metrics.evaluate_pose(p, t, "final")
    assert list(m) == ["final_mpjpe", "final_pa_mpjpe", "final_pck_3d", "final_auc_3d"]
    assert all(isinstance(v, np.ndarray) and v.shape == (64,) for v in m.values())
    assert m["final_mpjpe"].dtype == np.float32 and m["final_pa_mpjpe"].dtype == np.float64
    assert np.allclose(m["final_mpjpe"], g["pose_final_mpjpe"], rtol=1e-6, atol=1e-5)
    # float32 SVD on both sides (LAPACK per sample there, Jacobi here) -> float32-level agreement (mm)
    assert np.allclose(m["final_pa_mpjpe"], g["pose_final_pa_mpjpe"], rtol=1e-5, atol=2e-3)
    assert np.array_equal(m["final_pck_3d"], g["pose_final_pck_3d"])
    assert np.allclose(m["final_auc_3d"], g["pose_final_auc_3d"], rtol=1e-6)
    s = metrics.batch_compute_similarity_transform_numpy(p, t)
    assert s.dtype == torch.float64 and s.is_cuda
    assert np.allclose(s.cpu().numpy(), g["pose_s1_hat"], rtol=0, atol=2e-4)
    # the stand-alone functions of utils/loss.py
    s10 = torch.tensor(10.0, device="cuda")
    assert np.allclose(metrics.compute_mpjpe_batch(p, t).cpu().numpy() * 10, g["pose_final_mpjpe"], rtol=1e-6, atol=1e-5)
    assert np.array_equal(metrics.compute_pck_3d_batch(p * s10, t * s10).cpu().numpy() * 100, g["pose_final_pck_3d"])
    assert np.allclose(metrics.compute_auc_3d_batch(p * s10, t * s10).cpu().numpy() * 100, g["pose_final_auc_3d"], rtol=1e-6)


@pytest.mark.parametrize("B,J,seed", [(1, 16, 1), (63, 16, 2), (65, 15, 3), (1000, 5, 4), (257, 32, 5)])
def test_evaluate_pose_oracle_ragged(metrics, B, J, seed):
    pred, gt = synth.synth_eval_poses(B, J, seed=seed)
    m = metrics.evaluate_pose(torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda(), "x")
    o = mr.evaluate_pose(pred, gt)
    for k in POSE_KEYS:
        assert np.allclose(m["x_" + k], o[k], rtol=1e-5, atol=5e-3), k


def test_evaluate_pose_properties_large(metrics):
    """size-independent properties at 2^18 samples: PA-MPJPE is invariant to a similarity applied to the prediction,
    the aligned squared residual never exceeds the unaligned one; identical poses give 0 / 100 / 100; results do not depend on the
    position of a sample in the batch."""
    B = 1 << 18
    pred, gt = synth.synth_eval_poses(B, 16, seed=7)
    p, t = torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda()
    m = metrics.evaluate_pose(p, t, "a")
    assert np.isfinite(m["a_pa_mpjpe"]).all()
    s1 = metrics.batch_compute_similarity_transform_numpy(p, t)
    sq_after = ((s1 - t.double()) ** 2).sum((1, 2))
    sq_before = ((p.double() - t.double()) ** 2).sum((1, 2))
    assert bool((sq_after <= sq_before * (1 + 1e-9) + 1e-9).all())         # the alignment minimises the squared residual
    del s1, sq_after, sq_before
    c, s = np.cos(0.7), np.sin(0.7)
    R = torch.tensor([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]], device="cuda", dtype=torch.float32)
    m2 = metrics.evaluate_pose((p @ R.T) * 1.7 + 5.0, t, "b")
    assert np.allclose(m2["b_pa_mpjpe"][4:], m["a_pa_mpjpe"][4:], rtol=2e-4, atol=2e-2)
    m3 = metrics.evaluate_pose(t, t, "c")
    assert (m3["c_mpjpe"] == 0).all() and (m3["c_pa_mpjpe"] < 1e-3).all() and (m3["c_pck_3d"] == 100).all() and (m3["c_auc_3d"] == 100).all()
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(0)).cuda()
    m4 = metrics.evaluate_pose(p[perm], t[perm], "d")
    pn = perm.cpu().numpy()
    for k in POSE_KEYS:
        assert np.array_equal(m4["d_" + k], m["a_" + k][pn]), k


def test_evaluate_heatmap_golden(metrics, golden):
    g = golden["eval_metrics"]
    pred, gt = synth.synth_eval_heatmaps(6, 4, 15, seed=0)
    p, t = torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda()
    # the wrappers pass view slices (heatmap_mvf_ex.py:162-170): no copy is made for them
    for tag, sl in (("front", slice(0, 2)), ("back", slice(2, 4)), ("all", slice(0, 4))):
        m = metrics.evaluate(p[:, sl], t[:, sl], tag)
        assert list(m) == [tag + "_l1_error_heatmap", tag + "_pos_l1_error_heatmap", tag + "_mse_heatmap", tag + "_mse_pts2d"]
        assert not m[tag + "_l1_error_heatmap"].is_cuda and m[tag + "_l1_error_heatmap"].shape == (6,)
        assert m[tag + "_mse_heatmap"].is_cuda and m[tag + "_mse_heatmap"].ndim == 0
        for k, v in m.items():
            assert np.allclose(v.cpu().numpy(), g["hm_" + k], rtol=5e-6), k
    m = metrics.evaluate_heatmap(p, t, "p3d")
    assert list(m) == ["p3d_l1_error_heatmap", "p3d_pos_l1_error_heatmap"]
    for k, v in m.items():
        assert np.allclose(v.numpy(), g["hm_" + k], rtol=5e-6), k


@pytest.mark.parametrize("B,V,C,hs", [(1, 1, 1, 64), (3, 2, 15, 64), (5, 4, 16, 32), (2, 3, 7, 6)])
def test_evaluate_heatmap_oracle_shapes(metrics, B, V, C, hs):
    pred, gt = synth.synth_eval_heatmaps(B, V, C, hs, seed=B)
    pred[0, 0, 0, 1, 1] = 5.0; pred[0, 0, 0, hs - 1, hs - 1] = 5.0       # tie: the first maximum wins
    m = metrics.evaluate(torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda(), "x")
    o = mr.evaluate_heatmap(pred, gt)
    for k, v in o.items():
        assert np.allclose(m["x_" + k].cpu().numpy(), v, rtol=5e-6, atol=1e-7), k


def test_evaluate_heatmap_properties_large(metrics):
    """2048 frames x 4 views x 15 joints (2 x 2 GB): the per-frame errors of a batch equal those of its halves, the
    global MSE is the frame-weighted mean of the halves', identical inputs give zeros, the result is run-to-run
    bit-identical (fixed reduction order)."""
    from egorear_b200 import ops
    B = 2048
    kp = torch.from_numpy(synth.synth_keypoints(B, 4, 15, seed=3)).cuda()
    gt = ops.generate_target_batch(kp)                                    # [B,4,15,64,64]
    g = torch.Generator(device="cuda").manual_seed(5)
    pred = gt * 0.9 + torch.randn(gt.shape, device="cuda", generator=g) * 0.03
    m = metrics.evaluate(pred, gt, "a")
    m_again = metrics.evaluate(pred, gt, "a")
    for k in m:
        assert torch.equal(m[k].cpu(), m_again[k].cpu()), k
    h = B // 2
    lo, hi = metrics.evaluate(pred[:h], gt[:h], "a"), metrics.evaluate(pred[h:], gt[h:], "a")
    for k in ("a_l1_error_heatmap", "a_pos_l1_error_heatmap"):
        assert torch.equal(torch.cat([lo[k], hi[k]]), m[k]), k
    for k in ("a_mse_heatmap", "a_mse_pts2d"):
        assert abs(float(m[k]) - 0.5 * (float(lo[k]) + float(hi[k]))) <= 1e-6 * abs(float(m[k])), k
    # against torch on the same device (fp32 sums in another order)
    want_l1 = (pred - gt).abs().reshape(B, -1).sum(1)
    assert torch.allclose(m["a_l1_error_heatmap"], want_l1.cpu(), rtol=2e-5)
    assert abs(float(m["a_mse_heatmap"]) - float(torch.nn.functional.mse_loss(pred, gt))) <= 2e-5 * float(m["a_mse_heatmap"])
    z = metrics.evaluate(gt, gt, "z")
    assert all(float(v.abs().max()) == 0.0 for v in z.values())


def test_eval_argument_validation(metrics):
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        metrics.evaluate_pose(torch.zeros(2, 16, 3), torch.zeros(2, 16, 3), "x")
    with pytest.raises(RuntimeError, match=r"\[B, J, 3\]"):
        metrics.evaluate_pose(torch.zeros(2, 16, 2).cuda(), torch.zeros(2, 16, 2).cuda(), "x")
    with pytest.raises(RuntimeError, match="at most 32 joints"):
        metrics.evaluate_pose(torch.zeros(2, 40, 3).cuda(), torch.zeros(2, 40, 3).cuda(), "x")
    with pytest.raises(RuntimeError, match=r"\[B, V, C, H, W\]"):
        metrics.evaluate(torch.zeros(2, 15, 64, 64).cuda(), torch.zeros(2, 15, 64, 64).cuda(), "x")
    # empty batch: nothing launched, empty results
    m = metrics.evaluate_pose(torch.zeros(0, 16, 3).cuda(), torch.zeros(0, 16, 3).cuda(), "e")
    assert all(v.shape == (0,) for v in m.values())
