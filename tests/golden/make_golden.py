"""Generates the committed golden fixtures from the LIVE reference (run in the build container only):

    python tests/golden/make_golden.py

Imports /root/reference through oracle/ref_import.py (mmcv/timm/natsort stubbed, see there), drives the
reference's own functions/modules with the deterministic inputs and name-seeded weights of
egorear_b200/synth.py, and stores small (sub-sampled) outputs as .npz next to this script.  The GPU box has
no reference checkout: tests there compare against these files and against oracle/ (which
tests/test_oracle*.py pin to these same files).
"""
import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import                     # noqa: E402
from egorear_b200 import synth                    # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

EDGE_JOINTS = np.array([[-100.0, -100.0], [-30.0, 400.0], [871.9, 871.9], [900.0, 900.0], [915.0, 400.0], [0.0, 0.0],
                        [-20.0, -20.0], [436.0, 436.0], [6.8124, 6.8126], [865.1875, 20.4375], [-47.6875, 300.0],
                        [300.0, 912.8], [13.625 * 63.5, 13.625 * 0.5], [-6.8125, -6.8124], [858.374, 858.376],
                        [926.5, 926.4]], dtype=np.float64)


def sparse(hm):
    """[N,J,H,W] -> (flat indices int64, values float32) of the non-zeros"""
    flat = hm.reshape(-1)
    idx = np.flatnonzero(flat)
    return idx.astype(np.int64), flat[idx].astype(np.float32)


def golden_generate_target(fn):
    kp = synth.synth_keypoints(24, 4, 16, seed=7)                    # [24,4,16,2]
    joints = np.concatenate([kp.reshape(-1, 16, 2), EDGE_JOINTS[None]], axis=0)   # [97,16,2]
    out = np.stack([fn(j, image_size=872, heatmap_size=64, num_joints=16, sigma=1.0) for j in joints])
    idx, val = sparse(out)
    # second geometry: 15 joints, sigma=2, 48x48 map of a 640 px image (exercises the general path)
    j2 = synth.synth_keypoints(6, 1, 15, seed=8, lo=-40, hi=680).reshape(-1, 15, 2)
    out2 = np.stack([fn(j, image_size=640, heatmap_size=48, num_joints=15, sigma=2) for j in j2])
    idx2, val2 = sparse(out2)
    np.savez_compressed(os.path.join(HERE, "generate_target.npz"), joints=joints, idx=idx, val=val, shape=np.array(out.shape),
                        joints2=j2, idx2=idx2, val2=val2, shape2=np.array(out2.shape))
    print("generate_target:", out.shape, len(idx), "non-zeros;", out2.shape, len(idx2))


def decode_inputs():
    """Deterministic heatmaps incl. ties, a NaN, all-equal maps and threshold edge values."""
    g = torch.Generator().manual_seed(11)
    hm = torch.randn((6, 15, 64, 64), generator=g)
    hm[0, 0] = 0.25                                   # all equal -> index 0
    hm[0, 1, 10, 20] = 9.0; hm[0, 1, 30, 5] = 9.0     # tie -> first (row 10)
    hm[0, 2, 63, 63] = 50.0                           # last element
    hm[0, 3, 7, 9] = float("nan")                     # NaN propagates
    hm[0, 4] = -torch.rand((64, 64), generator=g)     # all negative
    hm[0, 5, 0, 0] = 0.5; hm[0, 5, 0, 1:] = 0.4999; hm[0, 5, 1:] = 0.1   # max exactly at threshold
    hm[1, 0] = float("-inf")
    hm[1, 1, 40, 40] = float("inf")
    return hm


def golden_decode(fn):
    hm = decode_inputs()
    out = {}
    for tag, (thr, norm) in {"model": (0.5, True), "eval": (1.0, False)}.items():
        p, m, v = fn(hm.clone(), threshold=thr, normalize=norm)
        out["preds_" + tag], out["maxvals_" + tag], out["valid_" + tag] = p.numpy(), m.numpy(), v.numpy()
    np.savez_compressed(os.path.join(HERE, "get_max_preds.npz"), **out)
    print("get_max_preds:", {k: v.shape for k, v in out.items()})


def soft_inputs():
    """finite heatmaps: random logits at several temperatures, rendered-Gaussian-like peaks, a constant map"""
    g = torch.Generator().manual_seed(12)
    hm = torch.randn((4, 15, 64, 64), generator=g)
    hm[1] *= 8.0                                      # peaky softmax
    hm[2] *= 0.05                                     # nearly uniform
    hm[3, 0] = 0.25                                   # constant -> centre of mass (31.5, 31.5)
    hm[3, 1] = 0.0; hm[3, 1, 5, 60] = 30.0            # one dominant cell
    hm[3, 2] = -50.0; hm[3, 2, 63, 0] = 40.0; hm[3, 2, 0, 63] = 40.0   # two equal peaks
    return hm


def golden_soft(fn):
    hm = soft_inputs()
    out = {}
    for tag, norm in (("raw", False), ("norm", True)):
        p, m = fn(hm.clone(), normalize=norm)
        out["preds_" + tag], out["maxvals_" + tag] = p.numpy(), m.numpy()
    np.savez_compressed(os.path.join(HERE, "get_max_preds_soft.npz"), **out)
    print("get_max_preds_soft:", {k: v.shape for k, v in out.items()})


def golden_models(cls, gen_target):
    B = 1
    feat, bfb = synth.synth_features(B, 4, seed=0)
    kp = synth.synth_keypoints(B, 4, 16, seed=3)
    hfa = torch.from_numpy(np.stack([[gen_target(kp[b, v], 872, 64, 16, 1.0)[1:] for v in range(4)] for b in range(B)]))
    out = {"kp": kp}
    # ---- 4-view mvfex (configs[1]) ----
    cfg = ref_import.load_model_cfg("ego4view_syn_heatmap_mvfex-n1_jqa.yaml")
    m = cls["EgoPoseFormerHeatmapMVFEX"](**copy.deepcopy(cfg)).eval()
    synth.fill_state_dict(m)
    m.forward_heatmap_feat_estimation = lambda img: (feat, [None, None, None, bfb])   # backbone is out of scope
    with torch.no_grad():
        for tag, h in (("hfa", hfa), ("self", None)):
            hp, ff = m(torch.zeros(B, 4, 3, 256, 256), heatmap_for_anchor=h)
            a2, mv, av = m.get_anchors_2d_from_hm(h if h is not None else hp[0])
            out["mv4_%s_hm_init" % tag] = hp[0][:, :, :, ::4, ::4].numpy()
            out["mv4_%s_hm_ref" % tag] = hp[1][:, :, :, ::4, ::4].numpy()
            out["mv4_%s_feat_ref" % tag] = ff[1][:, :, ::8, ::8, ::8].numpy()
            out["mv4_%s_anchors" % tag] = a2.numpy()
            out["mv4_%s_valid" % tag] = av.numpy()
            if tag == "hfa":          # the pose3d fixtures lift the Gaussian-anchored refinement
                feat_ref_full, hm_ref_full = ff[1], hp[1]
    # ---- 2-view stereo-front mvfex ----
    cfg2 = ref_import.load_model_cfg("ego4view_syn_heatmap_mvfex-n1_jqa_stereo_front.yaml")
    m2 = cls["EgoPoseFormerHeatmapMVFEX"](**copy.deepcopy(cfg2)).eval()
    synth.fill_state_dict(m2)
    m2.forward_heatmap_feat_estimation = lambda img: (feat[:, :2], [None, None, None, bfb[:, :2]])
    with torch.no_grad():
        hp2, ff2 = m2(torch.zeros(B, 2, 3, 256, 256), heatmap_for_anchor=hfa[:, :2])
    out["mv2_hm_init"] = hp2[0][:, :, :, ::4, ::4].numpy()
    out["mv2_hm_ref"] = hp2[1][:, :, :, ::4, ::4].numpy()
    out["mv2_feat_ref"] = ff2[1][:, :, ::8, ::8, ::8].numpy()
    # ---- pose3d (configs[2]) syn + rw ----
    for cam, cfgname in (("ego4view_syn", "ego4view_syn_pose3d.yaml"), ("ego4view_rw", "ego4view_rw_pose3d.yaml")):
        c3 = ref_import.load_model_cfg(cfgname)
        pc = copy.deepcopy(c3["pose3d_cfg"])
        pc.update(dict(num_views=4, image_size=[256, 256], use_pred_heatmap_init=True, camera_model=cam))
        p3 = cls["EgoPoseFormerPose3D"](**pc).eval()
        synth.fill_state_dict(p3)
        ctm = synth.synth_coord_trans_mat(B, seed=5) if cam == "ego4view_rw" else None
        with torch.no_grad():
            preds = p3(feat, feat_ref_full, hm_ref_full, ctm)
        out["pose_%s" % cam] = torch.stack(preds).numpy()
    np.savez_compressed(os.path.join(HERE, "models.npz"), **out)
    print("models:", {k: v.shape for k, v in out.items()})


def golden_pose_stereo(cls, gen_target):
    """2-view configs (ego4view_syn_pose3d_stereo_front.yaml / ..._stereo_back cameras): the stereo-front mvfex refinement
    lifted by a 2-view EgoPoseFormerPose3D (mlp_pred.0.0 is then Linear(16384 -> 1024))."""
    B = 1
    feat, bfb = synth.synth_features(B, 4, seed=0)
    kp = synth.synth_keypoints(B, 4, 16, seed=3)
    hfa = torch.from_numpy(np.stack([[gen_target(kp[b, v], 872, 64, 16, 1.0)[1:] for v in range(4)] for b in range(B)]))
    cfg2 = ref_import.load_model_cfg("ego4view_syn_heatmap_mvfex-n1_jqa_stereo_front.yaml")
    m2 = cls["EgoPoseFormerHeatmapMVFEX"](**copy.deepcopy(cfg2)).eval()
    synth.fill_state_dict(m2)
    m2.forward_heatmap_feat_estimation = lambda img: (feat[:, :2], [None, None, None, bfb[:, :2]])
    with torch.no_grad():
        hp2, ff2 = m2(torch.zeros(B, 2, 3, 256, 256), heatmap_for_anchor=hfa[:, :2])
    out = {}
    c3 = ref_import.load_model_cfg("ego4view_syn_pose3d_stereo_front.yaml")
    for cam, sl in (("ego4view_syn_stereo_front", slice(0, 2)), ("ego4view_syn_stereo_back", slice(2, 4)),
                    ("ego4view_rw_stereo_front", slice(0, 2))):
        pc = copy.deepcopy(c3["pose3d_cfg"])
        pc.update(dict(num_views=2, image_size=[256, 256], use_pred_heatmap_init=True, camera_model=cam))
        p3 = cls["EgoPoseFormerPose3D"](**pc).eval()
        synth.fill_state_dict(p3)
        ctm = synth.synth_coord_trans_mat(B, seed=5)[:, sl].contiguous() if "_rw" in cam else None
        with torch.no_grad():
            preds = p3(feat[:, sl], ff2[1], hp2[1], ctm)
        out["pose_%s" % cam] = torch.stack(preds).numpy()
        out["shape_mlp0_%s" % cam] = np.array(p3.state_dict()["mlp_pred.0.0.weight"].shape)
    np.savez_compressed(os.path.join(HERE, "pose3d_stereo.npz"), **out)
    print("pose3d_stereo:", {k: v.shape for k, v in out.items()}, out["shape_mlp0_ego4view_syn_stereo_front"])


INTEGRATE_CASES = (("sm100", True, 100.0), ("sm1", True, 1.0), ("relu", False, 100.0))
INTEGRATE_MAPS = ((0, 0), (1, 5), (2, 7), (3, 1), (3, 2), (3, 3))


def integrate_inputs():
    hm = soft_inputs()
    hm[3, 3] = -1.0                                   # relu variant: zero mass -> NaN coordinates
    return hm


def golden_integrate():
    """utils/util.py:80-109 (open3d, imported at the top of that module, is stubbed: not installed, not used here)"""
    import types
    ref_import.install()
    sys.modules.setdefault("open3d", types.ModuleType("open3d"))
    from pose_estimation.utils import util
    hm = integrate_inputs()
    out = {}
    for tag, sm, mult in INTEGRATE_CASES:
        c, p = util.integrate_tensor_2d(hm.clone(), softmax=sm, multiplier=mult)
        out["coords_" + tag] = c.numpy()
        out["maps_" + tag] = np.stack([p[b, j].numpy() for b, j in INTEGRATE_MAPS])
    np.savez_compressed(os.path.join(HERE, "integrate_tensor_2d.npz"), **out)
    print("integrate_tensor_2d:", {k: v.shape for k, v in out.items()})


PREPROCESS_CASES = ((872, 872, 256, 256), (480, 640, 256, 256), (100, 75, 256, 256), (256, 256, 256, 256), (301, 257, 64, 96))


def golden_preprocess():
    """datasets/*: transform(Image.open(p).convert("RGB").resize([256, 256], Image.BICUBIC)) with the real PIL +
    torchvision; inputs are regenerated from the seed by the tests, outputs are stored as a crop + a SHA-256 of the whole."""
    import hashlib
    from PIL import Image
    from torchvision import transforms
    tf = transforms.Compose([transforms.ToTensor(), transforms.Normalize((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))])
    out = {}
    for i, (H, W, oh, ow) in enumerate(PREPROCESS_CASES):
        img = synth.synth_images(1, H, W, seed=i)[0]
        pil = Image.fromarray(img).convert("RGB").resize([ow, oh], Image.BICUBIC)
        u8 = np.asarray(pil)
        f = tf(pil).float().numpy()
        out["u8_crop_%d" % i] = u8[: 48, : 48].copy()
        out["f32_crop_%d" % i] = f[:, : 48, : 48].copy()
        out["u8_sha_%d" % i] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(u8).tobytes()).digest(), np.uint8)
        out["f32_sha_%d" % i] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(f).tobytes()).digest(), np.uint8)
    np.savez_compressed(os.path.join(HERE, "preprocess.npz"), **out)
    print("preprocess:", {k: v.shape for k, v in out.items() if "crop" in k})


def golden_eval_metrics(w):
    """eval-time metrics: the wrappers' own methods, called unbound on a stand-in `self` that
    carries exactly the attributes they read (criteria, cm2mm, num_heatmap, get_anchors_2d_from_hm)."""
    import types
    H, P = w["PoseHeatmapMVFEXLightningModel"], w["Pose3DMVFEXLightningModel"]
    out = {}
    pred, gt = synth.synth_eval_poses(64, 16, seed=0)
    self_p = types.SimpleNamespace(cm2mm=10)
    m = P.evaluate_pose(self_p, torch.from_numpy(pred), torch.from_numpy(gt), "final")
    for k, v in m.items():
        out["pose_" + k] = np.asarray(v)
    out["pose_s1_hat"] = w["pose_metric"].batch_compute_similarity_transform_numpy(torch.from_numpy(pred), torch.from_numpy(gt)).numpy()
    ph, gh = synth.synth_eval_heatmaps(6, 4, 15, seed=0)
    self_h = types.SimpleNamespace(criteria=torch.nn.MSELoss(reduction="mean"), num_heatmap=15)
    self_h.get_anchors_2d_from_hm = lambda hm: H.get_anchors_2d_from_hm(self_h, hm)
    for tag, sl in (("front", slice(0, 2)), ("back", slice(2, 4)), ("all", slice(0, 4))):
        m = H.evaluate(self_h, torch.from_numpy(ph)[:, sl], torch.from_numpy(gh)[:, sl], tag)
        for k, v in m.items():
            out["hm_" + k] = v.numpy()
    m = P.evaluate_heatmap(self_p, torch.from_numpy(ph), torch.from_numpy(gh), "p3d")
    for k, v in m.items():
        out["hm_" + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "eval_metrics.npz"), **out)
    print("eval_metrics:", {k: (v.shape, v.dtype) for k, v in out.items()})


if __name__ == "__main__":
    assert ref_import.available(), "needs the reference checkout"
    if "--metrics-only" in sys.argv:            # regenerate eval_metrics.npz alone
        golden_eval_metrics(ref_import.import_wrappers())
        sys.exit(0)
    if "--integrate-only" in sys.argv:
        golden_integrate()
        sys.exit(0)
    if "--preprocess-only" in sys.argv:
        golden_preprocess()
        sys.exit(0)
    if "--pose-stereo-only" in sys.argv:
        golden_pose_stereo(ref_import.import_estimators(), ref_import.import_functions()["generate_target"])
        sys.exit(0)
    fns = ref_import.import_functions()
    golden_generate_target(fns["generate_target"])
    golden_decode(fns["get_max_preds"])
    golden_soft(fns["get_max_preds_soft_pytorch"])
    golden_eval_metrics(ref_import.import_wrappers())
    golden_integrate()
    golden_preprocess()
    golden_models(ref_import.import_estimators(), fns["generate_target"])
    golden_pose_stereo(ref_import.import_estimators(), fns["generate_target"])
