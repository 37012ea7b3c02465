"""CPU: pins the torch oracle (oracle/model_ref.py) to the golden vectors of the live reference modules."""
import numpy as np
import pytest
import torch

from conftest import orc_generate_target, rel_err
from egorear_b200 import calib, modules, synth
from oracle import model_ref, ref_import

from egorear_b200.configs import MVF_CFG, POSE3D_CFG as POSE_CFG


def build_mvfex(V=4, precision="fp32"):
    m = modules.EgoPoseFormerHeatmapMVFEX(num_views=V, image_size=[256, 256], num_heatmap=15, feat_down_stride=4,
                                          heatmap_threshold=0.5, encoder_cfg=None, mvf_cfg=MVF_CFG,
                                          camera_model="ego4view_syn", precision=precision, build_backbone=False)
    return synth.fill_state_dict(m).eval()


def build_pose3d(camera_model="ego4view_syn", precision="fp32"):
    m = modules.EgoPoseFormerPose3D(num_views=2 if "stereo" in camera_model else 4, image_size=[256, 256], use_pred_heatmap_init=True, camera_model=camera_model,
                                    precision=precision, **POSE_CFG)
    return synth.fill_state_dict(m).eval()


def anchor_heatmaps(oracle_lib, kp):
    """rendered Gaussians of the golden keypoints, joint 0 dropped like the datasets do ([1:])"""
    return torch.from_numpy(orc_generate_target(oracle_lib, kp)[:, :, 1:].copy())


@pytest.fixture(scope="module")
def mv4_oracle(golden, oracle_lib):
    feat, bfb = synth.synth_features(1, 4, seed=0)
    sd = build_mvfex(4).state_dict()
    hfa = anchor_heatmaps(oracle_lib, golden["models"]["kp"])
    with torch.no_grad():
        return model_ref.mvfex_hot_path(sd, feat, bfb, hfa), feat


def test_mirror_state_dict_keys():
    """the mirror exposes the parameter names/shapes the reference checkpoints carry"""
    sd = build_mvfex(4).state_dict()
    for k, shape in {"heatmap_refiner_front_left.transformer_layers.0.cross_attn.sampling_offsets.weight": (128, 256),
                     "heatmap_refiner_back_right.transformer_layers.0.cross_attn.spatial_shapes": (1, 2),
                     "heatmap_refiner_front_left.frame_feat_multi_view_pos_embed": (1, 4, 4096, 256),
                     "heatmap_refiner_front_right.joint_query_embed.weight": (15, 256),
                     "heatmap_refiner_back_left.conv_heatmap_layers.0.7.weight": (15, 128, 1, 1),
                     "heatmap_refiner_back_left.head_layers.0.head.3.weight": (128, 64, 1, 1),
                     "heatmap_refiner_back_left.frame_feat_refined_proj_layers.0.3.bias": (128,),
                     "heatmap_refiner_back_left.frame_feat_proj_layers.2.weight": (512, 256, 3, 3),
                     "heatmap_refiner_back_left.heatmap_proj.0.weight": (256, 4096),
                     "heatmap_refiner_back_left.fc_bfb.weight": (256, 512),
                     "heatmap_refiner_back_left.fc_query.0.weight": (256, 256),
                     "heatmap_refiner_back_left.post_norm.0.weight": (256,),
                     "conv_heatmap_layers_stereo_back.9.weight": (15, 128, 1, 1)}.items():
        assert tuple(sd[k].shape) == shape, k
    psd = build_pose3d().state_dict()
    for k, shape in {"feat_proj.weight": (128, 128, 1, 1), "layers.2.fuse_mlp.weight": (128, 512),
                     "query_gen_mlp.0.weight": (128, 4), "conv_frame_feat.7.weight": (128, 64, 3, 3),
                     "mlp_pred.0.0.weight": (2048, 32768), "mlp_pred.2.weight": (48, 128), "reg_mlp.1.2.weight": (3, 128),
                     "post_norm.2.bias": (128,), "layers.0.cross_attn.start_index": (1,)}.items():
        assert tuple(psd[k].shape) == shape, k


@pytest.mark.skipif(not ref_import.available(), reason="reference checkout not present (GPU box)")
def test_mirror_keys_equal_reference_keys():
    import copy
    cls = ref_import.import_estimators()
    cfg = ref_import.load_model_cfg("ego4view_syn_heatmap_mvfex-n1_jqa.yaml")
    ref = cls["EgoPoseFormerHeatmapMVFEX"](**copy.deepcopy(cfg))
    mine = modules.EgoPoseFormerHeatmapMVFEX(**copy.deepcopy(cfg), precision="fp32")
    rs, ms = ref.state_dict(), mine.state_dict()
    assert list(rs.keys()) == list(ms.keys())
    assert all(rs[k].shape == ms[k].shape and rs[k].dtype == ms[k].dtype for k in rs)
    mine.load_state_dict(rs, strict=True)
    c3 = ref_import.load_model_cfg("ego4view_syn_pose3d.yaml")
    ref3 = cls["EgoPoseFormerMVFEX"](**copy.deepcopy(c3))
    mine3 = modules.EgoPoseFormerMVFEX(**copy.deepcopy(c3), precision="fp32")
    assert list(ref3.state_dict().keys()) == list(mine3.state_dict().keys())
    mine3.load_state_dict(ref3.state_dict(), strict=True)


def test_oracle_mvfex_matches_golden(golden, mv4_oracle):
    (list_h, list_f, a2, av), _ = mv4_oracle
    g = golden["models"]
    assert np.array_equal(a2.numpy(), g["mv4_hfa_anchors"]) and np.array_equal(av.numpy(), g["mv4_hfa_valid"])
    assert rel_err(list_h[0][:, :, :, ::4, ::4].numpy(), g["mv4_hfa_hm_init"]) < 1e-5
    assert rel_err(list_h[1][:, :, :, ::4, ::4].numpy(), g["mv4_hfa_hm_ref"]) < 1e-5
    assert rel_err(list_f[1][:, :, ::8, ::8, ::8].numpy(), g["mv4_hfa_feat_ref"]) < 1e-5
    assert 0.3 < float(av.float().mean()) < 1.0      # anchors are a valid/invalid mix


def test_oracle_mvfex_self_anchor_and_stereo(golden, oracle_lib):
    g = golden["models"]
    feat, bfb = synth.synth_features(1, 4, seed=0)
    with torch.no_grad():
        lh, lf, a2, av = model_ref.mvfex_hot_path(build_mvfex(4).state_dict(), feat, bfb, None)
        assert np.array_equal(a2.numpy(), g["mv4_self_anchors"])
        assert rel_err(lh[1][:, :, :, ::4, ::4].numpy(), g["mv4_self_hm_ref"]) < 1e-5
        hfa = anchor_heatmaps(oracle_lib, g["kp"])
        lh2, lf2, _, _ = model_ref.mvfex_hot_path(build_mvfex(2).state_dict(), feat[:, :2], bfb[:, :2], hfa[:, :2])
        assert rel_err(lh2[1][:, :, :, ::4, ::4].numpy(), g["mv2_hm_ref"]) < 1e-5
        assert rel_err(lf2[1][:, :, ::8, ::8, ::8].numpy(), g["mv2_feat_ref"]) < 1e-5


def test_oracle_pose3d_matches_golden(golden, mv4_oracle):
    (list_h, list_f, _, _), feat = mv4_oracle
    for cam in ("ego4view_syn", "ego4view_rw"):
        ctm = synth.synth_coord_trans_mat(1, seed=5) if cam == "ego4view_rw" else None
        st = {}
        with torch.no_grad():
            preds = model_ref.pose3d_forward(build_pose3d(cam).state_dict(), feat, list_f[1], calib.load_calibration(None), cam,
                                             ctm, stages=st)
        got = torch.stack(preds).numpy()
        want = golden["models"]["pose_" + cam]
        mpjpe = np.linalg.norm(got - want, axis=-1).mean(axis=-1).max()
        assert mpjpe < 1e-4, (cam, mpjpe)          # cm; bound of the task is 0.01 cm
        if cam == "ego4view_syn":
            # the in-place quirk: the tensor left behind is p + (12, 0, 0)   
            d = (st["anchors_3d_after"] - preds[0]).numpy()
            assert np.allclose(d[..., 0], 12.0, atol=1e-4) and np.allclose(d[..., 1:], 0.0, atol=1e-4)
            assert 0.05 < float(st["anchors_valid"].float().mean()) < 0.95


STEREO_CAMS = (("ego4view_syn_stereo_front", slice(0, 2)), ("ego4view_syn_stereo_back", slice(2, 4)),
               ("ego4view_rw_stereo_front", slice(0, 2)))


def stereo_ctm(cam, sl):
    return synth.synth_coord_trans_mat(1, seed=5)[:, sl].contiguous() if "_rw" in cam else None


def stereo_case(golden, oracle_lib):
    """2-view stereo-front refinement (oracle) whose features the 2-view pose3d lifts, as in make_golden.golden_pose_stereo"""
    feat, bfb = synth.synth_features(1, 4, seed=0)
    hfa = anchor_heatmaps(oracle_lib, golden["models"]["kp"])
    with torch.no_grad():
        lh2, lf2, _, _ = model_ref.mvfex_hot_path(build_mvfex(2).state_dict(), feat[:, :2], bfb[:, :2], hfa[:, :2])
    return feat, lh2, lf2


def test_oracle_pose3d_stereo_matches_golden(golden, oracle_lib):
    """configs/ego4view_syn_pose3d_stereo_front.yaml (and the stereo-back camera pair): 2 views, mlp_pred.0.0 16384 -> 1024"""
    feat, _, lf2 = stereo_case(golden, oracle_lib)
    g = golden["pose3d_stereo"]
    for cam, sl in STEREO_CAMS:
        m = build_pose3d(cam)
        assert tuple(m.state_dict()["mlp_pred.0.0.weight"].shape) == tuple(g["shape_mlp0_" + cam])
        with torch.no_grad():
            preds = model_ref.pose3d_forward(m.state_dict(), feat[:, sl], lf2[1], calib.load_calibration(None), cam,
                                             stereo_ctm(cam, sl))
        mpjpe = np.linalg.norm(torch.stack(preds).numpy() - g["pose_" + cam], axis=-1).mean(axis=-1).max()
        assert mpjpe < 1e-4, (cam, mpjpe)


def test_msda_restatement_vs_grid_sample():
    """literal per-corner restatement (oracle) == the grid_sample restatement used for the live reference"""
    from oracle.ref_import import _MSDAGridSample
    g = torch.Generator().manual_seed(3)
    B, H, W, nh, hd, Q, P = 2, 64, 64, 4, 32, 16, 16
    value = torch.randn((B, H * W, nh, hd), generator=g)
    loc = torch.rand((B, Q, nh, 1, P, 2), generator=g) * 1.3 - 0.15       # some points outside the map
    aw = torch.softmax(torch.randn((B, Q, nh, P), generator=g), -1).view(B, Q, nh, 1, P)
    a = model_ref.ms_deform_attn(value, H, W, loc, aw)
    b = _MSDAGridSample.apply(value, torch.tensor([[H, W]]), torch.tensor([0]), loc, aw, 32)
    assert rel_err(a.numpy(), b.numpy()) < 1e-5


def test_msda_restatement_vs_transformers_deformable_detr():
    """second, independent opinion on the un-vendored mmcv op: Hugging Face's pure-PyTorch
    `MultiScaleDeformableAttention` of Deformable DETR (same MSDA definition, written by other people) on the same inputs,
    single level as in EgoRear, at both map sizes the hot path samples (64x64, 32x32), with points outside the map"""
    mod = pytest.importorskip("transformers.models.deformable_detr.modeling_deformable_detr")
    msda = mod.MultiScaleDeformableAttention()
    g = torch.Generator().manual_seed(5)
    B, nh, hd, Q, P = 2, 4, 32, 15, 16
    for shapes in ([(64, 64)], [(32, 32)]):
        (H, W), = shapes
        value = torch.randn((B, H * W, nh, hd), generator=g)
        loc = torch.rand((B, Q, nh, 1, P, 2), generator=g) * 1.3 - 0.15       # some points outside the map
        aw = torch.softmax(torch.randn((B, Q, nh, P), generator=g), -1).view(B, Q, nh, 1, P)
        ours = model_ref.ms_deform_attn(value, H, W, loc, aw)
        theirs = msda(value, shapes, torch.tensor(shapes), torch.tensor([0]), loc, aw, 64)
        assert theirs.shape == ours.shape
        assert rel_err(ours.numpy(), theirs.numpy()) < 1e-5
