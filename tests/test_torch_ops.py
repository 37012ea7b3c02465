"""CPU: the egr::* torch.library operators (egorear_b200/torch_ops.py) — schemas, fake implementations, and that the
mirror modules' forwards trace into ONE graph of opaque egr::* nodes (no graph break), which is what `torch.compile`
of `model.network` needs (reference: run.py:7-9, `compile: True` in every shipped config).  No kernel runs here: the
inputs are FakeTensors on a fake CUDA device."""
import copy

import pytest
import torch
from torch._subclasses.fake_tensor import FakeTensorMode

from test_oracle_model import build_mvfex, build_pose3d


def _fake(*shape, dtype=torch.float32):
    return torch.empty(shape, dtype=dtype, device="cuda")


def test_ops_registered_with_schemas():
    from egorear_b200 import torch_ops  # noqa: F401
    for name in ("decode_argmax", "decode_soft_argmax", "integrate_tensor_2d", "generate_target", "msda_forward",
                 "heatmap_head_1x1", "pack_joints", "mvfex_forward", "mvfex_refiner_forward", "pose3d_forward"):
        op = getattr(torch.ops.egr, name).default
        assert "egr::" + name in str(op._schema)


def test_fake_impls_give_reference_shapes():
    from egorear_b200 import ops
    with FakeTensorMode():
        hm = _fake(8, 15, 64, 64)
        p, m, v = ops.get_max_preds(hm, 0.5, True)
        assert p.shape == (8, 15, 2) and m.shape == (8, 15) and v.shape == (8, 15) and v.dtype == torch.bool
        p1, m1, v1 = ops.get_max_preds(_fake(1, 15, 64, 64), 0.5)          # the squeeze quirk of utils/loss.py:142
        assert p1.shape == (1, 15, 2) and m1.shape == (15,) and v1.shape == (15,)
        ps, ms = ops.get_max_preds_soft_pytorch(hm)
        assert ps.shape == (8, 15, 2) and ms.shape == (8, 15, 1)
        c, pr = ops.integrate_tensor_2d(hm)
        assert c.shape == (8, 15, 2) and pr.shape == hm.shape
        g = ops.generate_target_batch(_fake(5, 4, 16, 2, dtype=torch.float64))
        assert g.shape == (5, 4, 16, 64, 64) and g.dtype == torch.float32
        val = _fake(2, 4096, 4, 64)
        o = ops.ms_deform_attn(val, [[64, 64]], None, _fake(2, 15, 4, 1, 16, 2), _fake(2, 15, 4, 1, 16))
        assert o.shape == (2, 15, 256)
        h = ops.heatmap_head_1x1(_fake(6, 128, 64, 64), _fake(15, 128, 1, 1), _fake(15))
        assert h.shape == (6, 15, 64, 64)
        assert ops.pack_joints(_fake(3, 4, 15, 2), _fake(3, 16, 3)).shape == (3, 168)


def _targets(gm):
    return [str(n.target) for n in gm.graph.nodes if n.op == "call_function" and "egr" in str(n.target)]


def test_hot_path_traces_as_one_graph_of_egr_nodes():
    """the chained model (heatmap estimator -> decode -> pose3d -> pack) is ONE fx graph: no graph break inside the hot path"""
    from egorear_b200 import ops
    m = build_mvfex(4, "bf16")
    p = build_pose3d("ego4view_syn", "bf16")

    def fn(feat, bfb):
        lh, lf = m.forward_from_feats(feat, bfb, want_feat_refined=False)
        B, V, J, H, W = lh[-1].shape
        pts, _, _ = ops.get_max_preds(lh[-1].view(B * V, J, H, W), 0.5, False)
        poses = p(lf[0], lf[-1], lh[-1], None, staged=m)
        return ops.pack_joints(pts.view(B, -1), poses[-1]), lh[-1]

    with FakeTensorMode():
        feat, bfb = _fake(2, 4, 128, 64, 64), _fake(2, 4, 512, 8, 8)
    ex = torch._dynamo.export(fn)(feat, bfb)           # raises on any graph break
    t = _targets(ex.graph_module)
    assert t == ["egr.mvfex_forward", "egr.decode_argmax", "egr.pose3d_forward", "egr.pack_joints"], t
    # the engine keys are compile-time constants of the two modules
    node = next(n for n in ex.graph_module.graph.nodes if "pose3d_forward" in str(n.target))
    assert node.args[0] == p._egr_key and node.args[1] == m._egr_key


def test_standalone_modules_trace():
    from egorear_b200 import modules, synth
    from egorear_b200.configs import MVF_CFG
    r = modules.HeatmapMVF(image_size=[256, 256], feat_down_stride=4, detach_heatmap_feat=False, heatmap_threshold=0.5,
                           num_views=4, num_heatmap=15, precision="fp32", **MVF_CFG).eval()
    with FakeTensorMode():
        args = (_fake(2, 15, 64, 64), _fake(2, 128, 64, 64), _fake(2, 4, 128, 64, 64), _fake(2, 4, 15, 2),
                _fake(2, 4, 15, dtype=torch.bool), _fake(2, 512, 8, 8), _fake(2, 4, 512, 8, 8))
        val, loc, aw = _fake(2, 4096, 4, 64), _fake(2, 15, 4, 1, 16, 2), _fake(2, 15, 4, 1, 16)
    ex = torch._dynamo.export(r)(*args)
    assert _targets(ex.graph_module) == ["egr.mvfex_refiner_forward"]
    ex = torch._dynamo.export(lambda v, l, w: ops_mod().ms_deform_attn(v, [[64, 64]], None, l, w))(val, loc, aw)
    assert _targets(ex.graph_module) == ["egr.msda_forward"]


def ops_mod():
    from egorear_b200 import ops
    return ops


def test_copies_get_their_own_key():
    m = build_mvfex(4, "bf16")
    m2 = copy.deepcopy(m)
    assert m2._egr_key != m._egr_key and m2._engine is None
    from egorear_b200 import torch_ops
    assert torch_ops._module(m2._egr_key) is m2 and torch_ops._module(m._egr_key) is m
    key = m2._egr_key
    del m2
    import gc
    gc.collect()
    with pytest.raises(RuntimeError, match="not alive"):
        torch_ops._module(key)
