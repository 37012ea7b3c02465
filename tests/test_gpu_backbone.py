"""GPU: the backbone engine (egr_backbone_*, csrc/engine_backbone.cu) vs the PyTorch ResNet18 + EfficientFPN
modules it replaces (mirror of pose_estimation/models/backbones/resnet.py:6-152, torchvision resnet18, eval mode), fp32,
same weights.  Random BatchNorm statistics / affine parameters exercise the BN folding.

Bounds (max|a-b| / max|ref| on the maps the hot path consumes): bf16 operands 2e-2, fp16 operands 4e-3."""
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = {"bf16": 2e-2, "fp16": 4e-3}


def build(precision, V=4, seed=0):
    from egorear_b200 import synth
    from egorear_b200.configs import heatmap_mvfex_cfg
    from egorear_b200.modules import EgoPoseFormerHeatmapMVFEX
    torch.manual_seed(seed)
    m = EgoPoseFormerHeatmapMVFEX(**heatmap_mvfex_cfg(V, "ego4view_rw"), precision=precision, build_backbone=True)
    synth.fill_state_dict(m)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for name, mod in m.named_modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.weight.copy_(1.0 + 0.2 * torch.randn(mod.weight.shape, generator=g))
                mod.bias.copy_(0.1 * torch.randn(mod.bias.shape, generator=g))
                mod.running_mean.copy_(0.1 * torch.randn(mod.running_mean.shape, generator=g))
                mod.running_var.copy_(1.0 + 0.3 * torch.rand(mod.running_var.shape, generator=g))
    return m.cuda().eval()


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
@pytest.mark.parametrize("B", [1, 3])
def test_backbone_engine_vs_torch(precision, B):
    m = build(precision)
    img = torch.randn((B, 4, 3, 256, 256), generator=torch.Generator().manual_seed(5)).cuda()
    with torch.no_grad():
        feat, feats = m.forward_heatmap_feat_estimation(img)               # PyTorch fp32: [B,4,128,64,64], [..., [B,4,512,8,8]]
        xh, bfb = m.forward_backbone_staged(img)
    assert xh.shape == (4, B, 64, 64, 128) and xh.is_contiguous() and bfb.shape == (B, 4, 512, 8, 8)
    got = xh.permute(1, 0, 4, 2, 3).float()
    e_feat, e_bfb = rel(got, feat), rel(bfb, feats[-1])
    print("backbone %s B=%d: rel err FPN map %.2e, stride-32 map %.2e" % (precision, B, e_feat, e_bfb))
    assert e_feat < TOL[precision] and e_bfb < TOL[precision]
    assert float(got.min()) >= 0.0                                         # the FPN ends in ReLU


def test_backbone_engine_stereo_and_reload():
    m = build("bf16", V=2)
    img = torch.randn((2, 2, 3, 256, 256), generator=torch.Generator().manual_seed(6)).cuda()
    with torch.no_grad():
        feat, feats = m.forward_heatmap_feat_estimation(img)
        xh, bfb = m.forward_backbone_staged(img)
        assert rel(xh.permute(1, 0, 4, 2, 3).float(), feat) < TOL["bf16"]
        sd = {k: v.clone() for k, v in m.state_dict().items()}
        sd["heatmap_estimator_stereo_front.encoder.neck.fpn_convs.0.0.bias"] += 1.0
        m.load_state_dict(sd, strict=True)                                 # derived (folded, 16-bit) weights are rebuilt
        xh2, _ = m.forward_backbone_staged(img)
    d = (xh2.float() - xh.float())
    assert float(d.mean()) > 0.9


def test_model_forward_with_backbone_engine_matches_torch_backbone():
    """EgoPoseFormerHeatmapMVFEX.forward(img) with backbone_impl="egr" (engine -> staged hand-over -> hot path) vs the same
    module driven by its PyTorch backbone: the refined heatmaps agree within the bf16 bounds of both halves"""
    m = build("bf16")
    img = torch.randn((2, 4, 3, 256, 256), generator=torch.Generator().manual_seed(7)).cuda()
    with torch.no_grad():
        m.backbone_impl = "torch"
        lh_t, lf_t = m(img)
        m.backbone_impl = "egr"
        lh_e, lf_e = m(img)
    assert lf_e[0] is None and lf_t[0] is not None
    assert rel(lh_e[0], lh_t[0]) < 3e-2 and rel(lh_e[1], lh_t[1]) < 3e-2


def test_stem_variants_agree():
    """the fused tensor-core stem (im2col rows built in shared memory) vs the im2col-buffer + plain-GEMM form, and the FPN
    fuse step as GEMM + coalesced kernel vs gathered in the GEMM epilogue: same weights, same 16-bit rounding points"""
    from egorear_b200 import engine
    m = build("bf16")
    img = torch.randn((2, 4, 3, 256, 256), generator=torch.Generator().manual_seed(8)).cuda()
    with torch.no_grad():
        base, bfb0 = m.forward_backbone_staged(img)
        try:
            engine.set_option("stem_fused", 0)
            a, bfb1 = m.forward_backbone_staged(img)
        finally:
            engine.set_option("stem_fused", 1)
        try:
            engine.set_option("fpn_epi", 1)
            b, _ = m.forward_backbone_staged(img)
        finally:
            engine.set_option("fpn_epi", 0)
    assert rel(a.float(), base.float()) < 5e-3 and rel(bfb1, bfb0) < 5e-3       # summation order inside the k-blocks only
    assert rel(b.float(), base.float()) < 1e-2                                   # one 16-bit rounding of the pre-activation more / less
