"""GPU: training-side operators (egorear_b200/train.py, csrc/train_ops.cu) vs autograd through the oracle.

MSDA backward: the oracle's restatement of mmcv's op (oracle/model_ref.ms_deform_attn) is built from differentiable torch
ops, so torch.autograd through it gives the gradients mmcv's backward kernel computes (value, sampling locations,
attention weights).  Losses: nn.MSELoss / the reference's MpjpeLoss formula through torch autograd."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("nh,hd,Q", [(4, 64, 15), (4, 32, 16)])
def test_msda_backward_vs_autograd_of_oracle(nh, hd, Q):
    from egorear_b200.train import MultiScaleDeformableAttnFunction as Fn
    from oracle import model_ref
    B, H, W, P = 3, 64, 64, 16
    g = torch.Generator().manual_seed(8)
    value = torch.randn((B, H * W, nh, hd), generator=g)
    loc = torch.rand((B, Q, nh, 1, P, 2), generator=g) * 1.2 - 0.1            # some points leave the map
    aw = torch.softmax(torch.randn((B, Q, nh, P), generator=g), -1).view(B, Q, nh, 1, P)
    go = torch.randn((B, Q, nh * hd), generator=g)
    # oracle on the CPU (float64 for a clean reference gradient)
    v0, l0, a0 = (t.double().requires_grad_(True) for t in (value, loc, aw))
    out0 = model_ref.ms_deform_attn(v0, H, W, l0, a0)
    out0.backward(go.double())
    v1, l1, a1 = (t.cuda().requires_grad_(True) for t in (value, loc, aw))
    shapes = torch.tensor([[H, W]], dtype=torch.long, device="cuda")
    out1 = Fn.apply(v1, shapes, torch.zeros(1, dtype=torch.long, device="cuda"), l1, a1, 32)
    assert rel(out1.detach().cpu(), out0.detach()) < 1e-5
    out1.backward(go.cuda())
    assert rel(v1.grad.cpu(), v0.grad) < 1e-4
    assert rel(a1.grad.cpu(), a0.grad) < 1e-4
    assert rel(l1.grad.cpu(), l0.grad) < 1e-3        # derivative of the bilinear kernel: differences of neighbouring values, fp32


def test_msdeformattn_module_trains_through_the_extension():
    """the as-executed MSDeformAttn mirror: loss.backward() reaches every parameter when grad is enabled"""
    from egorear_b200 import modules, synth
    m = modules.DeformMultiViewAttn(feat_shape=(64, 64), embed_dim=256, num_heads=4)
    synth.fill_state_dict(m)
    m = m.cuda().train()
    g = torch.Generator().manual_seed(9)
    q = torch.randn((2, 15, 256), generator=g).cuda()
    mem = torch.randn((2, 4096, 256), generator=g).cuda()
    anchors = torch.rand((2, 15, 2), generator=g).cuda()
    out = m(q, mem, anchors)
    out.square().mean().backward()
    for n, p in m.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all() and float(p.grad.abs().max()) > 0, n


def test_losses_forward_backward():
    from egorear_b200 import train
    g = torch.Generator().manual_seed(10)
    pred = torch.randn((4, 4, 15, 64, 64), generator=g).cuda().requires_grad_(True)
    gt = torch.rand((4, 4, 15, 64, 64), generator=g).cuda()
    loss = train.mse_loss(pred, gt)
    ref = torch.nn.functional.mse_loss(pred.detach().double().requires_grad_(True), gt.double())
    assert abs(float(loss) - float(ref)) < 1e-6 * abs(float(ref))
    (loss * 3.0).backward()
    want = 3.0 * 2.0 * (pred.detach() - gt) / pred.numel()
    assert rel(pred.grad, want) < 1e-6
    p3 = (torch.randn((32, 16, 3), generator=g) * 30).cuda().requires_grad_(True)
    g3 = (torch.randn((32, 16, 3), generator=g) * 30).cuda()
    g3[0, 0] = p3.detach()[0, 0]                                             # a zero distance: gradient 0 there, not NaN
    l2 = train.MpjpeLoss()(p3, g3)
    pr = p3.detach().double().requires_grad_(True)
    ref = torch.mean(torch.linalg.norm(g3.double() - pr, dim=-1, ord=2))      # models/utils/pose_metric.py:14-15
    ref.backward()
    assert abs(float(l2) - float(ref)) < 1e-6 * float(ref)
    l2.backward()
    assert torch.isfinite(p3.grad).all()
    assert rel(p3.grad, pr.grad) < 1e-5
