"""GPU: the heatmap-head tail as a stage (egr_head_tail_stage) - `Upsample(2, bilinear, align_corners=True), ReLU,
Conv2d(128, 15, 1)` of conv_heatmap_layers (egoposeformer_heatmap_mvf_ex.py:108-110, :579-583) - against the same three
torch modules in float64 on the same fp16 input.  Two implementations: impl 1 interpolates on the tensor cores (integer
interpolation matrices against z as an MN-major operand, fp32 row blend), impl 0 on the CUDA cores.

Error budget: the interpolated activation is rounded to fp16 once (operand of the 1x1 conv): <= 2^-11 relative per element,
averaged over 128 channels; with `precise` the weights carry no rounding error (hi + lo pair).  Stated bound 3e-4 relative
to the largest heatmap value (precise), 6e-4 otherwise; measured ~1e-4 / ~2e-4."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _run(z, w, bias, wsel, B, G, J, precise, impl, hm_layout="bgj"):
    from egorear_b200 import _lib
    lib = _lib.load()
    dev = z.device
    hm = torch.full((B, G, J, 64, 64), float("nan"), device=dev)
    hm_t = torch.full((G, B, J, 64 * 64), float("nan"), device=dev, dtype=torch.float16 if precise else torch.bfloat16)
    sel = (ctypes.c_int32 * 4)(*([int(v) for v in wsel] + [0] * (4 - len(wsel))))
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.egr_head_tail_stage(ctypes.c_void_p(z.data_ptr()), ctypes.c_void_p(w.data_ptr()), ctypes.c_void_p(bias.data_ptr()),
                                       sel, B, G, J, ctypes.c_void_p(hm.data_ptr()), G * J * 4096, J * 4096,
                                       ctypes.c_void_p(hm_t.data_ptr()), int(precise), int(impl), st))
    torch.cuda.synchronize()
    return hm, hm_t


def _reference(z, w, bias, wsel, B, G, J):
    # z [G][B][1024][128] fp16 -> [B][G][J][64][64] float64
    x = z.double().reshape(G * B, 32, 32, 128).permute(0, 3, 1, 2)
    up = torch.relu(F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True))
    outs = []
    for g in range(G):
        outs.append(F.conv2d(up[g * B:(g + 1) * B], w[wsel[g]].double()[:, :, None, None], bias[wsel[g]].double()))
    return torch.stack(outs, dim=1)


@pytest.mark.parametrize("impl", [1, 0])
@pytest.mark.parametrize("precise", [1, 0])
@pytest.mark.parametrize("B,G,J,wsel", [(3, 4, 15, (0, 0, 1, 1)), (1, 1, 15, (0,)), (37, 2, 15, (1, 0)), (5, 4, 16, (3, 2, 1, 0)), (64, 3, 15, (2, 0, 2))])
def test_head_tail_stage(B, G, J, wsel, precise, impl):
    g = torch.Generator(device="cuda").manual_seed(7 + B)
    z = torch.randn((G, B, 1024, 128), generator=g, device="cuda").half()
    w = torch.randn((4, J, 128), generator=g, device="cuda") * 128 ** -0.5
    bias = torch.randn((4, J), generator=g, device="cuda")
    hm, hm_t = _run(z, w, bias, wsel, B, G, J, precise, impl)
    ref = _reference(z, w, bias, wsel, B, G, J)
    assert torch.isfinite(hm).all()
    err = float((hm.double() - ref).abs().max() / ref.abs().max())
    print("head tail impl %d precise %d B %d: rel err %.2e" % (impl, precise, B, err))
    assert err < (3e-4 if precise else 6e-4)
    # the 16-bit copy is the same values rounded once
    want_t = hm.permute(1, 0, 2, 3, 4).reshape(G, B, J, 4096).to(hm_t.dtype)
    assert torch.equal(hm_t, want_t)


def test_head_tail_impls_agree_exactly_on_grid_points():
    """even output rows / columns of the first and last line sit ON source pixels (weights 1, 0): both implementations must
    reproduce relu(z) . W there to fp32-accumulation accuracy, whatever the interpolation arithmetic"""
    B, G, J = 2, 2, 15
    g = torch.Generator(device="cuda").manual_seed(11)
    z = torch.randn((G, B, 1024, 128), generator=g, device="cuda").half()
    w = torch.randn((1, J, 128), generator=g, device="cuda") * 128 ** -0.5
    bias = torch.zeros((1, J), device="cuda")
    hm1, _ = _run(z, w, bias, (0, 0), B, G, J, 1, 1)
    corner = torch.relu(z.double().reshape(G, B, 32, 32, 128)[:, :, [0, 0, 31, 31], [0, 31, 0, 31], :])      # [G][B][4][128]
    want = torch.einsum("gbpc,jc->bgjp", corner, w[0].double())
    got = hm1[:, :, :, [0, 0, 63, 63], [0, 63, 0, 63]].double()
    assert float((got - want).abs().max() / want.abs().max()) < 1e-5


@pytest.mark.parametrize("impl", [1, 0])
@pytest.mark.parametrize("out_f16", [1, 0])
@pytest.mark.parametrize("n_img", [1, 5, 256])
def test_up2_relu_stage(n_img, out_f16, impl):
    """R1 tail: relu(up2(z)) as a 16-bit channels-last map, both implementations against float64 torch (one rounding of the
    result: <= 2^-11 relative for fp16 + the fp32 interpolation arithmetic, 2^-8 for bf16)"""
    from egorear_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(21 + n_img)
    z = torch.randn((n_img, 1024, 128), generator=g, device="cuda").half()
    dt = torch.float16 if out_f16 else torch.bfloat16
    out = torch.full((n_img, 4096, 128), float("nan"), device="cuda", dtype=dt)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.egr_up2_relu_stage(ctypes.c_void_p(z.data_ptr()), n_img, ctypes.c_void_p(out.data_ptr()), out_f16, impl, st))
    torch.cuda.synchronize()
    x = z.double().reshape(n_img, 32, 32, 128).permute(0, 3, 1, 2)
    ref = torch.relu(F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)).permute(0, 2, 3, 1).reshape(n_img, 4096, 128)
    assert torch.isfinite(out).all()
    err = float((out.double() - ref).abs().max() / ref.abs().max())
    # element by element: the correctly interpolated value rounded once (2^-11 / 2^-8 relative) + 5e-6 absolute for the
    # CUDA-core kernel's fp32 source coordinates (scale * dst in float, as PyTorch computes them; the tensor-core kernel's
    # integer weights are exact)
    rtol = 5e-4 if out_f16 else 4e-3
    excess = float(((out.double() - ref).abs() - rtol * ref.abs()).max())
    print("up2_relu impl %d f16 %d n %d: max err / max ref %.2e, worst excess over %.0e relative: %.2e" % (impl, out_f16, n_img, err, rtol, excess))
    # (the CUDA-core kernel interpolates a bf16-mode map in half2 arithmetic: one rounding per HFMA2)
    assert excess < (1e-7 if impl else 5e-6 if out_f16 else 5e-3)
