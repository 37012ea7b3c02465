"""GPU: one dense stage (egr_dense_stage) — the tcgen05 + TMA bf16 kernel and the fp32 SIMT kernel vs plain torch fp32
on the same bf16-rounded operands.

The tensor-core kernel accumulates bf16 x bf16 products in fp32 (TMEM); on operands that are already bf16 values the only
difference from the fp32 torch reference is summation order, so the bound is tight (2e-3 relative to the output range for
bf16 outputs = one bf16 ulp, 1e-4 for fp32 outputs).
"""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def dense(A, W, bias, D, M, N, K, lda, ldd, amode=0, epi=0, aux=None, kblk=0, kblk_stride=0, Hin=0, Win=0, Cin=0, Hout=0,
          Wout=0, groups=1, a_gs=0, w_gs=0, b_gs=0, d_gs=0, aux_gs=0, use_tc=1, ka=0, out_pair=0):
    from egorear_b200 import _lib
    lib = _lib.load()
    d = _lib.DenseDesc()
    d.A, d.W, d.bias, d.D, d.aux = _ptr(A), _ptr(W), _ptr(bias), _ptr(D), _ptr(aux)
    d.M, d.N, d.K, d.lda, d.ldd, d.amode, d.epi = M, N, K, lda, ldd, amode, epi
    d.kblk, d.kblk_stride = kblk, kblk_stride
    d.Hin, d.Win, d.Cin, d.Hout, d.Wout = Hin, Win, Cin, Hout, Wout
    d.groups, d.a_gs, d.w_gs, d.b_gs, d.d_gs, d.aux_gs = groups, a_gs, w_gs, b_gs, d_gs, aux_gs
    code = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}
    d.a_is_bf16 = code[A.dtype]
    d.d_is_bf16 = code[D.dtype]
    d.use_tc = use_tc
    d.ka = ka
    d.out_pair = out_pair
    _lib.check(lib.egr_dense_stage(ctypes.byref(d), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()


def act(x, epi):
    if epi in (1, 3):
        return torch.relu(x)
    if epi == 2:
        return F.gelu(x)
    return x


def check(got, want, out_dtype):
    tol = 6e-3 if out_dtype == torch.bfloat16 else 2e-4
    err = float((got.float() - want).abs().max() / want.abs().max().clamp_min(1e-20))
    assert err < tol, "rel err %.3e" % err
    return err


def rnd(shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(torch.bfloat16)


@pytest.mark.parametrize("M,N,K,groups,epi,out_dtype", [
    (128, 64, 64, 1, 0, torch.float32),          # one tile, one k-block
    (256, 128, 128, 1, 1, torch.bfloat16),
    (960, 256, 4096, 4, 1, torch.float32),       # Q1a: ragged M (960 = 7.5 tiles), long K, 4 weight sets
    (4096, 256, 128, 2, 1, torch.bfloat16),      # F1a-like
    (2048, 512, 256, 1, 2, torch.bfloat16),      # BN=256 path when tiles are plentiful / GELU
    (148 * 128 * 2, 256, 192, 1, 1, torch.bfloat16),   # >= 148 tiles at BN=256, 3 k-blocks, several tiles per CTA
    (37 * 128 + 5, 128, 576, 3, 0, torch.float32),
])
def test_tc_plain(M, N, K, groups, epi, out_dtype):
    A = rnd((groups, M, K), 1)
    W = rnd((groups, N, K), 2, K ** -0.5)
    bias = torch.randn((groups, N), device="cuda")
    D = torch.full((groups, M, N), float("nan"), device="cuda", dtype=out_dtype)
    dense(A, W, bias, D, M, N, K, K, N, epi=epi, groups=groups, a_gs=M * K, w_gs=N * K, b_gs=N, d_gs=M * N)
    want = act(torch.einsum("gmk,gnk->gmn", A.float(), W.float()) + bias[:, None, :], epi)
    check(D, want, out_dtype)


def test_tc_strided_rows_and_output():
    """lda > K and ldd > N (a stage writing into a wider buffer)"""
    M, N, K = 512, 64, 128
    Abuf = rnd((M, 256), 3)
    W = rnd((N, K), 4, K ** -0.5)
    Dbuf = torch.zeros((M, 192), device="cuda", dtype=torch.bfloat16)
    dense(Abuf[:, 64:], W, None, Dbuf[:, 128:], M, N, K, 256, 192)
    want = Abuf[:, 64:64 + K].float() @ W.float().t()
    check(Dbuf[:, 128:], want, torch.bfloat16)
    assert float(Dbuf[:, :128].abs().max()) == 0.0


@pytest.mark.parametrize("B", [2, 64])
def test_tc_ksplit_blocks_and_splitk(B):
    """pose3d Linear(32768 -> 2048): A is [V][B][8192] (k blocks of 8192 one view apart); small M triggers split-K"""
    V, KB, N = 4, 8192, 2048
    A = rnd((V, B, KB), 5)
    W = rnd((N, V * KB), 6, (V * KB) ** -0.5)
    bias = torch.randn((N,), device="cuda")
    D = torch.full((B, N), float("nan"), device="cuda")
    dense(A, W, bias, D, B, N, V * KB, KB, N, epi=2, kblk=KB, kblk_stride=B * KB)
    want = F.gelu(A.float().permute(1, 0, 2).reshape(B, V * KB) @ W.float().t() + bias)
    check(D, want, torch.float32)


@pytest.mark.parametrize("Hin,Cin,N,n_img,groups,out_dtype", [
    (64, 128, 256, 2, 1, torch.bfloat16),       # H1b / H2a geometry
    (64, 256, 512, 1, 2, torch.bfloat16),       # F1b geometry, two weight sets
    (64, 64, 128, 3, 1, torch.float32),         # P2b
    (16, 64, 128, 4, 1, torch.bfloat16),        # P2d: 8x8 outputs, two images per tile
])
def test_tc_conv3s2(Hin, Cin, N, n_img, groups, out_dtype):
    x = rnd((groups, n_img, Hin, Hin, Cin), 7)                       # NHWC
    w = rnd((groups, N, Cin, 3, 3), 8, (9 * Cin) ** -0.5)            # torch conv layout
    bias = torch.randn((groups, N), device="cuda")
    Wp = w.permute(0, 1, 3, 4, 2).contiguous().reshape(groups, N, 9 * Cin)   # [N][ky][kx][Cin]
    Ho = Hin // 2
    M = n_img * Ho * Ho
    D = torch.full((groups, M, N), float("nan"), device="cuda", dtype=out_dtype)
    dense(x, Wp, bias, D, M, N, 9 * Cin, 0, N, amode=1, epi=1, Hin=Hin, Win=Hin, Cin=Cin, groups=groups,
          a_gs=n_img * Hin * Hin * Cin, w_gs=N * 9 * Cin, b_gs=N, d_gs=M * N)
    for g in range(groups):
        ref = F.conv2d(x[g].float().permute(0, 3, 1, 2), w[g].float(), bias[g], stride=2, padding=1)
        want = torch.relu(ref).permute(0, 2, 3, 1).reshape(M, N)
        check(D[g], want, out_dtype)


def test_tc_relu_addup():
    """F1c epilogue: relu(acc + bias) + relu(bilinear_x2_align_corners(aux))"""
    n_img, Ho, N, K = 3, 32, 128, 512
    M = n_img * Ho * Ho
    A = rnd((M, K), 9)
    W = rnd((N, K), 10, K ** -0.5)
    bias = torch.randn((N,), device="cuda")
    aux = rnd((n_img, 16 * 16, N), 11)
    D = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
    dense(A, W, bias, D, M, N, K, K, N, epi=3, aux=aux, Hout=Ho, Wout=Ho)
    up = F.interpolate(aux.float().reshape(n_img, 16, 16, N).permute(0, 3, 1, 2), scale_factor=2, mode="bilinear",
                       align_corners=True)
    want = torch.relu(A.float() @ W.float().t() + bias) + torch.relu(up).permute(0, 2, 3, 1).reshape(M, N)
    check(D, want, torch.bfloat16)


def test_tc_matches_simt_kernel():
    """the two dense kernels agree on the same operands (fp32 SIMT on bf16-valued inputs)"""
    M, N, K = 1000, 128, 256
    A = rnd((M, K), 12)
    W = rnd((N, K), 13, K ** -0.5)
    bias = torch.randn((N,), device="cuda")
    D1 = torch.empty((M, N), device="cuda")
    D2 = torch.empty((M, N), device="cuda")
    dense(A, W, bias, D1, M, N, K, K, N, epi=1, use_tc=1)
    dense(A, W.float().contiguous(), bias, D2, M, N, K, K, N, epi=1, use_tc=0)
    check(D1, D2, torch.float32)


def test_tc_rejects_bad_shapes():
    from egorear_b200 import _lib
    A = rnd((128, 96), 1)
    W = rnd((64, 96), 2)
    D = torch.empty((128, 64), device="cuda")
    with pytest.raises(_lib.EgrError):
        dense(A, W, None, D, 128, 64, 96, 96, 64)        # K not a multiple of 64


def tf32_trunc(x):
    return (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("M,N,K,epi", [(4096, 64, 128, 1), (700, 128, 96, 0), (64, 2048, 4096, 2)])
def test_tc_tf32_plain(M, N, K, epi):
    """fp32 operands through kind::tf32 (pose3d proposal branch): 10-bit mantissa products, fp32 accumulate"""
    g = torch.Generator(device="cuda").manual_seed(21)
    A = torch.randn((M, K), generator=g, device="cuda")
    W = torch.randn((N, K), generator=g, device="cuda") * K ** -0.5
    bias = torch.randn((N,), device="cuda")
    D = torch.full((M, N), float("nan"), device="cuda")
    dense(A, W, bias, D, M, N, K, K, N, epi=epi)
    want = act(A.double() @ W.double().t() + bias.double(), epi).float()
    err = float((D - want).abs().max() / want.abs().max())
    want_t = act(tf32_trunc(A).double() @ tf32_trunc(W).double().t() + bias.double(), epi).float()
    err_t = float((D - want_t).abs().max() / want_t.abs().max())
    print("tf32 rel err vs fp64 %.2e, vs truncated-operand fp64 %.2e" % (err, err_t))
    assert err < 2e-3


def test_tc_tf32_conv3s2():
    Hin, Cin, N, n_img = 64, 64, 128, 2
    g = torch.Generator(device="cuda").manual_seed(22)
    x = torch.randn((n_img, Hin, Hin, Cin), generator=g, device="cuda")
    w = torch.randn((N, Cin, 3, 3), generator=g, device="cuda") * (9 * Cin) ** -0.5
    bias = torch.randn((N,), device="cuda")
    Wp = w.permute(0, 2, 3, 1).contiguous().reshape(N, 9 * Cin)
    M = n_img * (Hin // 2) ** 2
    D = torch.full((M, N), float("nan"), device="cuda")
    dense(x, Wp, bias, D, M, N, 9 * Cin, 0, N, amode=1, epi=1, Hin=Hin, Win=Hin, Cin=Cin)
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double(), bias.double(), stride=2, padding=1)
    want = torch.relu(ref).permute(0, 2, 3, 1).reshape(M, N).float()
    err = float((D - want).abs().max() / want.abs().max())
    assert err < 2e-3, err


@pytest.mark.parametrize("M,N,K,epi,out_dtype", [(4096, 64, 128, 1, torch.float16), (700, 128, 96 + 32, 0, torch.float32),
                                                  (64, 2048, 4096, 2, torch.float32)])
def test_tc_fp16_plain(M, N, K, epi, out_dtype):
    """fp16 operands through kind::f16 (pose3d proposal branch): TF32's 10-bit mantissa at half the bytes"""
    g = torch.Generator(device="cuda").manual_seed(31)
    A = torch.randn((M, K), generator=g, device="cuda").half()
    W = (torch.randn((N, K), generator=g, device="cuda") * K ** -0.5).half()
    bias = torch.randn((N,), device="cuda")
    D = torch.full((M, N), float("nan"), device="cuda", dtype=out_dtype)
    dense(A, W, bias, D, M, N, K, K, N, epi=epi)
    want = act(A.double() @ W.double().t() + bias.double(), epi).float()
    err = float((D.float() - want).abs().max() / want.abs().max())
    assert err < (2e-3 if out_dtype == torch.float16 else 2e-4), err


def test_tc_fp16_conv3s2_and_saturation():
    Hin, Cin, N, n_img = 64, 64, 128, 2
    g = torch.Generator(device="cuda").manual_seed(32)
    x = torch.randn((n_img, Hin, Hin, Cin), generator=g, device="cuda").half()
    w = (torch.randn((N, Cin, 3, 3), generator=g, device="cuda") * (9 * Cin) ** -0.5).half()
    bias = torch.randn((N,), device="cuda")
    Wp = w.permute(0, 2, 3, 1).contiguous().reshape(N, 9 * Cin)
    M = n_img * (Hin // 2) ** 2
    D = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float16)
    dense(x, Wp, bias, D, M, N, 9 * Cin, 0, N, amode=1, epi=1, Hin=Hin, Win=Hin, Cin=Cin)
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double(), bias.double(), stride=2, padding=1)
    want = torch.relu(ref).permute(0, 2, 3, 1).reshape(M, N).float()
    assert float((D.float() - want).abs().max() / want.abs().max()) < 2e-3
    # out-of-range results saturate to +-65504 instead of inf
    A = torch.full((128, 64), 200.0, device="cuda", dtype=torch.float16)
    W2 = torch.full((64, 64), 100.0, device="cuda", dtype=torch.float16)
    W2[1::2] *= -1
    D2 = torch.zeros((128, 64), device="cuda", dtype=torch.float16)
    dense(A, W2, None, D2, 128, 64, 64, 64, 64)
    assert torch.isfinite(D2).all() and float(D2.max()) == 65504.0 and float(D2.min()) == -65504.0


def _split16(w):
    hi = w.half()
    lo = (w - hi.float()).half()
    return hi, lo


@pytest.mark.parametrize("pair", [1, 0])
@pytest.mark.parametrize("M,N,K,groups", [(4096, 256, 128, 1), (1000, 128, 256, 4), (960, 256, 4096, 4), (148 * 128 + 77, 128, 128, 2),
                                          (40000, 256, 256, 1)])
def test_tc_fp16_split_weights(M, N, K, groups, pair):
    """EGR_PREC_FP16 dense stage: W = [W_hi | W_lo] along K (K' = 2K, ka = K: the k loop re-reads A): the result equals
    A_fp16 x W_fp32 to fp32-accumulation accuracy, i.e. the weight rounding error is gone (a single fp16 W is ~5e-4 off)"""
    g = torch.Generator(device="cuda").manual_seed(41)
    A = torch.randn((groups, M, K), generator=g, device="cuda").half()
    W = torch.randn((groups, N, K), generator=g, device="cuda") * K ** -0.5
    bias = torch.randn((groups, N), device="cuda")
    hi, lo = _split16(W)
    Ws = torch.cat([hi, lo], dim=2).contiguous()                       # [g][N][2K]
    D = torch.full((groups, M, N), float("nan"), device="cuda")
    from egorear_b200 import engine
    # pair=1: ring stages [A | W_hi | W_lo] (tiles up to 128 wide: A crosses L2 -> smem once); pair=0: the K wrap (A re-read)
    engine.set_option("pair", pair)
    try:
        dense(A, Ws, bias, D, M, N, 2 * K, K, N, epi=1, groups=groups, a_gs=M * K, w_gs=N * 2 * K, b_gs=N, d_gs=M * N, ka=K)
    finally:
        engine.set_option("pair", 0)
    want = torch.relu(torch.einsum("gmk,gnk->gmn", A.double(), W.double()) + bias.double()[:, None, :]).float()
    err = float((D - want).abs().max() / want.abs().max())
    single = torch.relu(torch.einsum("gmk,gnk->gmn", A.double(), hi.double()) + bias.double()[:, None, :]).float()
    err_single = float((single - want).abs().max() / want.abs().max())
    print("split-weight rel err %.2e (single fp16 weight: %.2e)" % (err, err_single))
    assert err < 2e-5 and err < 0.2 * err_single


def test_tc_fp16_split_weights_conv3s2():
    Hin, Cin, N, n_img = 64, 128, 256, 3
    g = torch.Generator(device="cuda").manual_seed(42)
    x = torch.randn((n_img, Hin, Hin, Cin), generator=g, device="cuda").half()
    w = torch.randn((N, Cin, 3, 3), generator=g, device="cuda") * (9 * Cin) ** -0.5
    bias = torch.randn((N,), device="cuda")
    Wp = w.permute(0, 2, 3, 1).contiguous().reshape(N, 9 * Cin)
    hi, lo = _split16(Wp)
    Ws = torch.cat([hi, lo], dim=1).contiguous()
    M = n_img * (Hin // 2) ** 2
    D = torch.full((M, N), float("nan"), device="cuda")
    dense(x, Ws, bias, D, M, N, 2 * 9 * Cin, 0, N, amode=1, epi=1, Hin=Hin, Win=Hin, Cin=Cin, ka=9 * Cin)
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double(), bias.double(), stride=2, padding=1)
    want = torch.relu(ref).permute(0, 2, 3, 1).reshape(M, N).float()
    assert float((D - want).abs().max() / want.abs().max()) < 2e-5


def tf32_rna(x):
    """nearest TF32 value, ties away (cvt.rna.tf32.f32)"""
    u = x.contiguous().view(torch.int32)
    return ((u + 0x1000) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("M,N,K,groups", [(960, 256, 256, 4), (960, 256, 3200, 4), (64, 256, 512, 4)])
def test_tc_tf32x3(M, N, K, groups):
    """the fp32-grade token Linear of EGR_PREC_FP16: rows [x | x_lo] (x raw fp32: the tensor core truncates it to TF32),
    weights [W_hi | W_hi | W_lo], K' = 3K with ka = 2K.  Error vs float64 ~1e-6, three orders below plain TF32."""
    g = torch.Generator(device="cuda").manual_seed(43)
    X = torch.randn((groups, M, K), generator=g, device="cuda")
    W = torch.randn((groups, N, K), generator=g, device="cuda") * K ** -0.5
    x_hi = tf32_trunc(X)
    A = torch.cat([X, tf32_rna(X - x_hi)], dim=2).contiguous()          # [g][M][2K]
    w_hi = tf32_rna(W)
    Ws = torch.cat([w_hi, w_hi, tf32_rna(W - w_hi)], dim=2).contiguous()   # [g][N][3K]
    D = torch.full((groups, M, N), float("nan"), device="cuda")
    dense(A, Ws, None, D, M, N, 3 * K, 2 * K, N, groups=groups, a_gs=M * 2 * K, w_gs=N * 3 * K, d_gs=M * N, ka=2 * K)
    want = torch.einsum("gmk,gnk->gmn", X.double(), W.double()).float()
    err = float((D - want).abs().max() / want.abs().max())
    plain = torch.einsum("gmk,gnk->gmn", tf32_rna(X).double(), w_hi.double()).float()
    err_plain = float((plain - want).abs().max() / want.abs().max())
    print("3xTF32 rel err %.2e (plain TF32: %.2e)" % (err, err_plain))
    # the residual error is the tensor core's own fp32 accumulation (partial sums are aligned and truncated, ~K * 2^-24 relative):
    # 4e-6 at K = 256, 4e-5 at K = 3200 — still an order of magnitude below plain TF32 and far inside the 1e-3 budget
    assert err < 1e-4 and err < 0.2 * err_plain


@pytest.mark.parametrize("H,Cin,N,n_img,groups", [(64, 64, 64, 3, 1), (32, 128, 128, 4, 2), (16, 256, 256, 8, 1), (8, 512, 512, 6, 2),
                                                   (64, 128, 128, 2, 1)])
def test_tc_conv3s1(H, Cin, N, n_img, groups):
    """implicit 3x3 stride-1 pad-1 conv (ResNet18 BasicBlock convs, EfficientFPN fpn_convs; models/backbones/resnet.py): one
    4-D TMA box per tap, the zero padding is the unit's out-of-bounds fill in x and y of every image"""
    g = torch.Generator(device="cuda").manual_seed(51)
    x = torch.randn((groups, n_img, H, H, Cin), generator=g, device="cuda").to(torch.bfloat16)
    w = (torch.randn((groups, N, Cin, 3, 3), generator=g, device="cuda") * (9 * Cin) ** -0.5).to(torch.bfloat16)
    bias = torch.randn((groups, N), device="cuda")
    Wp = w.permute(0, 1, 3, 4, 2).contiguous().reshape(groups, N, 9 * Cin)
    M = n_img * H * H
    D = torch.full((groups, M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    dense(x, Wp, bias, D, M, N, 9 * Cin, 0, N, amode=2, epi=1, Hin=H, Win=H, Cin=Cin, groups=groups,
          a_gs=M * Cin, w_gs=N * 9 * Cin, b_gs=N, d_gs=M * N)
    for gi in range(groups):
        ref = F.conv2d(x[gi].double().permute(0, 3, 1, 2), w[gi].double(), bias[gi].double(), stride=1, padding=1)
        want = torch.relu(ref).permute(0, 2, 3, 1).reshape(M, N).float()
        check(D[gi], want, torch.bfloat16)


def test_tc_conv1x1_stride2_centre_tap():
    """ResNet's downsample branch: 1x1 stride-2 conv = the 3x3 stride-2 mode with K == Cin (centre tap only)"""
    Hin, Cin, N, n_img = 64, 64, 128, 4
    g = torch.Generator(device="cuda").manual_seed(52)
    x = torch.randn((n_img, Hin, Hin, Cin), generator=g, device="cuda").to(torch.bfloat16)
    w = (torch.randn((N, Cin), generator=g, device="cuda") * Cin ** -0.5).to(torch.bfloat16)
    bias = torch.randn((N,), device="cuda")
    M = n_img * (Hin // 2) ** 2
    D = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    dense(x, w, bias, D, M, N, Cin, 0, N, amode=1, epi=0, Hin=Hin, Win=Hin, Cin=Cin)
    want = (x[:, ::2, ::2].double().reshape(M, Cin) @ w.double().t() + bias.double()).float()
    check(D, want, torch.bfloat16)


@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float16])
def test_tc_residual_epilogues(out_dtype):
    """epi 4: relu(acc + bias + aux[M][N]) (BasicBlock: out += identity; relu) on a conv3s1 stage; epi 5: relu(acc + bias +
    up2(aux)) (EfficientFPN fuse conv with its low-resolution half commuted in front of the upsample)"""
    H, Cin, N, n_img = 32, 128, 128, 3
    g = torch.Generator(device="cuda").manual_seed(53)
    x = torch.randn((n_img, H, H, Cin), generator=g, device="cuda").to(out_dtype)
    w = (torch.randn((N, Cin, 3, 3), generator=g, device="cuda") * (9 * Cin) ** -0.5).to(out_dtype)
    bias = torch.randn((N,), device="cuda")
    M = n_img * H * H
    idn = torch.randn((M, N), generator=g, device="cuda").to(out_dtype)
    D = torch.full((M, N), float("nan"), device="cuda", dtype=out_dtype)
    dense(x, w.permute(0, 2, 3, 1).contiguous().reshape(N, 9 * Cin), bias, D, M, N, 9 * Cin, 0, N, amode=2, epi=4, aux=idn,
          Hin=H, Win=H, Cin=Cin)
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double(), bias.double(), stride=1, padding=1).permute(0, 2, 3, 1).reshape(M, N)
    check(D, torch.relu(ref + idn.double()).float(), torch.bfloat16)
    # epi 5 on a 1x1 conv
    K = 128
    A = torch.randn((M, K), generator=g, device="cuda").to(out_dtype)
    W = (torch.randn((N, K), generator=g, device="cuda") * K ** -0.5).to(out_dtype)
    low = torch.randn((n_img, H // 2, H // 2, N), generator=g, device="cuda").to(out_dtype)
    D2 = torch.full((M, N), float("nan"), device="cuda", dtype=out_dtype)
    dense(A, W, bias, D2, M, N, K, K, N, epi=5, aux=low, Hout=H, Wout=H)
    up = F.interpolate(low.double().permute(0, 3, 1, 2), scale_factor=2, mode="bilinear", align_corners=True).permute(0, 2, 3, 1).reshape(M, N)
    check(D2, torch.relu(A.double() @ W.double().t() + bias.double() + up).float(), torch.bfloat16)


@pytest.mark.parametrize("M,N,K,epi", [(4096, 128, 256, 1), (1000, 256, 128, 0), (148 * 128 + 77, 128, 512, 1), (2048, 512, 128, 0)])
def test_tc_fp16_pair_chain(M, N, K, epi):
    """split ACTIVATIONS between two fp16 stages: stage 1 writes its output as the pair [hi | lo] (out_pair), stage 2 reads it
    as A = [x_hi | x_lo] against W = [W_hi | W_hi | W_lo] (ka = 2N, K' = 3N).  The chain then equals the float64 chain of the
    same fp16 INPUT to fp32-accumulation accuracy: no rounding of the intermediate, none of either weight matrix."""
    g = torch.Generator(device="cuda").manual_seed(44)
    A = torch.randn((M, K), generator=g, device="cuda").half()
    W1 = torch.randn((N, K), generator=g, device="cuda") * K ** -0.5
    b1 = torch.randn((N,), generator=g, device="cuda")
    N2 = 128
    W2 = torch.randn((N2, N), generator=g, device="cuda") * N ** -0.5
    b2 = torch.randn((N2,), generator=g, device="cuda")
    h1, l1 = _split16(W1)
    mid = torch.full((M, 2 * N), float("nan"), device="cuda", dtype=torch.float16)
    dense(A, torch.cat([h1, l1], dim=1).contiguous(), b1, mid, M, N, 2 * K, K, 2 * N, epi=epi, ka=K, out_pair=1)
    y1 = torch.einsum("mk,nk->mn", A.double(), W1.double()) + b1.double()
    if epi == 1:
        y1 = torch.relu(y1)
    hi, lo = mid[:, :N].double(), mid[:, N:].double()
    assert torch.isfinite(mid).all()
    err_mid = float((hi + lo - y1).abs().max() / y1.abs().max())
    err_hi = float((hi - y1).abs().max() / y1.abs().max())
    assert err_mid < 2e-5 and err_hi > 10 * err_mid
    h2, l2 = _split16(W2)
    out = torch.full((M, N2), float("nan"), device="cuda")
    dense(mid, torch.cat([h2, h2, l2], dim=1).contiguous(), b2, out, M, N2, 3 * N, 2 * N, N2, ka=2 * N)
    want = (torch.einsum("mk,nk->mn", y1, W2.double()) + b2.double()).float()
    err = float((out - want).abs().max() / want.abs().max())
    single = (torch.einsum("mk,nk->mn", hi, W2.double()) + b2.double()).float()
    err_single = float((single - want).abs().max() / want.abs().max())
    print("pair chain rel err %.2e (rounded intermediate: %.2e)" % (err, err_single))
    assert err < 3e-5 and err < 0.2 * err_single
