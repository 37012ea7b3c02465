"""GPU: G1 generate_target, D1 get_max_preds, H1' head and P3 reprojection through the C-ABI vs the oracles / golden."""
import numpy as np
import pytest
import torch

from conftest import dense_from_sparse, orc_generate_target, orc_get_max_preds, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from egorear_b200 import ops as o
    return o


def test_generate_target_golden_and_oracle(ops, golden, oracle_lib):
    g = golden["generate_target"]
    want = dense_from_sparse(g["idx"], g["val"], g["shape"])
    got = ops.generate_target_batch(g["joints"], 872, 64, 1.0).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # drop-in signature: numpy in, numpy out, first num_joints rows
    one = ops.generate_target(g["joints"][3], image_size=872, heatmap_size=64, num_joints=15, sigma=1)
    assert one.shape == (15, 64, 64) and np.array_equal(one, want[3, :15])
    # general geometry, sigma = 2 (patch comes from the caller's numpy: bit-exact too)
    want2 = dense_from_sparse(g["idx2"], g["val2"], g["shape2"])
    got2 = ops.generate_target_batch(g["joints2"], 640, 48, 2).cpu().numpy()
    assert np.array_equal(got2.view(np.uint32), want2.view(np.uint32))
    # seeded sweep vs the C oracle, incl. off-image joints (config 4 distribution)
    from egorear_b200 import synth
    kp = synth.synth_keypoints(2048, 4, 16, seed=1)
    got = ops.generate_target_batch(kp).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), orc_generate_target(oracle_lib, kp).view(np.uint32))


def test_generate_target_properties_full_size(ops):
    """size-independent properties on a large batch: support <= 49, max == 1.0 or clipped, decode(render) == mu"""
    from egorear_b200 import synth
    N = 16384
    kp = torch.from_numpy(synth.synth_keypoints(N, 4, 16, seed=2)).cuda()
    hm = ops.generate_target_batch(kp)                                    # [N,4,16,64,64] = 17 GB? no: 4.3 GB
    nnz = (hm != 0).flatten(3).sum(-1)
    assert int(nnz.max()) <= 49
    mu = torch.trunc(kp / 13.625 + 0.5)                                  # int() truncation
    inside = ((mu >= 0) & (mu < 64)).all(-1)
    preds, maxv, valid = ops.get_max_preds(hm.view(N * 4, 16, 64, 64), threshold=1.0, normalize=False)
    preds = preds.view(N, 4, 16, 2)
    assert torch.equal(preds[inside], mu[inside].float())                 # peak sits on mu, value exactly 1.0
    assert bool((maxv.view(N, 4, 16)[inside] == 1.0).all()) and bool(valid.view(N, 4, 16)[inside].all())
    gone = ((mu - 3 >= 64) | (mu + 4 < 0)).any(-1)
    assert int(nnz[gone].sum()) == 0                                      # the skip branch leaves an all-zero map


def test_generate_target_rejects_bad_sigma(ops):
    from egorear_b200._lib import EgrError
    with pytest.raises(EgrError):
        ops.generate_target_batch(np.zeros((1, 1, 2)), sigma=0.5)


def test_get_max_preds_golden_and_edges(ops, golden, oracle_lib):
    from test_oracle import _decode_inputs
    g = golden["get_max_preds"]
    hm = _decode_inputs().cuda()
    for tag, (thr, norm) in {"model": (0.5, True), "eval": (1.0, False)}.items():
        p, m, v = ops.get_max_preds(hm, threshold=thr, normalize=norm)
        assert np.array_equal(p.cpu().numpy().view(np.uint32), g["preds_" + tag].view(np.uint32))
        assert np.array_equal(m.cpu().numpy().view(np.uint32), g["maxvals_" + tag].view(np.uint32))
        assert np.array_equal(v.cpu().numpy(), g["valid_" + tag])
    # squeeze quirk of loss.py:142 and the asserts
    p, m, v = ops.get_max_preds(hm[:1, :1])
    assert m.shape == () and v.shape == ()
    with pytest.raises(AssertionError):
        ops.get_max_preds(hm[0])
    # bit-exact indices vs torch.max on the device and vs the C oracle, at a large size
    g2 = torch.Generator(device="cuda").manual_seed(3)
    big = torch.randn((4096, 15, 64, 64), generator=g2, device="cuda")
    big[::7, 3] = big[::7, 3].round()                                     # many exact ties
    p, m, v, idx = ops.get_max_preds(big, 0.5, True, return_index=True)
    tm, ti = torch.max(big.view(4096, 15, -1), dim=2)
    assert torch.equal(idx.long(), ti) and torch.equal(m, tm)
    op, om, ov, oi = orc_get_max_preds(oracle_lib, big[:64].cpu().numpy(), 0.5, True)
    assert np.array_equal(p[:64].cpu().numpy(), op) and np.array_equal(idx[:64].cpu().numpy(), oi)
    assert np.array_equal(v[:64].cpu().numpy(), ov)


def test_heatmap_head_1x1(ops):
    g = torch.Generator().manual_seed(0)
    feat = torch.randn((8, 128, 64, 64), generator=g)
    w = torch.randn((15, 128, 1, 1), generator=g) * 0.1
    b = torch.randn((15,), generator=g)
    want = torch.nn.functional.conv2d(feat, w, b)
    got = ops.heatmap_head_1x1(feat.cuda(), w.cuda(), b.cuda()).cpu()
    assert rel_err(got.numpy(), want.numpy()) < 1e-5


def test_reproject_fisheye_vs_oracle(ops):
    from egorear_b200 import calib, synth
    from oracle import model_ref
    g = torch.Generator().manual_seed(4)
    B = 64
    pts = torch.stack((torch.rand((B, 16), generator=g) * 120 - 60, torch.rand((B, 16), generator=g) * 120 - 60,
                       torch.rand((B, 16), generator=g) * 190 - 30), dim=-1)
    cams = calib.load_calibration(None)
    for cam in ("ego4view_syn", "ego4view_rw", "ego4view_syn_stereo_front", "ego4view_syn_stereo_back", "ego4view_rw_stereo_back"):
        V = 4 if cam in ("ego4view_syn", "ego4view_rw") else 2
        ctm = synth.synth_coord_trans_mat(B, seed=1)[:, :V].contiguous() if "rw" in cam else None
        p_ref = pts.clone()
        a_ref, v_ref = model_ref.reproject(p_ref, cams, cam, ctm)
        p_gpu = pts.clone().cuda()
        a, v = ops.reproject_fisheye(p_gpu, cam, ctm.cuda() if ctm is not None else None)
        assert float((a.cpu() - a_ref).abs().max()) < 2e-6, cam
        flips = (v.cpu() != v_ref)
        assert int(flips.sum()) == 0 or float((a_ref[flips] * (1 - a_ref[flips])).abs().max()) < 1e-5   # only on the border
        assert torch.allclose(p_gpu.cpu(), p_ref, atol=1e-5), cam         # the in-place quirk is reproduced
        assert 0.05 < float(v_ref.float().mean()) < 0.95
    with pytest.raises(ValueError):
        ops.reproject_fisheye(pts.cuda(), "pinhole")
    with pytest.raises(RuntimeError):
        ops.reproject_fisheye(pts.cuda(), "ego4view_rw", synth.synth_coord_trans_mat(B).double().cuda())


def test_msda_op_vs_oracle(ops):
    from oracle import model_ref
    g = torch.Generator().manual_seed(3)
    for (B, nh, hd, Q) in ((3, 4, 64, 15), (37, 4, 32, 16)):              # odd batch: no im2col_step restriction
        value = torch.randn((B, 4096, nh, hd), generator=g)
        loc = torch.rand((B, Q, nh, 1, 16, 2), generator=g) * 1.3 - 0.15
        aw = torch.softmax(torch.randn((B, Q, nh, 16), generator=g), -1).view(B, Q, nh, 1, 16)
        want = model_ref.ms_deform_attn(value, 64, 64, loc, aw)
        got = ops.MultiScaleDeformableAttnFunction.apply(value.cuda(), torch.tensor([[64, 64]]).cuda(), torch.tensor([0]).cuda(),
                                                         loc.cuda(), aw.cuda(), 32).cpu()
        assert rel_err(got.numpy(), want.numpy()) < 1e-5


def test_soft_argmax_decoder(ops, golden):
    """D1s get_max_preds_soft_pytorch through the C ABI vs the golden vectors of the live reference and the oracle.
    Floating point: coordinates within 1e-3 relative of the map size (64), i.e. 1e-4 px measured margin below."""
    import torch
    from conftest import soft_inputs
    from oracle import model_ref
    g = golden["get_max_preds_soft"]
    hm = soft_inputs()
    for tag, norm in (("raw", False), ("norm", True)):
        p, m = ops.get_max_preds_soft_pytorch(hm.cuda(), normalize=norm)
        assert p.shape == (4, 15, 2) and m.shape == (4, 15, 1)
        scale = 1.0 if norm else 64.0
        assert float(np.abs(p.cpu().numpy() - g["preds_" + tag]).max()) < 1e-4 * scale
        assert np.array_equal(m.cpu().numpy(), g["maxvals_" + tag])
    # larger random batch incl. a non-square map, vs the oracle
    gen = torch.Generator().manual_seed(3)
    big = torch.randn((37, 16, 48, 64), generator=gen) * 3.0
    p, m = ops.get_max_preds_soft_pytorch(big.cuda())
    rp, rm = model_ref.get_max_preds_soft_pytorch(big)
    assert float((p.cpu() - rp).abs().max()) < 5e-3 and torch.equal(m.cpu(), rm)
    with pytest.raises(AssertionError):
        ops.get_max_preds_soft_pytorch(big[0].cuda())
    e, _ = ops.get_max_preds_soft_pytorch(big[:0].cuda())
    assert e.shape == (0, 16, 2)


def test_integrate_tensor_2d_golden_and_oracle(ops, golden):
    """D1i integral decoder (utils/util.py:80-109) vs the live reference's vectors and the oracle"""
    from conftest import INTEGRATE_CASES, INTEGRATE_MAPS, integrate_inputs
    from oracle import model_ref
    g = golden["integrate_tensor_2d"]
    hm = integrate_inputs()
    for tag, sm, mult in INTEGRATE_CASES:
        c, p = ops.integrate_tensor_2d(hm.cuda(), softmax=sm, multiplier=mult)
        assert c.shape == (4, 15, 2) and p.shape == hm.shape and c.is_cuda and p.is_cuda
        assert np.allclose(c.cpu().numpy(), g["coords_" + tag], rtol=2e-6, atol=2e-5, equal_nan=True), tag
        got = np.stack([p[b, j].cpu().numpy() for b, j in INTEGRATE_MAPS])
        assert np.allclose(got, g["maps_" + tag], rtol=2e-6, atol=1e-35), tag
    # seeded sweep incl. a non-square map, against the oracle; every softmax map sums to 1
    gen = torch.Generator().manual_seed(3)
    for shape, sm, mult in (((3, 16, 64, 64), True, 100.0), ((2, 5, 48, 32), True, 7.5), ((2, 5, 48, 32), False, 1.0)):
        x = torch.rand(shape, generator=gen) - (0.0 if sm else 0.3)
        c, p = ops.integrate_tensor_2d(x.cuda(), sm, mult)
        rc, rp = model_ref.integrate_tensor_2d(x, sm, mult)
        assert torch.allclose(c.cpu(), rc, rtol=2e-6, atol=2e-5) and torch.allclose(p.cpu(), rp, rtol=2e-6, atol=1e-35)
        if sm:
            assert torch.allclose(p.sum((2, 3)).cpu(), torch.ones(shape[:2]), atol=1e-5)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.integrate_tensor_2d(hm)


def test_soft_decoders_propagate_nan():
    """a NaN in a map makes maxvals and both coordinates NaN, like torch.max / softmax in the reference
    (utils/loss.py:145-177, utils/util.py:80-109); the other maps are untouched"""
    from egorear_b200 import ops
    from oracle import model_ref
    g = torch.Generator().manual_seed(5)
    hm = torch.randn((2, 3, 64, 64), generator=g)
    hm[1, 2, 17, 40] = float("nan")
    p, m = ops.get_max_preds_soft_pytorch(hm.cuda())
    rp, rm = model_ref.get_max_preds_soft_pytorch(hm)
    assert torch.isnan(p[1, 2]).all() and torch.isnan(m[1, 2]).all() and torch.isnan(rp[1, 2]).all()
    ok = torch.ones((2, 3), dtype=torch.bool); ok[1, 2] = False
    assert torch.allclose(p.cpu()[ok], rp[ok], rtol=1e-4, atol=1e-3)
    for softmax in (True, False):
        c, pr = ops.integrate_tensor_2d(hm.cuda(), softmax=softmax, multiplier=3.0)
        rc, rpr = model_ref.integrate_tensor_2d(hm, softmax=softmax, multiplier=3.0)
        assert torch.isnan(c[1, 2]).all() and torch.isnan(rc[1, 2]).all()
        assert torch.allclose(c.cpu()[ok], rc[ok], rtol=1e-4, atol=1e-3)
        assert not torch.isnan(c.cpu()[ok]).any()


def test_misaligned_views_are_handled():
    """a contiguous view whose storage offset is not a multiple of 16 bytes must not fault the 128-bit kernels"""
    from egorear_b200 import ops
    from oracle import model_ref
    g = torch.Generator().manual_seed(6)
    base = torch.randn((1 + 2 * 15 * 4096,), generator=g)
    hm_c = base[1:].view(2, 15, 64, 64)
    hm = base.cuda()[1:].view(2, 15, 64, 64)
    assert hm.data_ptr() % 16 != 0 and hm.is_contiguous()
    p, m, v = ops.get_max_preds(hm, 0.5, True)
    rp, rm, rv = model_ref.get_max_preds(hm_c, 0.5, True)
    assert torch.equal(p.cpu(), rp) and torch.equal(v.cpu(), rv)
    ps, ms = ops.get_max_preds_soft_pytorch(hm)
    assert torch.allclose(ps.cpu(), model_ref.get_max_preds_soft_pytorch(hm_c)[0], rtol=1e-4, atol=1e-3)
    torch.cuda.synchronize()


def test_generate_target_streams_and_sigmas():
    """the Gaussian patch travels by value, so concurrent calls on two streams with different sigmas cannot
    overwrite each other's patch (there used to be one thread-local device buffer)"""
    from egorear_b200 import ops, synth
    kp = torch.from_numpy(synth.synth_keypoints(256, 4, 16, seed=2)).cuda()
    want1 = ops.generate_target_batch(kp, sigma=1).clone()
    want2 = ops.generate_target_batch(kp, sigma=2).clone()
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for it in range(8):
        with torch.cuda.stream(s1):
            a = ops.generate_target_batch(kp, sigma=1)
        with torch.cuda.stream(s2):
            b = ops.generate_target_batch(kp, sigma=2)
        outs.append((a, b))
    torch.cuda.synchronize()
    for a, b in outs:
        assert torch.equal(a, want1) and torch.equal(b, want2)
