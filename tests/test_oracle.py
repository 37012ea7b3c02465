"""CPU: pins the C oracle (oracle/gt_decode.c) to the golden vectors produced by the live reference, and — when the
reference checkout is present (build container) — to the reference functions themselves."""
import numpy as np
import pytest
import torch

from conftest import dense_from_sparse, orc_generate_target, orc_get_max_preds
from oracle import ref_import


def test_generate_target_matches_golden(golden, oracle_lib):
    g = golden["generate_target"]
    want = dense_from_sparse(g["idx"], g["val"], g["shape"])
    got = orc_generate_target(oracle_lib, g["joints"])
    assert got.dtype == np.float32 and np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_generate_target_gaussian_table(golden, oracle_lib):
    """the 10 distinct fp32 values of the sigma=1 patch and the edge-case KATs of §8c"""
    out = orc_generate_target(oracle_lib, np.array([[[436.0, 436.0]]]))[0, 0]
    bits = sorted(set(out[out != 0].view(np.uint32).tolist()), reverse=True)
    assert bits == [0x3F800000, 0x3F1B4598, 0x3EBC5AB1, 0x3E0A9555, 0x3DA81C2F, 0x3C960AAE, 0x3C360282, 0x3BDCC9FE,
                    0x3AC50F0C, 0x39016791]
    kat = {(-100.0, -100.0): (0, 0.0), (-30.0, 400.0): (21, 0.60653067), (871.9, 871.9): (9, None), (900.0, 900.0): (1, 1.2340980e-4),
           (915.0, 400.0): (0, 0.0), (0.0, 0.0): (16, 1.0)}
    for (x, y), (nnz, mx) in kat.items():
        o = orc_generate_target(oracle_lib, np.array([[[x, y]]]))[0, 0]
        assert int((o != 0).sum()) == nnz, (x, y)
        if mx is not None:
            assert abs(float(o.max()) - mx) < 1e-7
    # int() truncates toward zero: x = -20 -> mu 0
    o = orc_generate_target(oracle_lib, np.array([[[-20.0, -20.0]]]))[0, 0]
    assert o[0, 0] == 1.0


def test_generate_target_general_geometry(golden, oracle_lib):
    g = golden["generate_target"]
    want = dense_from_sparse(g["idx2"], g["val2"], g["shape2"])
    got = orc_generate_target(oracle_lib, g["joints2"], image_size=640, hs=48, sigma=2.0)
    # sigma != 1 goes through expf: bit-exact on the support, values within 1 ulp of numpy's exp
    assert np.array_equal(got != 0, want != 0)
    assert np.max(np.abs(got.view(np.int32).astype(np.int64) - want.view(np.int32).astype(np.int64))) <= 1


def _decode_inputs():
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    # only the input factory is needed; avoid executing the reference import at module import time
    src = open(spec.origin).read()
    ns = {"torch": torch}
    start = src.index("def decode_inputs()")
    end = src.index("def golden_decode")
    exec(src[start:end], ns)
    return ns["decode_inputs"]()


def test_get_max_preds_matches_golden(golden, oracle_lib):
    g = golden["get_max_preds"]
    hm = _decode_inputs().numpy()
    for tag, (thr, norm) in {"model": (0.5, True), "eval": (1.0, False)}.items():
        p, m, v, idx = orc_get_max_preds(oracle_lib, hm, thr, norm)
        assert np.array_equal(p.view(np.uint32), g["preds_" + tag].view(np.uint32))
        assert np.array_equal(m.view(np.uint32), g["maxvals_" + tag].view(np.uint32))   # NaN bit pattern included
        assert np.array_equal(v, g["valid_" + tag])
    # tie -> first index; all-equal -> 0; NaN wins
    _, _, _, idx = orc_get_max_preds(oracle_lib, hm, 0.5, False)
    assert idx[0, 0] == 0 and idx[0, 1] == 10 * 64 + 20 and idx[0, 2] == 4095 and idx[0, 3] == 7 * 64 + 9


@pytest.mark.skipif(not ref_import.available(), reason="reference checkout not present (GPU box)")
def test_oracle_vs_live_reference(oracle_lib):
    fns = ref_import.import_functions()
    rng = np.random.default_rng(5)
    joints = rng.uniform(-80, 950, size=(200, 16, 2))
    want = np.stack([fns["generate_target"](j, 872, 64, 16, 1.0) for j in joints])
    got = orc_generate_target(oracle_lib, joints)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    hm = torch.randn(5, 15, 64, 64)
    p, m, v = fns["get_max_preds"](hm.clone(), 0.5, True)
    op, om, ov, _ = orc_get_max_preds(oracle_lib, hm.numpy(), 0.5, True)
    assert np.array_equal(op, p.numpy()) and np.array_equal(om, m.numpy()) and np.array_equal(ov, v.numpy())


def test_soft_argmax_oracle_matches_golden(golden):
    """oracle restatement of get_max_preds_soft_pytorch (utils/loss.py:145-177) vs vectors from the live reference"""
    from conftest import soft_inputs
    from oracle import model_ref
    g = golden["get_max_preds_soft"]
    hm = soft_inputs()
    for tag, norm in (("raw", False), ("norm", True)):
        p, m = model_ref.get_max_preds_soft_pytorch(hm, norm)
        assert p.shape == (4, 15, 2) and m.shape == (4, 15, 1)
        assert np.allclose(p.numpy(), g["preds_" + tag], rtol=0, atol=1e-5)
        assert np.array_equal(m.numpy(), g["maxvals_" + tag])
    # known answers: constant map and two equal opposite peaks -> the centre; one dominant cell -> that cell
    assert np.allclose(g["preds_raw"][3, 0], [31.5, 31.5], atol=1e-4) and np.allclose(g["preds_raw"][3, 2], [31.5, 31.5], atol=1e-4)
    assert np.allclose(g["preds_raw"][3, 1], [60.0, 5.0], atol=1e-4)
    if ref_import.available():
        fn = ref_import.import_functions()["get_max_preds_soft_pytorch"]
        p2, m2 = fn(hm.clone(), False)
        assert np.array_equal(p2.numpy(), g["preds_raw"]) and np.array_equal(m2.numpy(), g["maxvals_raw"])


def test_integrate_tensor_2d_oracle_matches_golden(golden):
    """oracle restatement of integrate_tensor_2d (utils/util.py:80-109) vs vectors from the live reference"""
    from conftest import INTEGRATE_CASES, INTEGRATE_MAPS, integrate_inputs
    from oracle import model_ref
    g = golden["integrate_tensor_2d"]
    hm = integrate_inputs()
    for tag, sm, mult in INTEGRATE_CASES:
        c, p = model_ref.integrate_tensor_2d(hm, sm, mult)
        assert c.shape == (4, 15, 2) and p.shape == hm.shape
        assert np.allclose(c.numpy(), g["coords_" + tag], rtol=1e-6, atol=1e-5, equal_nan=True)
        got = np.stack([p[b, j].numpy() for b, j in INTEGRATE_MAPS])
        assert np.allclose(got, g["maps_" + tag], rtol=1e-6, atol=0)
    # known answers: constant map -> centre; one dominant cell -> that cell; zero relu mass -> NaN (0/0, like torch)
    assert np.allclose(g["coords_sm1"][3, 0], [31.5, 31.5], atol=1e-4)
    assert np.allclose(g["coords_sm100"][3, 1], [60.0, 5.0], atol=1e-4)
    assert np.isnan(g["coords_relu"][3, 3]).all() and np.isfinite(g["coords_sm100"]).all()


def test_preprocess_oracle_matches_golden(golden):
    """oracle restatement of Pillow's bicubic resize + ToTensor + Normalize, bit-for-bit against PIL + torchvision vectors"""
    import hashlib
    from conftest import PREPROCESS_CASES
    from egorear_b200 import synth
    from oracle import preprocess_ref as pr
    g = golden["preprocess"]
    for i, (H, W, oh, ow) in enumerate(PREPROCESS_CASES):
        img = synth.synth_images(1, H, W, seed=i)[0]
        f, u8 = pr.preprocess(img, oh, ow)
        assert u8.shape == (oh, ow, 3) and f.shape == (3, oh, ow) and f.dtype == np.float32
        assert np.array_equal(u8[:48, :48], g["u8_crop_%d" % i])
        assert np.array_equal(f[:, :48, :48].view(np.uint32), g["f32_crop_%d" % i].view(np.uint32))
        assert hashlib.sha256(np.ascontiguousarray(u8).tobytes()).digest() == g["u8_sha_%d" % i].tobytes()
        assert hashlib.sha256(np.ascontiguousarray(f).tobytes()).digest() == g["f32_sha_%d" % i].tobytes()
    # the inputs exercise both clips of the 8-bit passes
    u8 = pr.resize_bicubic_u8(synth.synth_images(1, 872, 872, seed=0)[0], 256, 256)
    assert (u8 == 0).any() and (u8 == 255).any()
    # identity size: Pillow skips the pass, the coefficient table reduces to the centre tap
    img = synth.synth_images(1, 64, 48, seed=9)[0]
    assert np.array_equal(pr.resize_bicubic_u8(img, 64, 48), img)
    try:
        from PIL import Image
    except ImportError:
        return
    img = synth.synth_images(1, 333, 517, seed=11)[0]
    want = np.asarray(Image.fromarray(img).resize([200, 120], Image.BICUBIC))
    assert np.array_equal(pr.resize_bicubic_u8(img, 120, 200), want)


def test_preprocess_oracle_vs_pil_random_geometries():
    """seeded sweep of input / output sizes (down- and up-scaling, odd sizes, 1-pixel axes) against PIL itself"""
    PIL = pytest.importorskip("PIL")
    from PIL import Image
    from egorear_b200 import synth
    from oracle import preprocess_ref as pr
    rng = np.random.default_rng(21)
    geos = [(int(rng.integers(1, 400)), int(rng.integers(1, 400)), int(rng.integers(1, 200)), int(rng.integers(1, 200))) for _ in range(14)]
    geos += [(872, 1, 256, 1), (1, 872, 1, 256), (2, 2, 255, 255), (1000, 3, 7, 300), (1000, 9, 500, 4), (1000, 10, 7, 300)]   # incl. Pillow's tall-strip rule
    for H, W, oh, ow in geos:
        img = synth.synth_images(1, H, W, seed=H * 1000 + W)[0]
        want = np.asarray(Image.fromarray(img).resize([ow, oh], Image.BICUBIC))
        assert np.array_equal(pr.resize_bicubic_u8(img, oh, ow), want), (H, W, oh, ow, PIL.__version__)
