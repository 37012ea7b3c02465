"""GPU: dataset image preprocessing through the C-ABI, bit-exact against PIL + torchvision vectors."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import PREPROCESS_CASES
from egorear_b200 import synth
from oracle import preprocess_ref as pr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from egorear_b200 import ops as o
    return o


def test_preprocess_golden_bit_exact(ops, golden):
    g = golden["preprocess"]
    for i, (H, W, oh, ow) in enumerate(PREPROCESS_CASES):
        img = synth.synth_images(1, H, W, seed=i)
        f, u8 = ops.preprocess_images(torch.from_numpy(img).cuda(), size=(ow, oh), return_resized=True)
        assert f.shape == (1, 3, oh, ow) and f.dtype == torch.float32 and u8.shape == (1, oh, ow, 3) and u8.dtype == torch.uint8
        f, u8 = f[0].cpu().numpy(), u8[0].cpu().numpy()
        assert np.array_equal(u8[:48, :48], g["u8_crop_%d" % i]), (H, W, oh, ow)
        assert np.array_equal(f[:, :48, :48].view(np.uint32), g["f32_crop_%d" % i].view(np.uint32)), (H, W, oh, ow)
        assert hashlib.sha256(np.ascontiguousarray(u8).tobytes()).digest() == g["u8_sha_%d" % i].tobytes()
        assert hashlib.sha256(np.ascontiguousarray(f).tobytes()).digest() == g["f32_sha_%d" % i].tobytes()


@pytest.mark.parametrize("shape,size", [((2, 4, 218, 218), (64, 64)), ((3, 37, 53), (96, 40)), ((1, 1, 16, 500), (256, 8)),
                                        ((5, 640, 7), (3, 256))])
def test_preprocess_oracle_shapes(ops, shape, size):
    """leading batch/view dims, odd sizes (unaligned rows), extreme aspect ratios; other mean / std"""
    lead, (H, W) = shape[:-2], shape[-2:]
    n = int(np.prod(lead))
    img = synth.synth_images(n, H, W, seed=H)
    mean, std = (0.5, 0.25, 0.125), (0.5, 0.3, 0.2)
    f, u8 = ops.preprocess_images(torch.from_numpy(img).cuda().view(*lead, H, W, 3), size=size, mean=mean, std=std,
                                  return_resized=True)
    assert f.shape == (*lead, 3, size[1], size[0])
    f, u8 = f.reshape(n, 3, size[1], size[0]).cpu().numpy(), u8.reshape(n, size[1], size[0], 3).cpu().numpy()
    for k in range(n):
        rf, ru = pr.preprocess(img[k], size[1], size[0], mean, std)
        assert np.array_equal(u8[k], ru)
        assert np.array_equal(f[k].view(np.uint32), rf.view(np.uint32))


def test_preprocess_properties_full_size(ops):
    """64 frames x 4 views of 872x872 (SURVEY §8d config sizes): constant images stay constant (the normalised
    coefficients sum to one), the result does not depend on the position in the batch, and a sample of images equals
    the oracle bit for bit."""
    img = torch.from_numpy(synth.synth_images(8, 872, 872, seed=3)).cuda()
    batch = img.repeat(32, 1, 1, 1).view(64, 4, 872, 872, 3)
    batch[5, 2] = 77
    f, u8 = ops.preprocess_images(batch, return_resized=True)
    assert f.shape == (64, 4, 3, 256, 256)
    assert bool((u8[5, 2] == 77).all())
    flat = u8.view(256, 256, 256, 3)
    for k in (0, 9, 100, 255):
        if k != 5 * 4 + 2:
            assert torch.equal(flat[k], flat[k % 8])
    for k in (0, 7):
        rf, ru = pr.preprocess(img[k].cpu().numpy())
        assert np.array_equal(flat[k].cpu().numpy(), ru)
        assert np.array_equal(f.view(256, 3, 256, 256)[k].cpu().numpy().view(np.uint32), rf.view(np.uint32))


def test_preprocess_argument_validation(ops):
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.preprocess_images(torch.zeros(1, 8, 8, 3, dtype=torch.uint8))
    with pytest.raises(RuntimeError, match="uint8"):
        ops.preprocess_images(torch.zeros(1, 8, 8, 3).cuda())
    with pytest.raises(RuntimeError, match="vertical-first"):         # Pillow's tall-strip rule is rejected, not mis-ordered
        ops.preprocess_images(torch.zeros(1, 1000, 3, 3, dtype=torch.uint8).cuda(), size=(300, 7))
    out = ops.preprocess_images(torch.zeros(0, 8, 8, 3, dtype=torch.uint8).cuda(), size=(4, 4))
    assert out.shape == (0, 3, 4, 4)
