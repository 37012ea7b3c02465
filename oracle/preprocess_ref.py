"""ORACLE (test infrastructure, never on the product path): numpy restatement of the reference's image preprocessing.

SURVEY §8f row 3.  Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may import this file.

Reference (pose_estimation/datasets/ego4view_syn/ego4view_syn_heatmap_mvf.py and the other five dataset classes, e.g.
datasets/ego4view_rw/ego4view_rw_heatmap_mvf.py:40-41, 96-99):

    img = Image.open(path).convert("RGB")
    x = transform(img.resize([256, 256], Image.BICUBIC)).float().numpy()
    transform = Compose([ToTensor(), Normalize((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))])

The arithmetic of `Image.resize(..., BICUBIC)` lives in a third-party dependency that is not under /root/reference:
Pillow (unpinned in README.md:133; 12.2.0 in this image), src/libImaging/Resample.c.  Its published algorithm, restated
here: two separable passes (horizontal first, except for Pillow's tall-strip rule, see vertical_first) over 8-bit data
with an 8-bit intermediate image; per output index the
bicubic kernel (a = -0.5, support 2) is stretched by the down-scaling factor (antialiasing), sampled at the input pixel
centres inside [center - support, center + support], normalised in double and converted to fixed point with 22
fractional bits; a pixel is clip8((2^21 + sum k_i * p_i) >> 22).  ToTensor divides by 255 in float32, Normalize subtracts
the float32 mean and divides by the float32 std.

Parity pinned: tests/test_oracle.py checks this file bit-for-bit against tests/golden/preprocess.npz, which
tests/golden/make_golden.py produced with the real PIL + torchvision transforms in the build container (and against
PIL itself whenever PIL is importable).
"""
import numpy as np

PRECISION_BITS = 32 - 8 - 2
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def _bicubic(x):
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the full-image box.
    -> bounds int32 [out, 2] (xmin, count), kk int32 [out, ksize] (22-bit fixed point), ksize"""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = np.array([_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)], np.float64)
        ww = 0.0
        for v in w:                       # sequential double sum, like the C loop
            ww += v
        if ww != 0.0:
            w = w / ww
        fixed = np.where(w < 0, -0.5 + w * (1 << PRECISION_BITS), 0.5 + w * (1 << PRECISION_BITS))
        kk[xx, :xmax] = np.trunc(fixed).astype(np.int32)          # C (int) cast truncates toward zero
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def _clip8(v):
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def _pass_horizontal(img, out_w):
    H, W, C = img.shape
    bounds, kk, _ = precompute_coeffs(W, out_w)
    src = img.astype(np.int64)
    out = np.empty((H, out_w, C), np.uint8)
    for xx in range(out_w):
        x0, n = bounds[xx]
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(src[:, x0:x0 + n, :], kk[xx, :n].astype(np.int64), axes=([1], [0]))
        out[:, xx, :] = _clip8(acc)
    return out


def _pass_vertical(img, out_h):
    H, W, C = img.shape
    bounds, kk, _ = precompute_coeffs(H, out_h)
    src = img.astype(np.int64)
    out = np.empty((out_h, W, C), np.uint8)
    for yy in range(out_h):
        y0, n = bounds[yy]
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(kk[yy, :n].astype(np.int64), src[y0:y0 + n], axes=([0], [0]))
        out[yy] = _clip8(acc)
    return out


def vertical_first(in_h, in_w, out_h):
    """PIL/Image.py (Pillow >= 11; 12.2.0 pinned here), Image.resize: very tall strips that shrink vertically are
    resampled vertically first (`if self.size[1] > self.size[0] * 100 and size[1] < self.size[1]`).  Never true for
    camera frames; the CUDA path rejects such inputs instead of silently using the other order."""
    return in_h > in_w * 100 and out_h < in_h


def resize_bicubic_u8(img, out_h, out_w):
    """img uint8 [H, W, C] -> uint8 [out_h, out_w, C], bit-exact restatement of Image.resize((out_w, out_h), BICUBIC):
    horizontal pass, 8-bit intermediate, vertical pass (a pass whose size does not change is skipped)."""
    H, W, C = img.shape
    if vertical_first(H, W, out_h):
        img = _pass_vertical(img, out_h)
        return _pass_horizontal(img, out_w) if out_w != W else img
    if out_w != W:
        img = _pass_horizontal(img, out_w)
    if out_h != H:
        img = _pass_vertical(img, out_h)
    return img


def to_tensor_normalize(u8, mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """ToTensor + Normalize on uint8 [H, W, 3] -> float32 [3, H, W]: (x / 255 - mean) / std, every step in float32."""
    x = u8.astype(np.float32).transpose(2, 0, 1) / np.float32(255)
    m = np.asarray(mean, np.float32)[:, None, None]
    s = np.asarray(std, np.float32)[:, None, None]
    return (x - m) / s


def preprocess(img, out_h=256, out_w=256, mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """one decoded RGB image uint8 [H, W, 3] -> (float32 [3, out_h, out_w], the resized uint8 image)"""
    r = resize_bicubic_u8(img, out_h, out_w)
    return to_tensor_normalize(r, mean, std), r
