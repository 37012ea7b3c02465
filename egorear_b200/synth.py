"""Deterministic synthetic weights and inputs (bench.py, smoke(), tests and tests/golden/make_golden.py).

There is no network for checkpoints or datasets, so every run uses random-init weights of the shipped
architectures and synthetic fisheye-shaped inputs.  Stock initialisation leaves the hot path degenerate
(SURVEY §7 trap 7: zero sampling_offsets / attention_weights, anchors never valid), so `fill_state_dict`
writes a NAME-SEEDED fill: each tensor depends only on its state_dict key and shape — not on module construction
order — which makes the reference modules (built in the survey container) and this package's mirror modules
carry bit-identical parameters without shipping 500 MB of weights.
"""
import math
import zlib

import numpy as np
import torch


def _gen(key, salt=0):
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) + 7919 * salt) & 0x7FFFFFFF)
    return g


def _ring_bias(n_heads=4, n_points=16):
    # stock MSDeformAttn offset bias (models/utils/deform_attn.py:69-81)
    th = torch.arange(n_heads, dtype=torch.float32) * (2.0 * math.pi / n_heads)
    g = torch.stack([th.cos(), th.sin()], -1)
    g = (g / g.abs().max(-1, keepdim=True)[0]).view(n_heads, 1, 1, 2).repeat(1, 1, n_points, 1)
    g = g * torch.arange(1, n_points + 1, dtype=torch.float32).view(1, 1, -1, 1)
    return g.reshape(-1)


def synth_tensor(key, shape, dtype=torch.float32):
    """Value of parameter `key` (hot-path keys only; see fill_state_dict)."""
    g = _gen(key)
    shape = tuple(shape)
    n = lambda std: torch.randn(shape, generator=g, dtype=torch.float32) * std
    u = lambda a: (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * a
    if key.endswith("sampling_offsets.weight"):
        t = n(0.05)
    elif key.endswith("sampling_offsets.bias"):
        t = _ring_bias() + n(0.25)
    elif key.endswith("attention_weights.weight"):
        t = n(0.2)
    elif key.endswith("attention_weights.bias"):
        t = n(0.5)
    elif key.endswith("frame_feat_multi_view_pos_embed"):
        t = n(0.02)
    elif key.endswith("joint_query_embed.weight"):
        t = n(1.0)
    elif key == "mlp_pred.2.bias" or key.endswith(".mlp_pred.2.bias"):
        # plausible device-relative pose (cm).  The fisheye cameras look along +z (theta = atan(-z/norm), rho(0) = image
        # circle radius), so body joints have z > 0: x,y ~ U(-50,50), z ~ U(-20,150) -> mixed in / out of FOV
        r = torch.rand((shape[0] // 3, 3), generator=g)
        t = torch.stack((r[:, 0] * 100 - 50, r[:, 1] * 100 - 50, r[:, 2] * 170 - 20), dim=1).reshape(shape)
    elif key.endswith(".weight") and len(shape) == 1:
        t = 1.0 + n(0.1)                           # every 1-D ".weight" on the hot path is a LayerNorm gain
    elif key.endswith(".weight") and len(shape) >= 2:
        fan_in = int(np.prod(shape[1:]))
        t = u(math.sqrt(6.0 / fan_in))            # keeps activations O(1) through ReLU stacks
    elif key.endswith(".bias"):
        t = u(0.1)
    else:
        t = n(0.02)
    return t.to(dtype)


def is_hot_path_key(key):
    return "heatmap_estimator_stereo" not in key and ".encoder." not in key


def fill_state_dict(module, hot_path_only=True):
    """Overwrite (in place) every floating-point hot-path entry of module.state_dict() with its name-seeded value."""
    sd = module.state_dict()
    with torch.no_grad():
        for k, v in sd.items():
            if not v.is_floating_point():
                continue
            if hot_path_only and not is_hot_path_key(k):
                continue
            v.copy_(synth_tensor(_strip_prefix(k), v.shape).to(v.device))
    return module


def _strip_prefix(k):
    # EgoPoseFormerMVFEX nests the two estimators; the fill is defined on the estimators' own key names
    for p in ("heatmap_estimator.", "pose3d_estimator."):
        if k.startswith(p):
            return k[len(p):]
    return k


def synth_state_dict(shapes):
    """{key: shape} -> {key: tensor} without building a module."""
    return {k: synth_tensor(k, s) for k, s in shapes.items()}


# ------------------------------------------------------------------------------------------------------
# inputs
# ------------------------------------------------------------------------------------------------------
def synth_features(B, V=4, seed=0, device="cpu"):
    """Backbone outputs: FPN map [B,V,128,64,64] (>= 0: the FPN ends in ReLU, resnet.py:113-119) and the
    stride-32 ResNet map [B,V,512,8,8] (>= 0: ends in a BasicBlock ReLU)."""
    feat = torch.empty((B, V, 128, 64, 64))
    bfb = torch.empty((B, V, 512, 8, 8))
    for b in range(B):                      # one stream per frame: frame b does not depend on the batch size
        g = torch.Generator(device="cpu")
        g.manual_seed((1000 + seed) * 100003 + b)
        feat[b] = torch.relu(torch.randn((V, 128, 64, 64), generator=g))
        bfb[b] = torch.relu(torch.randn((V, 512, 8, 8), generator=g))
    return feat.to(device), bfb.to(device)


def synth_keypoints(n_frames, V=4, J=16, seed=0, lo=-60.0, hi=932.0):
    """float64 [n_frames, V, J, 2] pixel coordinates, ~12 % off-image per axis."""
    return np.random.default_rng(seed).uniform(lo, hi, size=(n_frames, V, J, 2))


def synth_coord_trans_mat(B, seed=0, device="cpu"):
    """fp32 [B,4,4,4] rigid device->camera transforms in metres mirroring the syn rig."""
    noise = torch.empty((B, 4, 3))
    for b in range(B):                      # one stream per frame: frame b does not depend on the batch size
        g = torch.Generator(device="cpu")
        g.manual_seed((2000 + seed) * 100003 + b)
        noise[b] = torch.randn((4, 3), generator=g) * 0.005
    M = torch.zeros((B, 4, 4, 4), dtype=torch.float32)
    base_t = torch.tensor([[0.06, 0.0, 0.0], [-0.06, 0.0, 0.0], [-0.06, 0.37, 0.0], [0.06, 0.37, 0.0]])
    for v in range(4):
        R = torch.eye(3)
        if v >= 2:
            R = torch.diag(torch.tensor([-1.0, -1.0, 1.0]))     # 180 deg about z
        M[:, v, :3, :3] = R
        M[:, v, :3, 3] = base_t[v] + noise[:, v]
        M[:, v, 3, 3] = 1.0
    return M.to(device)


def synth_eval_poses(B, J=16, seed=0):
    """(pred, gt) float32 [B,J,3] in cm for the eval-metric tests/bench: gt inside the body volume, pred = a random
    similarity of gt plus joint noise whose scale varies per sample, so that PCK / AUC land strictly between 0 and 1.
    Samples 0-3 are edge cases: exact match, point reflection, coplanar prediction, pure similarity."""
    rng = np.random.default_rng(4000 + seed)
    gt = rng.uniform(-60, 60, (B, J, 3))
    gt[..., 2] -= 80.0
    noise = rng.normal(0.0, 1.0, (B, J, 3)) * rng.uniform(0.2, 8.0, (B, 1, 1))
    ang = rng.normal(0.0, 0.08, (B, 3))
    ca, sa, cb, sb, cc, sc = (f(ang[:, i]) for i in range(3) for f in (np.cos, np.sin))
    zero, one = np.zeros(B), np.ones(B)
    Rx = np.stack([one, zero, zero, zero, ca, -sa, zero, sa, ca], -1).reshape(B, 3, 3)
    Ry = np.stack([cb, zero, sb, zero, one, zero, -sb, zero, cb], -1).reshape(B, 3, 3)
    Rz = np.stack([cc, -sc, zero, sc, cc, zero, zero, zero, one], -1).reshape(B, 3, 3)
    R = Rz @ Ry @ Rx
    mu = gt.mean(1, keepdims=True)
    pred = np.einsum("bij,bkj->bki", R, gt - mu) * rng.uniform(0.9, 1.1, (B, 1, 1)) + mu + rng.normal(0, 2.0, (B, 1, 3)) + noise
    if B > 0:
        pred[0] = gt[0]
    if B > 1:
        pred[1] = -gt[1]
    if B > 2:
        pred[2, :, 2] = 0.0
    if B > 3:
        pred[3] = gt[3] * 2.5 + 7.0
    return pred.astype(np.float32), gt.astype(np.float32)


def synth_eval_heatmaps(B, V=2, C=15, hs=64, seed=0):
    """(pred, gt) float32 [B,V,C,hs,hs]: gt = unit-peak Gaussians (sigma 1) at random cells, ~15 % of them off the map
    (all-zero target -> invalid under the wrappers' threshold 1.0) and ~10 % attenuated below 1.0; pred = gt moved by
    up to two cells, rescaled, plus dense noise (so arg-max positions differ and negative values occur)."""
    rng = np.random.default_rng(5000 + seed)
    yy, xx = np.meshgrid(np.arange(hs), np.arange(hs), indexing="ij")
    n = B * V * C
    cx = rng.integers(-8, hs + 8, n)
    cy = rng.integers(-8, hs + 8, n)
    amp = np.where(rng.uniform(size=n) < 0.1, 0.7, 1.0)
    sx = cx + rng.integers(-2, 3, n)
    sy = cy + rng.integers(-2, 3, n)
    gt = np.zeros((n, hs, hs), np.float32)
    pred = np.zeros((n, hs, hs), np.float32)
    for i in range(n):
        if 0 <= cx[i] < hs and 0 <= cy[i] < hs:
            gt[i] = (amp[i] * np.exp(-((xx - cx[i]) ** 2 + (yy - cy[i]) ** 2) / 2.0)).astype(np.float32)
            gt[i][gt[i] < 1e-4] = 0.0
        pred[i] = (0.8 * np.exp(-((xx - sx[i]) ** 2 + (yy - sy[i]) ** 2) / 2.0)).astype(np.float32)
    pred += rng.normal(0.0, 0.02, pred.shape).astype(np.float32)
    return pred.reshape(B, V, C, hs, hs), gt.reshape(B, V, C, hs, hs)


def synth_images(N, H=872, W=872, seed=0):
    """uint8 [N,H,W,3] stand-ins for decoded camera frames: smooth gradients + texture noise + saturated / black blocks
    and one-pixel lines (the negative lobes of the bicubic kernel over- and undershoot there, exercising the 8-bit
    clipping of both resampling passes)."""
    rng = np.random.default_rng(6000 + seed)
    yy, xx = np.meshgrid(np.linspace(0, 1, H), np.linspace(0, 1, W), indexing="ij")
    out = np.empty((N, H, W, 3), np.uint8)
    for n in range(N):
        ph = rng.uniform(0, 6.28, 3)
        fr = rng.uniform(2, 9, 3)
        img = np.stack([127.5 + 127.5 * np.sin(fr[c] * (xx * (c + 1) + yy * (3 - c)) + ph[c]) for c in range(3)], -1)
        img += rng.normal(0, 25, (H, W, 3))
        img = np.clip(img, 0, 255)
        for _ in range(6):
            y0, x0 = rng.integers(0, max(H - 8, 1)), rng.integers(0, max(W - 8, 1))
            h, w = rng.integers(2, max(H // 6, 3)), rng.integers(2, max(W // 6, 3))
            img[y0:y0 + h, x0:x0 + w] = rng.choice([0.0, 255.0], 3)
        img[rng.integers(0, H)] = 255.0
        img[:, rng.integers(0, W)] = 0.0
        out[n] = img.astype(np.uint8)
    return out
