"""Dataset-side drop-in (egorear_b200/datasets.py): the reference's OWN dataset classes, with the resize
+ normalise and the .npy heatmap read substituted at their call sites, must hand the wrappers bit-identical batches.

A one-frame Ego4View-syn tree is written to tmp_path (4 fisheye frames, the frame JSON with pts2d, and the fisheye_hm
.npy files rendered by the reference's generate_target, as generate_heatmap.py would).  Needs the reference (checkout or
its file copy oracle/_ref) — skipped otherwise."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import ref_import

needs_ref = pytest.mark.skipif(not ref_import.available(), reason="reference checkout / oracle/_ref not present")
CAMS = ["camera_front_left", "camera_front_right", "camera_back_left", "camera_back_right"]


def make_tree(root, n_frames=2, size=872):
    from PIL import Image
    from egorear_b200 import datasets as egd, synth
    gen = ref_import.import_functions()["generate_target"]
    seq = os.path.join(root, "rp_test", "seq_000")
    kp = synth.synth_keypoints(n_frames, 4, 16, seed=3)
    imgs = synth.synth_images(n_frames * 4, size, size, seed=1).reshape(n_frames, 4, size, size, 3)
    for f in range(n_frames):
        name = "%06d" % f
        joints = {j: {"%s_pts2d" % c: kp[f, v, i].tolist() for v, c in enumerate(CAMS)} for i, j in enumerate(egd.JOINTS)}
        os.makedirs(os.path.join(seq, "json_smplx_gendered"), exist_ok=True)
        with open(os.path.join(seq, "json_smplx_gendered", name + ".json"), "w") as fh:
            json.dump({"joints": joints}, fh)
        for v, c in enumerate(CAMS):
            os.makedirs(os.path.join(seq, "fisheye_rgb", c), exist_ok=True)
            os.makedirs(os.path.join(seq, "fisheye_hm", c), exist_ok=True)
            Image.fromarray(imgs[f, v]).save(os.path.join(seq, "fisheye_rgb", c, name + ".jpg"), quality=92)
            np.save(os.path.join(seq, "fisheye_hm", c, name + ".npy"), gen(kp[f, v], 872, 64, 16, 1.0))   # generate_heatmap.py:62-79
    info = os.path.join(root, "info.txt")
    with open(info, "w") as fh:
        fh.write("rp_test\n")
    return info, kp


def load_both(tmp_path):
    """reference samples first (unpatched module), then the same dataset object after patch_datasets()"""
    import importlib
    from egorear_b200 import datasets as egd
    ref_import.install()
    info, kp = make_tree(str(tmp_path))
    name = "pose_estimation.datasets.ego4view_syn.ego4view_syn_heatmap_mvf"
    mod = importlib.import_module(name)
    if getattr(mod, "_egr_patched", False):
        mod = importlib.reload(mod)
    ds = mod.Ego4ViewSynHeatmapMVFDataset(str(tmp_path), info, camera_pos="all")
    order = sorted(range(len(ds)), key=lambda i: ds.frame_dataset[i])
    ref = [ds[i] for i in order]
    assert egd.patch_datasets(modules=[name]) == [name]
    ds2 = mod.Ego4ViewSynHeatmapMVFDataset(str(tmp_path), info, camera_pos="all")
    raw = [ds2[i] for i in sorted(range(len(ds2)), key=lambda i: ds2.frame_dataset[i])]
    importlib.reload(mod)                      # leave the module as found for other tests
    return ref, raw, kp


@needs_ref
def test_patched_dataset_returns_raw_frames_and_keypoints(tmp_path):
    ref, raw, kp = load_both(tmp_path)
    assert len(ref) == len(raw) == 2
    for f, (r, w) in enumerate(zip(ref, raw)):
        assert r["frame_path"] == w["frame_path"]
        assert r["img"].shape == (4, 3, 256, 256) and r["img"].dtype == torch.float32
        assert w["img"].shape == (4, 872, 872, 3) and w["img"].dtype == torch.uint8
        assert r["gt_heatmap"].shape == (4, 15, 64, 64)
        assert w["gt_heatmap"].shape == (4, 15, 1, 2) and w["gt_heatmap"].dtype == torch.float64
        assert np.array_equal(w["gt_heatmap"].numpy()[:, :, 0], kp[f][:, 1:])          # joint 0 dropped like the reference
    # the DataLoader's default collate works on the raw fields
    batch = torch.utils.data.default_collate(raw)
    assert batch["img"].shape == (2, 4, 872, 872, 3) and batch["gt_heatmap"].shape == (2, 4, 15, 1, 2)


@needs_ref
@pytest.mark.gpu
def test_gpu_batch_transform_equals_reference_dataset(tmp_path):
    from egorear_b200 import datasets as egd
    ref, raw, _ = load_both(tmp_path)
    want = torch.utils.data.default_collate(ref)
    batch = torch.utils.data.default_collate(raw)
    batch = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
    got = egd.gpu_batch_transform(batch)
    assert got["img"].shape == want["img"].shape and got["gt_heatmap"].shape == want["gt_heatmap"].shape
    assert torch.equal(got["img"].cpu().view(torch.int32), want["img"].view(torch.int32))                  # PIL + torchvision, bit for bit
    assert torch.equal(got["gt_heatmap"].cpu().view(torch.int32), want["gt_heatmap"].view(torch.int32))    # generate_target, bit for bit
    again = egd.gpu_batch_transform(got)          # idempotent
    assert again["img"] is got["img"]
