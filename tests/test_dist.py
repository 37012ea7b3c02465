"""CPU, world_size 2, gloo: the frame sharding + single all-gather of the multi-GPU harness, and patch()."""
import os
import sys
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_frames, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from egorear_b200 import dist as egd
    r, lr, w = egd.init(backend="gloo")
    assert (r, w) == (rank, world)
    s, e = egd.shard_range(n_frames, rank, world)
    # every frame's packed row is a function of its global index: the gather must restore frame order
    rows = torch.arange(s, e, dtype=torch.float32)[:, None] * torch.ones(1, 168) + torch.arange(168)[None, :] / 1000.0
    if n_frames % world == 0:
        out = egd.gather_rows(rows, world)
    else:
        out = egd.gather_ragged_rows(rows, n_frames, rank, world)
    want = torch.arange(n_frames, dtype=torch.float32)[:, None] * torch.ones(1, 168) + torch.arange(168)[None, :] / 1000.0
    ok = torch.equal(out, want)
    t = egd.max_over_ranks(10.0 + rank, torch.device("cpu"))
    egd.barrier()
    q.put((rank, ok, t, (s, e)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [128, 37])
def test_shard_and_gather_world2(n_frames):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300) + n_frames % 7
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res)
    assert all(t == 11.0 for _, _, t, _ in res)              # max over ranks
    (s0, e0), (s1, e1) = res[0][3], res[1][3]
    assert s0 == 0 and e0 == s1 and e1 == n_frames and (e0 - s0) - (e1 - s1) in (0, 1)


def test_shard_range_covers_everything():
    from egorear_b200.dist import shard_range
    for n in (0, 1, 7, 64, 1000003):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_patch_rebinds_reference_names():
    """patch() swaps the hot-path names inside (stand-in) reference modules; absent modules are skipped"""
    import egorear_b200
    from egorear_b200 import modules as M, ops
    fake = types.ModuleType("pose_estimation.utils.loss")
    fake.get_max_preds = lambda *a, **k: None
    est = types.ModuleType("pose_estimation.models.estimator.egoposeformer_mvf_ex")
    for n in ("EgoPoseFormerMVFEX", "EgoPoseFormerPose3D", "EgoPoseFormerTransformerLayer", "DeformStereoAttn",
              "EgoformerSpatialMHA", "EgoPoseFormerHeatmapMVFEX"):
        setattr(est, n, object)
    saved = {k: sys.modules.get(k) for k in ("pose_estimation", "pose_estimation.utils", fake.__name__,
                                             "pose_estimation.models", "pose_estimation.models.estimator", est.__name__)}
    try:
        for k in saved:
            sys.modules[k] = types.ModuleType(k)
        sys.modules[fake.__name__], sys.modules[est.__name__] = fake, est
        done = egorear_b200.patch(precision="fp32", modules=[fake.__name__, est.__name__])
        assert fake.get_max_preds is ops.get_max_preds
        assert est.DeformStereoAttn is M.DeformStereoAttn
        assert issubclass(est.EgoPoseFormerPose3D, M.EgoPoseFormerPose3D) and est.EgoPoseFormerPose3D.__name__ == "EgoPoseFormerPose3D"
        assert set(done) == {fake.__name__, est.__name__}
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    with pytest.raises(ValueError):
        egorear_b200.patch(precision="fp8")
