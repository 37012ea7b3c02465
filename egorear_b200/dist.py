"""Multi-GPU harness: frames shard across ranks (weights replicated), one all-gather of the packed joints.

SURVEY §8e: frames are independent (no cross-frame op on the hot path); the 4 views of a frame stay on one GPU.
One process per GPU (torchrun), `torch.distributed` over NCCL for the single collective: an all-gather of
672 B/frame ([B_local, 4*15*2 + 16*3] fp32).  No data-path collective elsewhere.  The helpers are backend-agnostic
so the sharding / gather ordering is covered by world_size-2 gloo tests on CPU.
"""
import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init(backend=None):
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    return rank, local_rank, world


def shard_range(n_frames, rank, world):
    """Contiguous frame block of `rank`: the first (n_frames % world) ranks get one extra frame."""
    base, rem = divmod(n_frames, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_rows(local, world=None):
    """all-gather of equally sized [B_local, C] row blocks -> [world*B_local, C] in rank order (frame order)."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return local
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())
    return out


def gather_ragged_rows(local, n_frames, rank, world):
    """all-gather when n_frames % world != 0: pad every block to the largest shard, gather, drop the padding."""
    if world == 1:
        return local
    sizes = [shard_range(n_frames, r, world) for r in range(world)]
    mx = max(e - s for s, e in sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    allr = gather_rows(pad, world).view(world, mx, *local.shape[1:])
    return torch.cat([allr[r, : e - s] for r, (s, e) in enumerate(sizes)], dim=0)


def max_over_ranks(value, device):
    """device-timed milliseconds -> max over ranks (the number a multi-GPU step is judged by)"""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def shutdown():
    if dist.is_initialized():
        dist.destroy_process_group()
