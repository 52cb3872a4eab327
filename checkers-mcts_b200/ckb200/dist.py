"""Multi-GPU plumbing: games shard by index across ranks, no collective on the data
path; one gather pools the finished self-play records on rank 0 at the end of an iteration
(the reference's Pool.map return + merge_data, training_pipeline.py:277-284, 323-332)."""
import numpy as np


def shard(n_games_total, rank, world):
    """game g lives on rank g mod world.  -> (game_id_base, game_id_stride, n_local)"""
    n_local = (n_games_total - rank + world - 1) // world if n_games_total > rank else 0
    return rank, world, n_local


def owner(game_id, world):
    return game_id % world


def gather_records(records, rank, world, device=None, dst=0):
    """records: structured numpy array (any dtype) -> concatenated array of all ranks on ``dst``
    (None elsewhere), ordered by rank.  Uses the default torch.distributed group (NCCL on GPUs,
    gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return records
    dev = torch.device("cpu") if device is None else torch.device(device)
    payload = torch.from_numpy(np.ascontiguousarray(records).view(np.uint8).reshape(-1).copy()).to(dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([payload.numel()], dtype=torch.int64, device=dev))
    sizes = [int(s.item()) for s in sizes]
    mx = max(max(sizes), 1)
    padded = torch.zeros(mx, dtype=torch.uint8, device=dev)
    padded[:payload.numel()] = payload
    out = [torch.empty(mx, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == dst else None
    dist.gather(padded, out, dst=dst)
    if rank != dst:
        return None
    parts = [out[r][:sizes[r]].cpu().numpy().view(records.dtype) for r in range(world)]
    return np.concatenate(parts) if parts else records[:0]
