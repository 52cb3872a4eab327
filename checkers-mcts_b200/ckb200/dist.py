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


def rank_world():
    """(rank, world, local_rank) of the default torch.distributed group; (0, 1, None) outside one"""
    import os
    try:
        import torch.distributed as dist
    except ImportError:
        return 0, 1, None
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1, None
    return dist.get_rank(), dist.get_world_size(), int(os.environ.get("LOCAL_RANK", dist.get_rank()))


def broadcast_weights(blob, rank, world, device=None, src=0):
    """the float32 weight blob of ``src`` on every rank (iteration start); ``blob`` may be None
    on the other ranks"""
    import torch
    import torch.distributed as dist
    if world == 1:
        return blob
    dev = torch.device("cpu") if device is None else torch.device(device)
    n = torch.tensor([0 if blob is None else int(np.asarray(blob).size)], dtype=torch.int64, device=dev)
    dist.broadcast(n, src=src)
    buf = torch.empty(int(n.item()), dtype=torch.float32, device=dev)
    if rank == src:
        buf.copy_(torch.from_numpy(np.ascontiguousarray(blob, dtype=np.float32).reshape(-1)))
    dist.broadcast(buf, src=src)
    return buf.cpu().numpy()


def broadcast_int(value, rank, world, device=None, src=0):
    """one 63-bit integer (the run's seed) from ``src`` to every rank"""
    import torch
    import torch.distributed as dist
    if world == 1:
        return value
    dev = torch.device("cpu") if device is None else torch.device(device)
    t = torch.tensor([int(value) if rank == src else 0], dtype=torch.int64, device=dev)
    dist.broadcast(t, src=src)
    return int(t.item())
