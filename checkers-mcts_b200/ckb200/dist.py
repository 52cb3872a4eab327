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


def gather_packed(hdr, words, rank, world, device=None, dst=0):
    """Packed records (``RECORD_HDR_DTYPE`` headers + uint32 child words, see ckb200.records.pack) of every rank on
    ``dst``: exact-size point-to-point transfers (one grouped NCCL send/recv batch on GPUs, gloo send/recv in the CPU
    tests) straight between the buffers that are handed in -- torch tensors (device-resident: no host bounce) or numpy
    arrays.  No padding to the largest rank.  -> (hdr bytes tensor, words tensor, per-rank (n_hdr_bytes, n_words)) on
    ``dst``, None elsewhere."""
    import torch
    import torch.distributed as dist
    dev = torch.device("cpu") if device is None else torch.device(device)

    def as_tensor(a, np_dtype):
        if isinstance(a, torch.Tensor):
            return a.reshape(-1)
        flat = np.ascontiguousarray(a).reshape(-1)
        flat = flat.view(np_dtype) if flat.size else np.zeros(0, dtype=np_dtype)
        return torch.from_numpy(flat.copy()).to(dev)

    h = as_tensor(hdr, np.uint8)
    w = as_tensor(words, np.int32)
    mine = torch.tensor([h.numel(), w.numel()], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, mine)
    sizes = [(int(t[0].item()), int(t[1].item())) for t in sizes]
    if rank != dst:
        ops = []
        if h.numel():
            ops.append(dist.P2POp(dist.isend, h, dst))
        if w.numel():
            ops.append(dist.P2POp(dist.isend, w, dst))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return None
    H = torch.empty(sum(a for a, _ in sizes), dtype=torch.uint8, device=dev)
    W = torch.empty(sum(b for _, b in sizes), dtype=torch.int32, device=dev)
    ops, ho, wo = [], 0, 0
    for r, (a, b) in enumerate(sizes):
        if r == dst:
            H[ho:ho + a] = h
            W[wo:wo + b] = w
        else:
            if a:
                ops.append(dist.P2POp(dist.irecv, H[ho:ho + a], r))
            if b:
                ops.append(dist.P2POp(dist.irecv, W[wo:wo + b], r))
        ho += a
        wo += b
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return H, W, sizes


def gather_engine_records(engine, rank, world, device, dst=0, want_info=False, unpack=True):
    """The iteration-end gather (the reference's Pool.map return + merge_data, training_pipeline.py:277-284, 323-332):
    every rank packs the records of its finished games ON THE DEVICE into two torch buffers (ck_records_pack_device,
    ~60-75 bytes per record instead of 372), ``dst`` receives them with exact-size NCCL point-to-point transfers straight
    into device memory and decodes once on the host.  -> (RECORD_DTYPE array of all ranks ordered by rank on ``dst`` /
    None elsewhere, milliseconds), or with ``want_info`` a dict in place of the milliseconds: ``gather_ms`` = packing +
    NCCL transfers (everything up to "all records are in rank 0's HBM"), ``decode_ms`` = D2H copy + numpy decode into
    RECORD_DTYPE, ``ms`` their sum, ``bytes`` on the wire.  ``unpack=False`` returns (headers, words) instead of decoding."""
    import time
    import torch
    from . import lib as L
    from .lib_types import RECORD_HDR_DTYPE
    dev = torch.device(device)
    torch.cuda.synchronize(dev)
    t0 = time.time()
    n, nw = engine.records_packed_sizes()
    h = torch.empty(max(n, 1) * RECORD_HDR_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    w = torch.empty(max(nw, 1), dtype=torch.int32, device=dev)
    if n:
        engine.records_pack_device(h.data_ptr(), n, w.data_ptr(), max(nw, 1))
    h, w = h[:n * RECORD_HDR_DTYPE.itemsize], w[:nw]
    if world > 1:
        got = gather_packed(h, w, rank, world, device=dev, dst=dst)
    else:
        got = (h, w, [(h.numel(), w.numel())])
    torch.cuda.synchronize(dev)
    t1 = time.time()
    out = None
    nbytes = 0
    if got is not None:
        H, W, sizes = got
        nbytes = H.numel() + 4 * W.numel()
        hdr = H.cpu().numpy().view(RECORD_HDR_DTYPE)
        words = W.cpu().numpy().view(np.uint32)
        out = L.records_unpack(hdr, words) if unpack else (hdr, words)      # host-side C twin of records.unpack
    t2 = time.time()
    info = dict(ms=1000.0 * (t2 - t0), gather_ms=1000.0 * (t1 - t0), decode_ms=1000.0 * (t2 - t1), bytes=nbytes)
    return (out, info) if want_info else (out, info["ms"])


def warm_up_p2p(rank, world, device, dst=0):
    """NCCL opens its point-to-point channels lazily at the first send/recv between two ranks (hundreds of milliseconds
    for seven peers); one tiny exchange up front keeps that one-off cost out of a timed gather"""
    import torch
    if world > 1:
        gather_packed(torch.zeros(40, dtype=torch.uint8, device=device), torch.zeros(1, dtype=torch.int32, device=device),
                      rank, world, device=device, dst=dst)


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


def broadcast_strings(values, rank, world, src=0):
    """a short list of strings (file names; None allowed) from ``src`` to every rank"""
    import torch.distributed as dist
    if world == 1:
        return values
    box = [list(values) if rank == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]
