"""CPU test of the N > 1 host logic with two gloo processes: game sharding and the iteration-end
record gather (no GPU, no kernels)."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

from helpers import PKG, ROOT


def _worker(rank, world, port, q):
    for p in (ROOT, PKG):
        sys.path.insert(0, p)
    import torch.distributed as dist
    from ckb200 import dist as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dt = np.dtype([("game", "<i4"), ("ply", "<i4"), ("q", "<f4"), ("visits", "<u4", (5,))])
    base, stride, n_local = D.shard(11, rank, world)
    recs = np.zeros(n_local * 3, dtype=dt)              # three records per local game
    for i in range(n_local):
        for k in range(3):
            recs[i * 3 + k] = (base + i * stride, k, rank + 0.5, [rank, i, k, 7, 9])
    out = D.gather_records(recs, rank, world)
    # packed records: exact-size point-to-point transfers, ragged sizes (rank 1 sends nothing at all in the second call)
    from ckb200 import records as R
    from ckb200.lib_types import RECORD_DTYPE
    full = np.zeros(n_local * 2, dtype=RECORD_DTYPE)
    for i in range(len(full)):
        full[i]["game"], full[i]["ply"], full[i]["n_children"] = base + (i // 2) * stride, i % 2, 3
        full[i]["action"][:3], full[i]["visits"][:3] = (149, 151, 300 + rank), (5 + i, 7, 9)
        full[i]["pos"] = (0xFFF, 0xFFF00000, 0, rank)
    full["mask"] = R.unpack(*R.pack(full))["mask"]
    for send in (full, full[:0] if rank == 1 else full):
        h, w = R.pack(send)
        got = D.gather_packed(h, w, rank, world)
        if rank == 0:
            H, W, sizes = got
            back = R.unpack(H.numpy().view(h.dtype), W.numpy().view(np.uint32))
            assert back[:len(send)].tobytes() == send.tobytes() and sizes[0] == (h.nbytes, len(w))
            q.put(("packed", len(back)))
        else:
            assert got is None
    # iteration start: the weights and the seed of rank 0 reach every rank
    assert D.rank_world()[:2] == (rank, world)
    blob = D.broadcast_weights(np.arange(1000, dtype=np.float32) * 0.5 if rank == 0 else None, rank, world)
    assert blob.dtype == np.float32 and (blob == np.arange(1000, dtype=np.float32) * 0.5).all()
    assert D.broadcast_int(0x1234567890ABCDEF >> 1 if rank == 0 else 7, rank, world) == 0x1234567890ABCDEF >> 1
    if rank == 0:
        q.put((n_local, out.tobytes(), out.dtype.descr))
    else:
        assert out is None
        q.put((n_local, None, None))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_gather_two_ranks():
    from ckb200 import dist as D
    # every game has exactly one owner, ids are base + i*stride
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            base, stride, n = D.shard(4096 + 3, r, world)
            ids = [base + i * stride for i in range(n)]
            assert all(D.owner(g, world) == r for g in ids)
            seen += ids
        assert sorted(seen) == list(range(4096 + 3))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(4)]
    packed = sorted(r[1] for r in res if r[0] == "packed")
    assert packed == [12, 22]                         # rank 0: 6 games x 2 records; rank 1: 5 games x 2 (then nothing)
    res = [r for r in res if r[0] != "packed"]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n_locals = sorted(r[0] for r in res)
    assert n_locals == [5, 6]
    blob = [r for r in res if r[1] is not None][0]
    dt = np.dtype([("game", "<i4"), ("ply", "<i4"), ("q", "<f4"), ("visits", "<u4", (5,))])
    out = np.frombuffer(blob[1], dtype=dt)
    assert len(out) == 11 * 3
    assert sorted(set(out["game"].tolist())) == list(range(11))
    assert (out["q"][out["game"] % 2 == 1] == 1.5).all() and (out["q"][out["game"] % 2 == 0] == 0.5).all()
    assert (out["visits"][:, 3] == 7).all()
