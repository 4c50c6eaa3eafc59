"""world_size-2 gloo test of the multi-GPU host logic: utterance sharding and the hypothesis all-gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gigaam_b200 import dist as gdist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        T = 9                                         # row pitch: a function of the (shared) padded input length
        s, e = gdist.shard_bounds(batch, rank, world)
        ids = torch.zeros((e - s, T), dtype=torch.int32)
        frames = torch.zeros((e - s, T), dtype=torch.int32)
        counts = torch.zeros((e - s,), dtype=torch.int32)
        for i, u in enumerate(range(s, e)):           # utterance u emits (u % 5) tokens u, u+1, ...
            n = u % 5
            ids[i, :n] = torch.arange(u, u + n, dtype=torch.int32)
            frames[i, :n] = torch.arange(n, dtype=torch.int32) + 1
            counts[i] = n
        if e == s:                                    # empty shard (more ranks than utterances): still joins the gather
            ids = frames = counts = None
        gi, gf, gc = gdist.gather_hypotheses(ids, frames, counts, batch, T, torch.device("cpu"))
        ok = gi.shape[0] == batch
        for u in range(batch):
            n = u % 5
            ok &= int(gc[u]) == n and gi[u, :n].tolist() == list(range(u, u + n)) and gf[u, :n].tolist() == list(range(1, n + 1))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_the_batch():
    for batch in (1, 2, 5, 64, 257):
        for world in (1, 2, 3, 8):
            spans = [gdist.shard_bounds(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_gather_hypotheses_world2_uneven_and_empty_shard():
    """One packed all-gather per call: an uneven split (5 utterances on 2 ranks) and a batch smaller than the world
    (1 utterance on 2 ranks: rank 1's shard is empty and must neither hang nor corrupt the result)."""
    world = 2
    for batch in (5, 1):
        port = _free_port()
        with mp.Manager() as mgr:
            ret = mgr.dict()
            mp.spawn(_worker, args=(world, port, batch, ret), nprocs=world, join=True)
            assert dict(ret) == {0: True, 1: True}, batch


def test_pack_unpack_roundtrip():
    rows, width, world, batch = 3, 4, 3, 7          # shards of 3, 2, 2
    bufs = []
    for r in range(world):
        s, e = gdist.shard_bounds(batch, r, world)
        ids = torch.arange(s * width, e * width, dtype=torch.int32).view(e - s, width)
        bufs.append(gdist.pack_hypotheses(ids, ids + 1000, torch.arange(s, e, dtype=torch.int32), rows, width, "cpu"))
    gi, gf, gc = gdist.unpack_gathered(torch.stack(bufs), batch, world, rows, width)
    assert gi.flatten().tolist() == list(range(batch * width)) and torch.equal(gf, gi + 1000) and gc.tolist() == list(range(batch))
