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
        T = 7 + rank                                  # ranks hold different row widths
        s, e = gdist.shard_bounds(batch, rank, world)
        ids = torch.zeros((e - s, T), dtype=torch.int32)
        frames = torch.zeros((e - s, T), dtype=torch.int32)
        counts = torch.zeros((e - s,), dtype=torch.int32)
        for i, u in enumerate(range(s, e)):           # utterance u emits (u % 5) tokens u, u+1, ...
            n = u % 5
            ids[i, :n] = torch.arange(u, u + n, dtype=torch.int32)
            frames[i, :n] = torch.arange(n, dtype=torch.int32) + 1
            counts[i] = n
        gi, gf, gc = gdist.gather_hypotheses(ids, frames, counts, batch)
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


def test_gather_hypotheses_world2_uneven():
    world, batch = 2, 5
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, batch, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}
