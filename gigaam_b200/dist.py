"""Multi-GPU layout of the path: utterances are independent, so a batch is sharded across ranks with no
data-path collective; the only exchange is ONE all-gather of the device-resident hypotheses when the batch was actually
split.  One process per GPU (torchrun).

On GPUs the gather is the library's own `gam_gather_hyps` (include/gigaam_b200.h): ids / frames / counts of a shard live
in one packed int32 buffer that a single ncclAllGather (NVLink 5 / NVSwitch) replicates to every rank, stream-ordered
behind the greedy kernels and capturable in the same CUDA graph -- no shape exchange, no host synchronisation.
`torch.distributed` only bootstraps it (ships the 128-byte NCCL id).  The CPU tests (gloo) drive the same packing
through `torch.distributed.all_gather_into_tensor`."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

Tensor = torch.Tensor


def shard_bounds(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of `batch` utterances: the first (batch % world) ranks get one extra."""
    base, extra = divmod(batch, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batch(wav: Tensor, lengths: Tensor, rank: int, world: int) -> Tuple[Tensor, Tensor]:
    s, e = shard_bounds(wav.shape[0], rank, world)
    return wav[s:e], lengths[s:e]


def shard_rows(batch: int, world: int) -> int:
    """Rows every rank's packed buffer carries: the largest shard (shorter shards pad with counts = 0)."""
    return (batch + world - 1) // world


def pack_hypotheses(ids: Optional[Tensor], frames: Optional[Tensor], counts: Optional[Tensor], rows: int, width: int,
                    device) -> Tensor:
    """[ids rows x width | frames rows x width | counts rows] int32; missing rows (short or empty shard) have counts 0."""
    buf = torch.zeros(2 * rows * width + rows, dtype=torch.int32, device=device)
    if ids is not None and ids.shape[0] > 0:
        b, w = ids.shape
        buf[: rows * width].view(rows, width)[:b, :w] = ids
        buf[rows * width: 2 * rows * width].view(rows, width)[:b, :w] = frames
        buf[2 * rows * width: 2 * rows * width + b] = counts
    return buf


def unpack_gathered(gathered: Tensor, batch: int, world: int, rows: int, width: int) -> Tuple[Tensor, Tensor, Tensor]:
    """[world, 2 rows width + rows] -> (ids, frames [batch, width], counts [batch]) in global-batch order."""
    g = gathered.view(world, -1)
    ids = g[:, : rows * width].reshape(world, rows, width)
    frames = g[:, rows * width: 2 * rows * width].reshape(world, rows, width)
    counts = g[:, 2 * rows * width:].reshape(world, rows)
    if batch == world * rows:        # even split: pure reshapes (no host-built index: capturable in a CUDA graph)
        return ids.reshape(batch, width), frames.reshape(batch, width), counts.reshape(batch)
    keep = torch.cat([torch.arange(r * rows, r * rows + (shard_bounds(batch, r, world)[1] - shard_bounds(batch, r, world)[0]))
                      for r in range(world)]).to(g.device)
    return (ids.reshape(world * rows, width)[keep].contiguous(), frames.reshape(world * rows, width)[keep].contiguous(),
            counts.reshape(world * rows)[keep].contiguous())


class HypothesisGather:
    """The library's NCCL all-gather bound to one engine (one per process / GPU).  `torch.distributed` (any backend that
    can broadcast 128 bytes between the ranks) is used once, to ship rank 0's NCCL id."""

    def __init__(self, engine, group=None):
        self.engine = engine
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        lib = engine.lib
        uid = (C.c_uint8 * 128)()
        if self.rank == 0 and lib.gam_comm_unique_id(uid) != 0:
            raise RuntimeError("NCCL is not loadable in this process: gam_comm_unique_id failed")
        t = torch.tensor(list(uid), dtype=torch.uint8, device=engine.device if dist.get_backend(group) == "nccl" else "cpu")
        dist.broadcast(t, src=0, group=group)
        uid = (C.c_uint8 * 128)(*t.cpu().tolist())
        with torch.cuda.device(engine.device):
            rc = lib.gam_comm_init(engine.handle, uid, self.rank, self.world)
        from . import _lib
        _lib.check(lib, engine.handle, rc, "gam_comm_init")

    def all_gather(self, packed: Tensor) -> Tensor:
        """packed: int32 device buffer (same length on every rank) -> [world, n] on every rank, on the current stream."""
        eng = self.engine
        out = torch.empty((self.world, packed.numel()), dtype=torch.int32, device=eng.device)
        with torch.cuda.device(eng.device):
            rc = eng.lib.gam_gather_hyps(eng.handle, packed.data_ptr(), packed.numel(), out.data_ptr(), eng._stream())
        from . import _lib
        _lib.check(eng.lib, eng.handle, rc, "gam_gather_hyps")
        return out


def gather_hypotheses(ids: Optional[Tensor], frames: Optional[Tensor], counts: Optional[Tensor], batch: int, width: int,
                      device, gather: Optional[HypothesisGather] = None, group=None) -> Tuple[Tensor, Tensor, Tensor]:
    """One packed all-gather of a shard's hypotheses into global-batch order.  `ids` etc. may be None (empty shard).
    `width` = row pitch of the id / frame matrices, identical on all ranks (T' for CTC, T' x max_symbols for RNN-T)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return ids, frames, counts
    rows = shard_rows(batch, world)
    packed = pack_hypotheses(ids, frames, counts, rows, width, device)
    if gather is not None:
        out = gather.all_gather(packed)
    else:
        out = torch.empty((world, packed.numel()), dtype=torch.int32, device=device)
        dist.all_gather_into_tensor(out.view(-1), packed, group=group)
    return unpack_gathered(out, batch, world, rows, width)


def transcribe_sharded(model, wav: Tensor, lengths: Tensor, gather: Optional[HypothesisGather] = None
                       ) -> List[Tuple[str, List[int], List[int]]]:
    """Every rank calls this with the SAME global batch; each encodes + decodes its shard on its own GPU and all ranks
    return the full list of hypotheses (detokenised on every rank from the gathered ids).  With more ranks than
    utterances some shards are empty: those ranks skip the compute and still join the gather."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    w, l = shard_batch(wav, lengths, rank, world)
    dev = model._device
    eng = model._get_engine()
    T = eng.encoded_frames(eng.logmel_frames(wav.shape[1]))
    width = T if eng.head_type == 1 else T * eng.max_symbols
    ids = frames = counts = None
    if w.shape[0] > 0:
        if w.shape[0] == 1:
            # a lone utterance gets no attention mask (gigaam/encoder.py:616-626: masks exist only for B > 1), so its padding
            # must not reach the encoder: trim it, as a caller of the unsharded reference would never pad a batch of one
            w = w[:, : int(l[0])]
        enc, enc_len = model(w.to(dev), l.to(dev))
        ids, frames, counts = model.decoding.decode_device(model.head, enc, enc_len)
    if gather is None and world > 1 and dev.type == "cuda" and dist.get_backend() == "nccl":
        gather = model.__dict__.get("_hyp_gather")
        if gather is None or gather.engine is not eng:
            gather = HypothesisGather(eng)
            model.__dict__["_hyp_gather"] = gather
    ids, frames, counts = gather_hypotheses(ids, frames, counts, wav.shape[0], width, dev, gather)
    return model.decoding.to_hypotheses(ids.cpu(), frames.cpu(), counts.cpu())
