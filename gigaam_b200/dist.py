"""Multi-GPU layout of the path: utterances are independent, so a batch is sharded across ranks
with no data-path collective; the only exchange is one all-gather of the device-resident hypotheses
(ids / frames / counts) when the batch was actually split.  One process per GPU (torchrun), backend NCCL on
GPUs (NVLink 5 / NVSwitch) and gloo in the CPU tests."""
from __future__ import annotations

from typing import List, Tuple

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


def gather_hypotheses(ids: Tensor, frames: Tensor, counts: Tensor, batch: int, group=None
                      ) -> Tuple[Tensor, Tensor, Tensor]:
    """All-gather per-rank [B_local, W] id / frame matrices and [B_local] counts into global-batch order.
    Ranks may hold different B_local (uneven split) and different widths W: rows are padded to the maxima."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return ids, frames, counts
    dev = ids.device
    shape = torch.tensor([ids.shape[0], ids.shape[1]], dtype=torch.int64, device=dev)
    shapes = [torch.empty_like(shape) for _ in range(world)]
    dist.all_gather(shapes, shape, group=group)
    max_b = max(int(s[0]) for s in shapes)
    max_w = max(int(s[1]) for s in shapes)
    payload = torch.zeros((max_b, 2 * max_w + 1), dtype=torch.int32, device=dev)
    b, w = ids.shape
    payload[:b, :w] = ids
    payload[:b, max_w:max_w + w] = frames
    payload[:b, 2 * max_w] = counts
    out = torch.empty((world * max_b, 2 * max_w + 1), dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(out, payload, group=group)
    rows = torch.cat([torch.arange(r * max_b, r * max_b + int(shapes[r][0]), device=dev) for r in range(world)])
    out = out[rows]
    assert out.shape[0] == batch, (out.shape, batch)
    return out[:, :max_w].contiguous(), out[:, max_w:2 * max_w].contiguous(), out[:, 2 * max_w].contiguous()


def transcribe_sharded(model, wav: Tensor, lengths: Tensor) -> List[Tuple[str, List[int], List[int]]]:
    """Every rank calls this with the SAME global batch; each encodes + decodes its shard on its own GPU and all
    ranks return the full list of hypotheses (detokenised on every rank from the gathered ids)."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    w, l = shard_batch(wav, lengths, rank, world)
    dev = model._device
    enc, enc_len = model(w.to(dev), l.to(dev))
    ids, frames, counts = model.decoding.decode_device(model.head, enc, enc_len)
    ids, frames, counts = gather_hypotheses(ids, frames, counts, wav.shape[0])
    return model.decoding.to_hypotheses(ids.cpu(), frames.cpu(), counts.cpu())
