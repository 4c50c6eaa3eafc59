"""Host-side batch pipeline around the reference surface (`model(wav, len)` + `model.decoding.decode`):
the H2D copy of batch i+1 (copy stream, pinned source) and the D2H copy + detokenisation of batch i-1 overlap the
device work of batch i.  Nothing here changes what is computed; it is the serving loop a caller such as the
reference's `transcribe_longform` / `train_utils/eval.py` would drive, written so the GPU never waits for PCIe."""
from __future__ import annotations

from typing import Iterable, Iterator, List, Tuple

import torch

Tensor = torch.Tensor


class BatchPipeline:
    def __init__(self, model):
        self.model = model
        self.dev = model._device
        self.copy_stream = torch.cuda.Stream(device=self.dev)

    def _upload(self, batch):
        wav, lengths = batch
        with torch.cuda.stream(self.copy_stream):
            wav_d = wav.to(self.dev, non_blocking=True)
            len_d = lengths.to(self.dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        return wav_d, len_d, ev

    @torch.inference_mode()
    def run(self, host_batches: Iterable[Tuple[Tensor, Tensor]]) -> Iterator[List[Tuple[str, List[int], List[int]]]]:
        """host_batches: iterable of (wav [B, N] float32 pinned host tensor, lengths [B] int64) -> hypotheses per batch."""
        model = self.model
        compute = torch.cuda.current_stream(self.dev)
        it = iter(host_batches)
        try:
            nxt = self._upload(next(it))
        except StopIteration:
            return
        prev = None
        while nxt is not None:
            wav_d, len_d, ev = nxt
            try:
                nxt = self._upload(next(it))          # next batch's PCIe transfer runs under this batch's kernels
            except StopIteration:
                nxt = None
            compute.wait_event(ev)
            wav_d.record_stream(compute)
            len_d.record_stream(compute)
            enc, enc_len = model(wav_d, len_d)
            ids, frames, counts = model.decoding.decode_device(model.head, enc, enc_len)
            host = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in (ids, frames, counts)]
            for h, t in zip(host, (ids, frames, counts)):
                h.copy_(t, non_blocking=True)
            done = torch.cuda.Event()
            done.record(compute)
            if prev is not None:
                prev[1].synchronize()
                yield model.decoding.to_hypotheses(*prev[0])
            prev = (host, done)
        if prev is not None:
            prev[1].synchronize()
            yield model.decoding.to_hypotheses(*prev[0])
