"""Host-side batch pipeline around the reference surface (`model(wav, len)` + `model.decoding.decode`):
the H2D copy of batch i+1 (copy stream, pinned source) and the D2H copy + detokenisation of batch i-1 overlap the
device work of batch i, and the ~270 kernel launches of a batch are replayed as ONE CUDA graph per input shape
(captured from the very same `model(...)` / `decode_device(...)` calls), so the GPU waits neither for PCIe nor for
the host's launch loop.  Nothing here changes what is computed; it is the serving loop a caller such as the
reference's `transcribe_longform` / `train_utils/eval.py` would drive."""
from __future__ import annotations

from typing import Dict, Iterable, Iterator, List, Tuple

import torch

Tensor = torch.Tensor


class _ShapeGraph:
    """The whole device step for one (B, N) input shape, captured once: static inputs -> static outputs."""

    def __init__(self, model, B: int, N: int, dev: torch.device):
        self.wav = torch.zeros((B, N), dtype=torch.float32, device=dev)
        self.len = torch.full((B,), N, dtype=torch.int64, device=dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):                      # eager warm-up: plans, workspaces, tensor maps
                self._step(model)
        side.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=side):
            self.out = self._step(model)
        torch.cuda.current_stream(dev).wait_stream(side)
        # the graph baked the pointers of the engine's scratch buffers for this shape: own them, so the engine's
        # per-kind LRU may forget the shape without freeing memory this graph still writes
        self._held = model._get_engine().held_workspaces(B, N)

    def _step(self, model):
        enc, enc_len = model(self.wav, self.len)
        return model.decoding.decode_device(model.head, enc, enc_len)


class BatchPipeline:
    def __init__(self, model, use_graph: bool = True, max_graphs: int = 4):
        self.model = model
        self.dev = model._device
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.use_graph = use_graph
        self.max_graphs = max_graphs
        self._graphs: Dict[Tuple[int, int], _ShapeGraph] = {}

    def _upload(self, batch):
        wav, lengths = batch
        with torch.cuda.stream(self.copy_stream):
            wav_d = wav.to(self.dev, non_blocking=True)
            len_d = lengths.to(self.dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        return wav_d, len_d, ev

    def _graph_for(self, B: int, N: int) -> _ShapeGraph:
        g = self._graphs.get((B, N))
        if g is None:
            if len(self._graphs) >= self.max_graphs:
                self._graphs.pop(next(iter(self._graphs)))
            g = _ShapeGraph(self.model, B, N, self.dev)
            self._graphs[(B, N)] = g
        return g

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
            if self.use_graph:
                g = self._graph_for(wav_d.shape[0], wav_d.shape[1])
                g.wav.copy_(wav_d, non_blocking=True)   # device-to-device refill of the graph's static input
                g.len.copy_(len_d, non_blocking=True)
                g.graph.replay()
                ids, frames, counts = g.out
            else:
                enc, enc_len = model(wav_d, len_d)
                ids, frames, counts = model.decoding.decode_device(model.head, enc, enc_len)
            host = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in (ids, frames, counts)]
            for h, t in zip(host, (ids, frames, counts)):
                h.copy_(t, non_blocking=True)           # stream-ordered before the next replay overwrites the outputs
            done = torch.cuda.Event()
            done.record(compute)
            if prev is not None:
                prev[1].synchronize()
                yield model.decoding.to_hypotheses(*prev[0])
            prev = (host, done)
        if prev is not None:
            prev[1].synchronize()
            yield model.decoding.to_hypotheses(*prev[0])
