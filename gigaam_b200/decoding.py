"""Greedy decoders with the reference's interface (gigaam/decoding.py).  `decode(head, encoded, lengths)`
returns the same `List[(text, token_ids, token_frames)]`; head projection, argmax, CTC collapse and the whole
RNN-T prediction/joint loop run on the device, then one D2H copy brings ids / frames / counts back for the
host-side detokenisation (which the reference also does on the host, decoding.py:93-96,207)."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from torch import Tensor


class Tokenizer:
    """gigaam/decoding.py:10-44 -- charwise vocabulary or a SentencePiece model."""

    def __init__(self, vocab: List[str], model_path: Optional[str] = None):
        self.charwise = model_path is None
        if self.charwise:
            self.vocab = vocab
        else:
            from sentencepiece import SentencePieceProcessor
            self.model = SentencePieceProcessor()
            self.model.load(model_path)

    def decode(self, tokens: List[int]) -> str:
        if self.charwise:
            return "".join(self.vocab[tok] for tok in tokens)
        return self.model.decode(tokens)

    def __len__(self):
        return len(self.vocab) if self.charwise else len(self.model)

    def id_to_str(self, token_id: int) -> str:
        if self.charwise:
            return self.vocab[token_id]
        return self.model.IdToPiece(token_id)


def _as_btd(encoded: Tensor) -> Tensor:
    """[B, d, T] (the encoder's transposed view) -> contiguous [B, T, d] without a copy when possible."""
    x = encoded.transpose(1, 2)
    return x if x.is_contiguous() else x.contiguous()


class _GreedyBase:
    def __init__(self, vocabulary: List[str], model_path: Optional[str] = None):
        self.tokenizer = Tokenizer(vocabulary, model_path)
        self.blank_id = len(self.tokenizer)

    def decode_device(self, head, encoded: Tensor, lengths: Tensor, packed: Optional[Tensor] = None
                      ) -> Tuple[Tensor, Tensor, Tensor]:
        """Device-resident result: ids [B, max_out] i32, frames [B, max_out] i32, counts [B] i32 (views of `packed`, an
        `Engine.packed_hypotheses` buffer, when one is given: the layout the multi-GPU gather ships)."""
        eng = head._engine()
        assert eng.num_classes == len(self.tokenizer) + 1, \
            f"Num classes {eng.num_classes} != len(vocab)+1 {len(self.tokenizer) + 1}"
        enc = _as_btd(encoded.to(device=eng.device, dtype=torch.float32))
        return eng.greedy(enc, lengths, packed)

    @torch.inference_mode()
    def decode(self, head, encoded: Tensor, lengths: Tensor) -> List[Tuple[str, List[int], List[int]]]:
        ids, frames, counts = self.decode_device(head, encoded, lengths)
        return self.to_hypotheses(ids.cpu(), frames.cpu(), counts.cpu())

    def to_hypotheses(self, ids: Tensor, frames: Tensor, counts: Tensor) -> List[Tuple[str, List[int], List[int]]]:
        out = []
        for b, n in enumerate(counts.tolist()):
            tok = ids[b, :n].tolist()
            out.append((self.tokenizer.decode(tok), tok, frames[b, :n].tolist()))
        return out


class CTCGreedyDecoding(_GreedyBase):
    """gigaam/decoding.py:47-96"""


class RNNTGreedyDecoding(_GreedyBase):
    """gigaam/decoding.py:99-207"""

    def __init__(self, vocabulary: List[str], model_path: Optional[str] = None, max_symbols_per_step: int = 10):
        super().__init__(vocabulary, model_path)
        self.max_symbols = max_symbols_per_step
