"""Word timestamps from the (token id, frame) pairs the greedy kernels emit -- host post-processing with the
reference's semantics (gigaam/timestamps_utils.py:8-53): a word starts at its first token's frame and ends one
frame after its last token; words split on a space token or a SentencePiece piece starting with U+2581."""
from __future__ import annotations

from typing import List

from .preprocess import SAMPLE_RATE
from .types import Word

_SP_SPACE = "▁"


def compute_frame_shift(audio_length_samples: int, seq_len: int) -> float:
    return audio_length_samples / SAMPLE_RATE / seq_len


def frames_to_words(tokenizer, token_ids: List[int], token_frames: List[int], frame_shift: float) -> List[Word]:
    words: List[Word] = []
    pieces: List[str] = []
    frames: List[int] = []

    def flush() -> None:
        text = "".join(pieces).strip()
        if text:
            words.append(Word(text=text, start=frames[0] * frame_shift, end=(frames[-1] + 1) * frame_shift))
        pieces.clear()
        frames.clear()

    for tok, fr in zip(token_ids, token_frames):
        piece = tokenizer.id_to_str(tok)
        if piece == " ":
            flush()
            continue
        if piece.startswith(_SP_SPACE):
            flush()
            piece = piece[1:]
        pieces.append(piece)
        frames.append(fr)
    flush()
    return words
