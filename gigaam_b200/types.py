"""Result types returned by the public API (same fields as the reference's gigaam/types.py:16-29)."""
from dataclasses import dataclass
from typing import List, Optional


@dataclass
class Word:
    text: str
    start: float
    end: float


@dataclass
class TranscriptionResult:
    text: str
    words: Optional[List[Word]] = None

    def __str__(self) -> str:
        return self.text
