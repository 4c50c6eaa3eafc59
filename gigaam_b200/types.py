"""Result types returned by the public API (same fields and helpers as the reference's gigaam/types.py:16-67)."""
from dataclasses import dataclass
from typing import Iterator, List, Optional


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


@dataclass
class Segment:
    text: str
    start: float
    end: float
    words: Optional[List[Word]] = None


@dataclass
class LongformTranscriptionResult:
    """What `transcribe_longform` returns: the segments in recording order (gigaam/types.py:38-67)."""
    segments: List[Segment]

    @property
    def words(self) -> List[Word]:
        return [w for seg in self.segments if seg.words for w in seg.words]

    @property
    def has_word_timestamps(self) -> bool:
        return len(self.segments) > 0 and self.segments[0].words is not None

    @property
    def text(self) -> str:
        return " ".join(seg.text for seg in self.segments)

    def __str__(self) -> str:
        return self.text

    def __iter__(self) -> Iterator[Segment]:
        return iter(self.segments)

    def __len__(self) -> int:
        return len(self.segments)
