"""Result records of the public API.  Field names, defaults, `str()` and the long-form helpers follow the reference's
result classes (gigaam/types.py:16-67) because callers and tests touch them (`result.words is None`, `str(result) == text`,
`for segment in result`); the implementation is a small slotted record base instead of dataclasses -- a long-form
transcript holds one `Word` per spoken word, and slotted objects are a third of the size."""
from __future__ import annotations

from typing import Any, Dict, Iterator, List, Optional, Tuple


class _Record:
    """Positional / keyword construction over `_fields`, value equality and a readable repr."""
    __slots__ = ()
    _fields: Tuple[str, ...] = ()
    _defaults: Dict[str, Any] = {}

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        if len(args) > len(self._fields):
            raise TypeError(f"{type(self).__name__} takes at most {len(self._fields)} positional arguments")
        given = dict(zip(self._fields, args))
        for key, value in kwargs.items():
            if key not in self._fields:
                raise TypeError(f"{type(self).__name__} has no field {key!r}")
            if key in given:
                raise TypeError(f"{type(self).__name__} got {key!r} twice")
            given[key] = value
        for name in self._fields:
            if name in given:
                setattr(self, name, given[name])
            elif name in self._defaults:
                setattr(self, name, self._defaults[name])
            else:
                raise TypeError(f"{type(self).__name__} missing required field {name!r}")

    def __eq__(self, other: object) -> bool:
        return type(other) is type(self) and all(getattr(self, n) == getattr(other, n) for n in self._fields)

    def __repr__(self) -> str:
        return f"{type(self).__name__}({', '.join(f'{n}={getattr(self, n)!r}' for n in self._fields)})"


class Word(_Record):
    """One word with its start / end time in seconds."""
    __slots__ = _fields = ("text", "start", "end")
    text: str
    start: float
    end: float


class TranscriptionResult(_Record):
    """`transcribe()` result: `words` stays None unless word timestamps were requested."""
    __slots__ = _fields = ("text", "words")
    _defaults = {"words": None}
    text: str
    words: Optional[List[Word]]

    def __str__(self) -> str:
        return self.text


class Segment(_Record):
    """One speech segment of a long recording (times in seconds from the start of the recording)."""
    __slots__ = _fields = ("text", "start", "end", "words")
    _defaults = {"words": None}
    text: str
    start: float
    end: float
    words: Optional[List[Word]]


class LongformTranscriptionResult(_Record):
    """`transcribe_longform()` result: the segments in recording order."""
    __slots__ = _fields = ("segments",)
    segments: List[Segment]

    @property
    def text(self) -> str:
        return " ".join(seg.text for seg in self.segments)

    @property
    def words(self) -> List[Word]:
        return [word for seg in self.segments for word in (seg.words or [])]

    @property
    def has_word_timestamps(self) -> bool:
        return bool(self.segments) and self.segments[0].words is not None

    def __str__(self) -> str:
        return self.text

    def __len__(self) -> int:
        return len(self.segments)

    def __iter__(self) -> Iterator[Segment]:
        return iter(self.segments)
