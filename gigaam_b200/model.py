"""Model wrappers with the reference's surface (gigaam/model.py:16-140): GigaAM.forward / embed_audio /
prepare_wav, GigaAMASR.transcribe / _decode, `_device`, `_dtype`, `cfg`, `preprocessor`, `encoder`, `head`,
`decoding`.  Hydra `_target_` instantiation is replaced by a small registry keyed on the same class names."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple, Union

import numpy as np
import torch
from torch import Tensor, nn

from .decoder import CTCHead, RNNTHead
from .decoding import CTCGreedyDecoding, RNNTGreedyDecoding
from .encoder import ConformerEncoder
from .engine import Engine
from .preprocess import SAMPLE_RATE, FeatureExtractor, load_audio
from .types import TranscriptionResult, Word

LONGFORM_THRESHOLD = 25 * SAMPLE_RATE

_REGISTRY = {
    "FeatureExtractor": FeatureExtractor, "ConformerEncoder": ConformerEncoder, "CTCHead": CTCHead,
    "RNNTHead": RNNTHead, "CTCGreedyDecoding": CTCGreedyDecoding, "RNNTGreedyDecoding": RNNTGreedyDecoding,
}


def _plain(obj):
    """OmegaConf-like containers -> plain dict / list."""
    if hasattr(obj, "items"):
        return {k: _plain(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)) or type(obj).__name__ == "ListConfig":
        return [_plain(v) for v in obj]
    return obj


def instantiate(section: Dict, default_cls: str):
    """Stand-in for hydra.utils.instantiate (gigaam/model.py:24-25,93-94): `_target_: gigaam.<mod>.<Class>`."""
    kw = dict(section)
    target = kw.pop("_target_", None)
    kw.pop("type", None)
    name = target.rsplit(".", 1)[-1] if target else default_cls
    if name not in _REGISTRY:
        raise ValueError(f"unknown component {target!r}")
    return _REGISTRY[name](**kw)


def normalize_cfg(cfg) -> Dict:
    """Bring a checkpoint cfg (Hydra-style, with `_target_`s) or a synthetic cfg to one plain-dict shape:
    sections `preprocessor`, `encoder`, optional `head` (with `type`) and `decoding`."""
    c = _plain(cfg)
    out = dict(c)
    head = c.get("head")
    if head is not None and "type" not in head:
        tgt = str(head.get("_target_", ""))
        head = dict(head)
        head["type"] = "rnnt" if "RNNT" in tgt or "decoder" in head else "ctc"
        out["head"] = head
    return out


class GigaAM(nn.Module):
    """Giga Acoustic Model (self-supervised encoder) -- drop-in for gigaam.model.GigaAM."""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self._ncfg = normalize_cfg(cfg)
        self.preprocessor = instantiate(self._ncfg["preprocessor"], "FeatureExtractor")
        self.encoder = instantiate(self._ncfg["encoder"], "ConformerEncoder")
        self.preprocessor._bind(self)
        self.encoder._bind(self)
        self.__dict__["_engine_obj"] = None

    # ---- engine lifecycle: rebuilt lazily whenever parameters move / change dtype / get reloaded
    def _invalidate_engine(self) -> None:
        self.__dict__["_engine_obj"] = None

    def _get_engine(self) -> Engine:
        eng = self.__dict__.get("_engine_obj")
        if eng is None:
            dev = self._device
            if dev.type != "cuda":
                raise RuntimeError("gigaam_b200 has no CPU path: move the model to a CUDA (sm_100a) device first")
            sd = {k: v for k, v in self.state_dict().items()}
            eng = Engine(self._engine_cfg(), sd, dev, pack_cache=self._pack_cache_path())
            self.__dict__["_engine_obj"] = eng
        return eng

    def _pack_cache_path(self) -> Optional[str]:
        """On-disk cache of the packed weights, set by load_model for checkpoints read from a file: keyed by the
        checkpoint's md5 and the parameter dtype the engine was built from (fp16_encoder rounds before packing)."""
        base = self.__dict__.get("_pack_cache_base")
        return None if base is None else f"{base}.{str(self._dtype).split('.')[-1]}.b200pack"

    def _engine_cfg(self) -> Dict:
        c = self._ncfg
        pre = {k: v for k, v in c["preprocessor"].items() if k != "_target_"}
        enc = dict(self.encoder.cfg)
        out = dict(model_name=c.get("model_name", "custom"), preprocessor=pre, encoder=enc)
        if c.get("head") is not None:
            out["head"] = {k: v for k, v in c["head"].items() if k != "_target_"}
            out["decoding"] = {k: v for k, v in c.get("decoding", {}).items() if k != "_target_"}
        return out

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        res = super().load_state_dict(state_dict, strict=strict, assign=assign)
        self._invalidate_engine()
        return res

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._invalidate_engine()
        return out

    # ---- reference surface
    def forward(self, features: Tensor, feature_lengths: Tensor) -> Tuple[Tensor, Tensor]:
        """wav [B, N], lengths [B] -> (encoded [B, d_model, T'], encoded_len [B] int32)  (gigaam/model.py:27-37)"""
        features, feature_lengths = self.preprocessor(features, feature_lengths)
        return self.encoder(features, feature_lengths)

    @property
    def _device(self) -> torch.device:
        return next(self.parameters()).device

    @property
    def _dtype(self) -> torch.dtype:
        return next(self.parameters()).dtype

    def prepare_wav(self, wav_file: Union[str, Tensor, np.ndarray]) -> Tuple[Tensor, Tensor]:
        """gigaam/model.py:47-55; additionally accepts an in-memory mono waveform (tensor / ndarray)."""
        if isinstance(wav_file, str):
            wav = load_audio(wav_file)
        else:
            wav = torch.as_tensor(wav_file, dtype=torch.float32).reshape(-1)
        wav = wav.to(self._device).to(self._dtype).unsqueeze(0)
        length = torch.full([1], wav.shape[-1], device=self._device)
        return wav, length

    def embed_audio(self, wav_file) -> Tuple[Tensor, Tensor]:
        """gigaam/model.py:57-63"""
        wav, length = self.prepare_wav(wav_file)
        return self.forward(wav, length)


class GigaAMASR(GigaAM):
    """Giga Acoustic Model for Speech Recognition -- drop-in for gigaam.model.GigaAMASR."""

    def __init__(self, cfg):
        super().__init__(cfg)
        head_cfg = self._ncfg["head"]
        dec_cfg = {k: v for k, v in self._ncfg["decoding"].items() if k != "type"}
        self.head = instantiate(head_cfg, "RNNTHead" if head_cfg.get("type") == "rnnt" else "CTCHead")
        self.head._bind(self)
        self.decoding = instantiate(dec_cfg, "RNNTGreedyDecoding" if head_cfg.get("type") == "rnnt" else "CTCGreedyDecoding")

    def _decode(self, encoded: Tensor, encoded_len: Tensor, wav_lens: Tensor, word_timestamps: bool = False
                ) -> List[Tuple[str, Optional[List[Word]]]]:
        """gigaam/model.py:96-124"""
        if not word_timestamps:
            return [(t, None) for t, _, _ in self.decoding.decode(self.head, encoded, encoded_len)]
        # tokens are grouped into words on the device (csrc/words.cu); one D2H copy brings ids and word records back
        ids, frames, counts = self.decoding.decode_device(self.head, encoded, encoded_len)
        rec = self._get_engine().group_words(ids, frames, counts, self._word_flags())
        return self._words_from_records(ids.cpu(), counts.cpu(), encoded_len.cpu(), wav_lens.cpu(), [t.cpu() for t in rec])

    def _word_flags(self) -> Tensor:
        """Per-token flag table of the device word grouping (timestamps_utils.token_flag_table), built once."""
        flags = self.__dict__.get("_token_flags")
        if flags is None or flags.device != self._device:        # device-resident: the grouping runs inside CUDA graphs
            from .timestamps_utils import token_flag_table
            flags = token_flag_table(self.decoding.tokenizer).to(self._device)
            self.__dict__["_token_flags"] = flags
        return flags

    def _words_from_records(self, ids: Tensor, counts: Tensor, encoded_len: Tensor, wav_lens: Tensor, rec: List[Tensor]
                            ) -> List[Tuple[str, Optional[List[Word]]]]:
        """Host copies of (ids, counts, encoded_len, wav_lens, gam_group_words records) -> [(text, words)] per utterance."""
        from .timestamps_utils import compute_frame_shift, words_from_device
        tok = self.decoding.tokenizer
        ws, we, wf, wn, nw = rec
        out: List[Tuple[str, Optional[List[Word]]]] = []
        for i, n in enumerate(counts.tolist()):
            row = ids[i, :n].tolist()
            k = int(nw[i])
            shift = compute_frame_shift(int(wav_lens[i]), int(encoded_len[i]))
            words = words_from_device(tok, row, ws[i, :k].tolist(), we[i, :k].tolist(), wf[i, :k].tolist(), wn[i, :k].tolist(), shift)
            out.append((tok.decode(row), words))
        return out

    @torch.inference_mode()
    def transcribe(self, wav_file, word_timestamps: bool = False) -> TranscriptionResult:
        """gigaam/model.py:126-140"""
        wav, length = self.prepare_wav(wav_file)
        if length.item() > LONGFORM_THRESHOLD:
            raise ValueError("Too long wav file, use 'transcribe_longform' method.")
        encoded, encoded_len = self.forward(wav, length)
        text, words = self._decode(encoded, encoded_len, length, word_timestamps)[0]
        return TranscriptionResult(text=text, words=words)

    @torch.inference_mode()
    def transcribe_longform(self, wav_file, word_timestamps: bool = False, fr_batch_size: int = 16, fr_num_workers: int = 0,
                            segments: Optional[List[Tensor]] = None, boundaries: Optional[List[Tuple[float, float]]] = None,
                            **kwargs):
        """gigaam/model.py:195-259.  Segmentation is pluggable: pass `segments` / `boundaries` from any VAD (the
        reference's pyannote pipeline, gigaam/vad_utils.py, is third party and not vendored); without them the
        recording is cut at low-energy points (`longform.split_on_energy`, kwargs forwarded).  Segments are
        length-bucketed into batches of `fr_batch_size`; `fr_num_workers` is accepted for signature compatibility."""
        from .longform import split_on_energy, transcribe_segments
        if segments is None:
            wav = load_audio(wav_file) if isinstance(wav_file, str) else torch.as_tensor(wav_file, dtype=torch.float32).reshape(-1)
            segments, boundaries = split_on_energy(wav, SAMPLE_RATE, **kwargs)
        elif boundaries is None:
            raise ValueError("boundaries are required when segments are given")
        return transcribe_segments(self, segments, boundaries, word_timestamps, fr_batch_size)

    @torch.inference_mode()
    def transcribe_batch(self, wav: Tensor, lengths: Tensor) -> List[str]:
        """Batched entry (the path eval.py / transcribe_longform drive: model(wav, len) -> decoding.decode)."""
        encoded, encoded_len = self.forward(wav, lengths)
        return [t for t, _, _ in self.decoding.decode(self.head, encoded, encoded_len)]
