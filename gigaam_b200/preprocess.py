"""Log-mel front end with the reference's interface (gigaam/preprocess.py:12-98); the arithmetic is the fused
frame -> window -> DFT -> |.|^2 -> mel -> log CUDA kernel behind `gam_logmel`."""
from __future__ import annotations

import warnings
import wave
from subprocess import CalledProcessError, run
from typing import Tuple

import numpy as np
import torch
from torch import Tensor

from . import synthetic
from ._params import Bound, attach

SAMPLE_RATE = 16000


def load_audio(audio_path: str, sample_rate: int = SAMPLE_RATE) -> Tensor:
    """Same contract as the reference (gigaam/preprocess.py:12-40): mono float32 in [-1, 1] at `sample_rate`,
    decoded by ffmpeg.  When ffmpeg is not installed, 16-bit PCM WAV files already at `sample_rate` are read
    with the standard library instead (host I/O, not part of the accelerated path)."""
    cmd = ["ffmpeg", "-nostdin", "-threads", "0", "-i", audio_path, "-f", "s16le", "-ac", "1", "-acodec", "pcm_s16le",
           "-ar", str(sample_rate), "-"]
    try:
        audio = run(cmd, capture_output=True, check=True).stdout
    except CalledProcessError as exc:
        raise RuntimeError("Failed to load audio") from exc
    except FileNotFoundError:
        try:
            with wave.open(audio_path, "rb") as wf:
                if wf.getsampwidth() != 2 or wf.getframerate() != sample_rate:
                    raise RuntimeError("Failed to load audio: ffmpeg is missing and the file is not 16-bit PCM at "
                                       f"{sample_rate} Hz")
                pcm = np.frombuffer(wf.readframes(wf.getnframes()), dtype=np.int16)
                if wf.getnchannels() > 1:
                    pcm = pcm.reshape(-1, wf.getnchannels()).astype(np.float32).mean(axis=1).astype(np.int16)
                audio = pcm.tobytes()
        except (wave.Error, OSError) as exc:
            raise RuntimeError("Failed to load audio") from exc
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=UserWarning)
        return torch.frombuffer(bytearray(audio), dtype=torch.int16).float() / 32768.0


class FeatureExtractor(Bound):
    """Drop-in for gigaam.preprocess.FeatureExtractor (same ctor kwargs, buffers and `out_len`)."""

    def __init__(self, sample_rate: int, features: int, **kwargs):
        super().__init__()
        self.hop_length = kwargs.get("hop_length", sample_rate // 100)
        self.win_length = kwargs.get("win_length", sample_rate // 40)
        self.n_fft = kwargs.get("n_fft", sample_rate // 40)
        self.center = kwargs.get("center", True)
        attach(self, "featurizer.0.spectrogram.window", synthetic.hann_window(self.win_length))
        attach(self, "featurizer.0.mel_scale.fb", synthetic.mel_filterbank(self.n_fft // 2 + 1, features, sample_rate))

    def out_len(self, input_lengths: Tensor) -> Tensor:
        """gigaam/preprocess.py:78-92"""
        if self.center:
            return input_lengths.div(self.hop_length, rounding_mode="floor").add(1).long()
        return (input_lengths - self.win_length).div(self.hop_length, rounding_mode="floor").add(1).long()

    def forward(self, input_signal: Tensor, length: Tensor) -> Tuple[Tensor, Tensor]:
        eng = self._engine()
        wav = input_signal.to(device=eng.device, dtype=torch.float32)
        return eng.logmel(wav), self.out_len(length)
