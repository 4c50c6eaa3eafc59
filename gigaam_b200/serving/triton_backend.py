"""Triton python-backend model that puts the B200 engine behind the reference's Triton I/O contract.

The reference serves ASR as an ENSEMBLE of three models (triton_scripts/repos/gigaam_ctc_onnx/config.pbtxt:1-79,
gigaam_rnnt_onnx likewise): `preprocessing` (python, CPU) -> `ctc_encoder_onnx` / rnnt encoder (ONNX or TensorRT) ->
`ctc_postprocessing` (python, CPU), with the client-facing signature

    audio_batch    FP32  [-1]   all utterances of the request concatenated           (config.pbtxt:5-10)
    audio_lengths  INT64 [-1]   samples per utterance                                 (config.pbtxt:11-15)
    texts          STRING [-1]  one transcript per utterance                          (config.pbtxt:18-24)

Here ONE model with that same signature replaces the ensemble: the waveforms never leave the GPU between log-mel,
encoder and greedy decode, and nothing round-trips through fp16 feature tensors between Triton models
(repos/preprocessing/config.pbtxt:18-22).  Drop `model.py` (this file, or a two-line shim importing TritonPythonModel
from it) into `<repo>/gigaam_b200_asr/1/` next to a `config.pbtxt` written by `config_pbtxt()`; the model name, device
and optional checkpoint directory come from `parameters`.

`triton_python_backend_utils` only exists inside a Triton server process, so it is imported inside `execute` exactly as
the reference's backends do (repos/preprocessing/1/model.py:46); everything else is plain functions the tests call.
"""
from __future__ import annotations

import json
from typing import Any, Dict, List, Sequence

import numpy as np
import torch


def config_pbtxt(name: str = "gigaam_b200_asr", model_name: str = "v3_ctc", max_utterances: int = 64, gpu: int = 0) -> str:
    """config.pbtxt with the ensemble's client-facing I/O (triton_scripts/repos/gigaam_ctc_onnx/config.pbtxt:5-24)."""
    return f'''name: "{name}"
backend: "python"
max_batch_size: 0

input [
  {{ name: "audio_batch"   data_type: TYPE_FP32  dims: [-1] }},
  {{ name: "audio_lengths" data_type: TYPE_INT64 dims: [-1] }}
]
output [
  {{ name: "texts" data_type: TYPE_STRING dims: [-1] }}
]
parameters {{ key: "model_name" value {{ string_value: "{model_name}" }} }}
parameters {{ key: "max_utterances_per_step" value {{ string_value: "{max_utterances}" }} }}
instance_group [{{ kind: KIND_GPU gpus: [{gpu}] count: 1 }}]
'''


def split_concatenated(audio_batch: np.ndarray, audio_lengths: Sequence[int]) -> List[np.ndarray]:
    """The request layout of the reference's preprocessing backend (repos/preprocessing/1/model.py:55-66)."""
    audio_batch = np.asarray(audio_batch).reshape(-1)
    out, start = [], 0
    for n in audio_lengths:
        n = int(n)
        if n < 0 or start + n > audio_batch.size:
            raise ValueError(f"audio_lengths sum past audio_batch ({start + n} > {audio_batch.size})")
        out.append(audio_batch[start:start + n])
        start += n
    return out


def transcribe_concatenated(model, audio_batch: np.ndarray, audio_lengths: Sequence[int], max_utterances: int = 64) -> List[str]:
    """`audio_batch` / `audio_lengths` of one request -> transcripts in request order.  Utterances are length-bucketed
    into device batches of at most `max_utterances` (longform.plan_batches) and padded per batch; every batch is one
    `model(wav, lengths)` + `model.decoding.decode(...)` on the GPU."""
    from ..longform import plan_batches
    utts = split_concatenated(audio_batch, audio_lengths)
    texts: List[str] = [""] * len(utts)
    dev = model._device
    lengths = [u.size for u in utts]
    with torch.inference_mode():
        for batch in plan_batches(lengths, max_utterances):
            longest = max(lengths[i] for i in batch)
            wav = torch.zeros((len(batch), longest), dtype=torch.float32).pin_memory()
            for row, i in enumerate(batch):
                wav[row, : lengths[i]] = torch.from_numpy(np.ascontiguousarray(utts[i], dtype=np.float32))
            lens = torch.tensor([lengths[i] for i in batch], dtype=torch.int64)
            enc, enc_len = model(wav.to(dev, non_blocking=True), lens.to(dev))
            for row, (text, _, _) in enumerate(model.decoding.decode(model.head, enc, enc_len)):
                texts[batch[row]] = text
    return texts


class TritonPythonModel:
    """python-backend entry points (initialize / execute / finalize) -- same shape as the reference's backends."""

    def initialize(self, args: Dict[str, Any]) -> None:
        import gigaam_b200 as gigaam
        cfg = json.loads(args["model_config"])
        params = {k: v.get("string_value", "") for k, v in cfg.get("parameters", {}).items()}
        device = f"cuda:{args.get('model_instance_device_id', '0')}"
        self.max_utterances = int(params.get("max_utterances_per_step", "64") or 64)
        self.model = gigaam.load_model(params.get("model_name", "v3_ctc"), device=device,
                                       download_root=params.get("download_root") or None,
                                       synthetic=params.get("synthetic", "") == "1" or None)

    def execute(self, requests: Any) -> List[Any]:
        import triton_python_backend_utils as pb_utils  # type: ignore

        responses = []
        for request in requests:
            audio = pb_utils.get_input_tensor_by_name(request, "audio_batch").as_numpy()
            lengths = pb_utils.get_input_tensor_by_name(request, "audio_lengths").as_numpy()
            try:
                texts = transcribe_concatenated(self.model, audio, lengths, self.max_utterances)
                arr = np.array([t.encode("utf-8") for t in texts], dtype=object)
                responses.append(pb_utils.InferenceResponse(output_tensors=[pb_utils.Tensor("texts", arr)]))
            except Exception as exc:  # one bad request must not take the instance down
                responses.append(pb_utils.InferenceResponse(output_tensors=[], error=pb_utils.TritonError(str(exc))))
        return responses

    def finalize(self) -> None:
        self.model = None
