"""Generate tests/golden/*.npz from the REAL reference (run in the build container, where /root/reference
exists) and pin oracle/gigaam_oracle.py against it.

    python oracle/make_golden.py            # writes fixtures, asserts oracle == reference

The reference cannot be imported as a package offline (hydra / omegaconf / soundfile are absent,
gigaam/model.py:3-4, gigaam/utils.py:9), so those three are stubbed and its hot-path classes are built
directly from kwargs: gigaam.preprocess.FeatureExtractor, gigaam.encoder.ConformerEncoder,
gigaam.decoder.CTCHead / RNNTHead, gigaam.decoding.CTCGreedyDecoding / RNNTGreedyDecoding.
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


from oracle.ref_loader import build_reference  # noqa: E402


def run_case(model_name: str, batch: int, seconds: float, ragged: bool, out_name: str, seed: int = 0):
    from gigaam_b200 import synthetic
    from oracle import gigaam_oracle as orc

    ck = synthetic.synthetic_checkpoint(model_name, seed=seed)
    cfg, sd = ck["cfg"], ck["state_dict"]
    wav, wav_len = synthetic.synthetic_audio(batch, seconds, seed=1234, ragged=ragged)
    root, decoding = build_reference(cfg, sd)
    with torch.inference_mode():
        mel_ref, mel_len_ref = root.preprocessor(wav, wav_len)
        enc_ref, enc_len_ref = root.encoder(mel_ref, mel_len_ref)        # GigaAM.forward on CPU (model.py:33-35)
        pre_ref, _ = root.encoder.pre_encode(x=mel_ref.transpose(1, 2), lengths=mel_len_ref)
        dec_ref = decoding.decode(root.head, enc_ref, enc_len_ref) if decoding is not None else None
        # ---- oracle restatement on the same inputs
        mel_o = orc.log_mel(wav, sd, cfg["preprocessor"])
        pre = cfg["preprocessor"]
        mel_len_o = orc.logmel_out_len(wav_len, pre.get("hop_length", 160), pre.get("win_length", 400), pre.get("center", True))
        enc_o, enc_len_o, stages = orc.encoder_forward(mel_o, mel_len_o, sd, cfg["encoder"], return_all=True)
    valid = (torch.arange(enc_ref.shape[2])[None, :] < enc_len_ref[:, None])

    def rel(a, b):
        return float((a - b).norm() / b.norm())

    report = {
        "mel_max_abs": float((mel_o - mel_ref).abs().max()),
        "mel_len_equal": bool(torch.equal(mel_len_o, mel_len_ref)),
        "pre_encode_rel": rel(stages[0][valid], pre_ref[valid]),
        "enc_rel_valid": rel(enc_o.transpose(1, 2)[valid], enc_ref.transpose(1, 2)[valid]),
        "enc_len_equal": bool(torch.equal(enc_len_o, enc_len_ref)),
    }
    assert report["mel_len_equal"] and report["enc_len_equal"], report
    assert report["mel_max_abs"] < 2e-3, report
    assert report["enc_rel_valid"] < 1e-4, report
    arrays = dict(
        wav_seed=np.int64(1234), batch=np.int64(batch), seconds=np.float64(seconds), ragged=np.bool_(ragged),
        weight_seed=np.int64(seed),
        wav_len=wav_len.numpy(), mel=mel_ref.numpy().astype(np.float32), mel_len=mel_len_ref.numpy(),
        pre_encode=pre_ref.numpy().astype(np.float32),
        enc=enc_ref.numpy().astype(np.float32), enc_len=enc_len_ref.numpy(),
    )
    if dec_ref is not None:
        head = cfg["head"]["type"]
        dec_o = orc.ctc_greedy(enc_ref, enc_len_ref, sd) if head == "ctc" else orc.rnnt_greedy(
            enc_ref, enc_len_ref, sd, cfg["decoding"]["max_symbols_per_step"])
        for b, (text, ids, frames) in enumerate(dec_ref):
            assert ids == dec_o[b][0] and frames == dec_o[b][1], (model_name, b, "oracle decode != reference decode")
            arrays[f"ids_{b}"] = np.asarray(ids, dtype=np.int64)
            arrays[f"frames_{b}"] = np.asarray(frames, dtype=np.int64)
        report["tokens"] = [len(d[1]) for d in dec_ref]
        report["tokens_per_frame"] = float(sum(report["tokens"]) / max(int(enc_len_ref.sum()), 1))
        if head == "ctc":
            lg = orc.ctc_logits(enc_ref, sd)
            top2 = lg.topk(2, dim=-1).values
            arrays["ctc_margin"] = (top2[..., 0] - top2[..., 1]).numpy().astype(np.float32)
    out = ROOT / "tests" / "golden" / out_name
    np.savez_compressed(out, **arrays)
    print(out_name, report, f"{out.stat().st_size / 1024:.0f} KiB")
    return report


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 8)
    cases = {
        "v2_ctc": dict(batch=2, seconds=2.0, ragged=True, out_name="v2_ctc_b2_2s.npz"),
        "v2_rnnt": dict(batch=2, seconds=2.0, ragged=True, out_name="v2_rnnt_b2_2s.npz"),
        # v3 shape (RECALLED): conv1d k5 subsampling, LayerNorm conv-norm, depthwise k5, n_fft 320, center=False
        "v3_e2e_rnnt": dict(batch=2, seconds=2.0, ragged=True, out_name="v3_e2e_rnnt_b2_2s.npz"),
        # v1 shape: the rel_pos attention branch (encoder.py:191-228, 307-334); 6 s so that T' = 151 spans two key blocks
        "v1_ctc": dict(batch=2, seconds=6.0, ragged=True, out_name="v1_ctc_b2_6s.npz"),
    }
    for name in (sys.argv[1:] or list(cases)):
        run_case(name, **cases[name])
