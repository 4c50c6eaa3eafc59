"""Calibrate the synthetic RNN-T blank bias (gigaam_b200/synthetic.py:RNNT_BLANK_BIAS) by bisection so the
seeded random head emits a realistic token rate (~0.5 tok/frame at V+1=34, ~0.2 at V+1=1025).
Test infrastructure: runs the oracle on CPU.   python oracle/calibrate_rnnt.py v2_rnnt 0.5
"""
from __future__ import annotations

import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main(model_name: str, target: float):
    from gigaam_b200 import synthetic
    from oracle import gigaam_oracle as orc

    synthetic.RNNT_BLANK_BIAS[model_name] = 0.0
    ck = synthetic.synthetic_checkpoint(model_name, seed=0)
    cfg, sd = ck["cfg"], ck["state_dict"]
    wav, wav_len = synthetic.synthetic_audio(4, 5.0, seed=1234, ragged=True)
    with torch.inference_mode():
        enc, enc_len = orc.model_forward(wav, wav_len, sd, cfg)
    base = sd["head.joint.joint_net.1.bias"][-1].item()

    def rate(bias: float) -> float:
        sd["head.joint.joint_net.1.bias"][-1] = base + bias
        with torch.inference_mode():
            dec = orc.rnnt_greedy(enc, enc_len, sd, cfg["decoding"]["max_symbols_per_step"])
        return sum(len(d[0]) for d in dec) / float(enc_len.sum())

    lo, hi = 0.0, 40.0
    for _ in range(14):
        mid = 0.5 * (lo + hi)
        r = rate(mid)
        print(f"bias {mid:.4f} -> {r:.3f} tok/frame")
        if r > target:
            lo = mid
        else:
            hi = mid
    print("calibrated bias ~", round(0.5 * (lo + hi), 3))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "v2_rnnt", float(sys.argv[2]) if len(sys.argv) > 2 else 0.5)
