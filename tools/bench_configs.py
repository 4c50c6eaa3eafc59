"""Device-timed throughput of every BASELINE.json config on one GPU (the contract bench is bench.py = configs[1]).

    python tools/bench_configs.py [c1 c2 c3 c4 c5]
"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

import torch  # noqa: E402

import gigaam_b200 as gigaam  # noqa: E402
from bench import ClockSampler  # noqa: E402  (nvidia-smi clocks / throttle reasons during the timed region)

CONFIGS = {
    "c1": ("v2_ctc", 1, 5.0), "c2": ("v2_ctc", 64, 10.0), "c3": ("v2_rnnt", 32, 15.0),
    "c4": ("v3_e2e_rnnt", 32, 10.0),   # per-GPU share of 256 x 10 s over 8 GPUs
    "c5": ("v2_ssl", 128, 25.0),
    # not BASELINE configs: the rel_pos (v1) model on the c2 / c5 shapes
    "v1c2": ("v1_ctc", 64, 10.0), "v1c5": ("v1_ssl", 128, 25.0),
    # not a BASELINE config either: the c2 batch with ragged lengths (uniform 1 .. 10 s inside the 10 s buffer) -- what varlen
    # execution buys: only the frames that exist run through the encoder
    "c2r": ("v2_ctc", 64, 10.0),
}


def run(name):
    model_name, B, sec = CONFIGS[name]
    dev = torch.device("cuda", 0)
    model = gigaam.load_model(model_name, device=dev, synthetic=True)
    eng = model._get_engine()
    wav, wav_len = gigaam.synthetic_audio(B, sec, seed=1234)
    if name.endswith("r"):
        g = torch.Generator().manual_seed(7)
        wav_len = (torch.rand(B, generator=g) * 0.9 + 0.1).mul(sec * 16000).long()
        wav_len[0] = int(sec * 16000)
        wav = wav * (torch.arange(wav.shape[1])[None, :] < wav_len[:, None])
    audio_fraction = float(wav_len.sum()) / float(B * wav.shape[1])
    wav, wav_len = wav.to(dev), wav_len.to(dev)
    has_head = hasattr(model, "head")

    def step():
        enc, enc_len = model(wav, wav_len)
        if has_head:
            return model.decoding.decode_device(model.head, enc, enc_len)
        return enc, enc_len

    import time
    for _ in range(3):
        out = step()
    torch.cuda.synchronize()
    n = max(5, int(400.0 / max(1.0, B * sec / 50.0)))          # ~0.4 s of device time: enough nvidia-smi samples
    sampler = ClockSampler(0)
    sampler.start()
    time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    clocks = sampler.stop(t0, t1)
    ms = e0.elapsed_time(e1) / n
    eng.profile_begin()
    step()
    prof = eng.profile_end()
    res = {"config": name, "model": model_name, "batch": B, "seconds": sec, "ms_per_batch": round(ms, 3),
           "utt_per_s": round(B / ms * 1e3, 1), "rtfx": round(B * sec * audio_fraction / ms * 1e3), "audio_fraction": round(audio_fraction, 3),
           "steps": n, "clocks": clocks,
           "classes_ms": {k: round(v[0], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}}
    if has_head:
        counts = out[2]
        res["tokens_per_frame"] = round(float(counts.sum()) / float(B * eng.encoded_frames(eng.logmel_frames(wav.shape[1]))), 3)
    print(json.dumps(res), flush=True)
    del model
    torch.cuda.empty_cache()


if __name__ == "__main__":
    for c in (sys.argv[1:] or list(CONFIGS)):
        run(c)
