"""Development harness run under gpurun: every step is a separate subprocess (a sticky CUDA error in one kernel
must not hide the state of the others) with its own timeout; logs land in gpurun_out/check_<step>.log.

    python tools/gpu_check.py            # all steps
    python tools/gpu_check.py gemm attn  # selected steps
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
OUT = ROOT / "gpurun_out"

STEPS = ["gemm", "attn", "logmel", "pre_encode", "layers", "ctc", "rnnt", "e2e", "timing"]


def _engine(model="v2_ctc", n_layers=None, seed=0):
    import torch
    from gigaam_b200 import synthetic
    from gigaam_b200.engine import Engine
    ck = synthetic.synthetic_checkpoint(model, seed=seed, n_layers=n_layers)
    eng = Engine(ck["cfg"], ck["state_dict"], torch.device("cuda", 0))
    return ck, eng


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def step_gemm():
    import ctypes as C
    import torch
    import torch.nn.functional as F
    ck, eng = _engine(n_layers=1)
    lib, h = eng.lib, eng.handle
    dev = torch.device("cuda", 0)
    g = torch.Generator(device="cpu").manual_seed(1)
    ok = True
    for (M, N, K) in [(128, 256, 64), (1000, 768, 768), (16064, 3072, 768), (777, 768, 3072), (300, 1536, 768)]:
        A = (torch.randn(M, K, generator=g) * 0.5).half().to(dev)
        W = (torch.randn(N, K, generator=g) / K ** 0.5).half().to(dev)
        bias = torch.randn(N, generator=g).to(dev)
        ref = A.float() @ W.float().t() + bias
        s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for kind, name in [(0, "bias_f16"), (1, "silu_f16"), (2, "glu_f16"), (3, "res_f32"), (4, "bias_f32")]:
            if kind == 2:
                out = torch.zeros(M, N // 2, dtype=torch.float16, device=dev)
                r4 = ref.view(M, N // 256, 2, 128)
                want = (r4[:, :, 0] * torch.sigmoid(r4[:, :, 1])).reshape(M, N // 2)
                ldo = N // 2
                res = None
            elif kind in (0, 1):
                out = torch.zeros(M, N, dtype=torch.float16, device=dev)
                want = ref if kind == 0 else F.silu(ref)
                ldo, res = N, None
            else:
                out = torch.zeros(M, N, dtype=torch.float32, device=dev)
                res = torch.randn(M, N, generator=g).to(dev) if kind == 3 else None
                want = res + 0.5 * ref if kind == 3 else ref
                ldo = N
            rc = lib.gam_test_gemm(h, kind, A.data_ptr(), W.data_ptr(), bias.data_ptr(), res.data_ptr() if res is not None else None,
                                   out.data_ptr(), M, N, K, ldo, 0.5, s)
            torch.cuda.synchronize()
            err = rel(out.float(), want)
            mx = float((out.float() - want).abs().max())
            good = rc == 0 and err < 2e-3
            ok &= good
            print(f"gemm M={M} N={N} K={K} {name}: rc={rc} rel={err:.3e} maxabs={mx:.3e} {'OK' if good else 'FAIL'}", flush=True)
    return ok


def step_attn():
    import ctypes as C
    import torch
    ck, eng = _engine(n_layers=1)
    lib, h = eng.lib, eng.handle
    dev = torch.device("cuda", 0)
    g = torch.Generator(device="cpu").manual_seed(2)
    ok = True
    for (B, T, lens) in [(1, 128, None), (2, 51, [51, 30]), (3, 251, [251, 200, 97]), (2, 376, [376, 129]), (1, 626, None)]:
        d, H, dk = 768, 16, 48
        qkv = (torch.randn(B * T, 3 * d, generator=g)).half().to(dev)
        out = torch.zeros(B * T, d, dtype=torch.float16, device=dev)
        klen = torch.tensor(lens, dtype=torch.int32, device=dev) if lens else None
        s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        rc = lib.gam_test_attention(h, qkv.data_ptr(), klen.data_ptr() if klen is not None else None, out.data_ptr(), B, T, s)
        torch.cuda.synchronize()
        x = qkv.float().view(B, T, 3, H, dk)
        q, k, v = x[:, :, 0].transpose(1, 2), x[:, :, 1].transpose(1, 2), x[:, :, 2].transpose(1, 2)
        sc = q @ k.transpose(-1, -2) / dk ** 0.5
        if lens:
            valid = torch.arange(T, device=dev)[None, :] < klen[:, None]
            sc = sc.masked_fill(~valid[:, None, None, :], float("-inf"))
        want = (torch.softmax(sc, -1) @ v).transpose(1, 2).reshape(B * T, d)
        err = rel(out.float(), want)
        good = rc == 0 and err < 3e-3
        ok &= good
        print(f"attn B={B} T={T} lens={lens}: rc={rc} rel={err:.3e} maxabs={float((out.float()-want).abs().max()):.3e} {'OK' if good else 'FAIL'}", flush=True)
    return ok


def _inputs(batch, seconds, ragged):
    from gigaam_b200 import synthetic
    return synthetic.synthetic_audio(batch, seconds, seed=1234, ragged=ragged)


def step_logmel():
    import torch
    from oracle import gigaam_oracle as orc
    ck, eng = _engine(n_layers=1)
    ok = True
    for (B, sec, ragged) in [(2, 2.0, True), (3, 10.0, False), (1, 0.5, False)]:
        wav, wav_len = _inputs(B, sec, ragged)
        want = orc.log_mel(wav, ck["state_dict"], ck["cfg"]["preprocessor"])
        got = eng.logmel(wav.cuda())
        torch.cuda.synchronize()
        d = (got.cpu() - want).abs()
        good = tuple(got.shape) == tuple(want.shape) and float(d.max()) < 5e-3
        ok &= good
        print(f"logmel B={B} sec={sec}: shape={tuple(got.shape)} maxabs={float(d.max()):.3e} mean={float(d.mean()):.3e} {'OK' if good else 'FAIL'}", flush=True)
    return ok


def _stage_compare(n_layers_list, model="v2_ctc", B=2, sec=2.0, ragged=True, full_layers=16):
    import torch
    from oracle import gigaam_oracle as orc
    ck, eng = _engine(model, n_layers=full_layers)
    cfg, sd = ck["cfg"], ck["state_dict"]
    wav, wav_len = _inputs(B, sec, ragged)
    pre = cfg["preprocessor"]
    with torch.inference_mode():
        mel = orc.log_mel(wav, sd, pre)
        mel_len = orc.logmel_out_len(wav_len, 160, 400, True)
        enc_o, len_o, stages = orc.encoder_forward(mel, mel_len, sd, cfg["encoder"], n_layers_run=max(n_layers_list), return_all=True)
    valid = torch.arange(stages[0].shape[1])[None, :] < len_o[:, None]
    ok = True
    for n in n_layers_list:
        enc, enc_len = eng.encode(mel.cuda(), mel_len.cuda(), n_layers_run=n)
        torch.cuda.synchronize()
        got = enc.cpu()
        want = stages[n]
        err = rel(got[valid], want[valid])
        finite = bool(torch.isfinite(got).all())
        good = err < (2e-3 if n == 0 else 5e-3) and finite and torch.equal(enc_len.cpu(), len_o)
        ok &= good
        print(f"encode n_layers_run={n}: rel(valid)={err:.3e} maxabs={float((got[valid]-want[valid]).abs().max()):.3e} finite={finite} len_ok={torch.equal(enc_len.cpu(), len_o)} {'OK' if good else 'FAIL'}", flush=True)
    return ok


def step_pre_encode():
    return _stage_compare([0], full_layers=1)


def step_layers():
    return _stage_compare([1, 2, 4, 8, 16], full_layers=16)


def step_ctc():
    import torch
    from oracle import gigaam_oracle as orc
    ck, eng = _engine("v2_ctc", n_layers=1)
    sd = ck["state_dict"]
    g = torch.Generator().manual_seed(3)
    ok = True
    for (B, T, lens) in [(2, 51, [51, 30]), (5, 251, [251, 250, 1, 0, 100])]:
        enc = torch.randn(B, T, 768, generator=g)
        enc_len = torch.tensor(lens, dtype=torch.int32)
        want = orc.ctc_greedy(enc.transpose(1, 2), enc_len, sd)
        ids, frames, counts = eng.greedy(enc.cuda(), enc_len.cuda())
        torch.cuda.synchronize()
        for b in range(B):
            n = int(counts[b])
            gi, gf = ids[b, :n].tolist(), frames[b, :n].tolist()
            good = gi == want[b][0] and gf == want[b][1]
            ok &= good
            print(f"ctc B={B} T={T} b={b}: n={n} want={len(want[b][0])} {'OK' if good else 'FAIL'}", flush=True)
    return ok


def step_rnnt():
    import torch
    from oracle import gigaam_oracle as orc
    ck, eng = _engine("v2_rnnt", n_layers=1)
    sd, cfg = ck["state_dict"], ck["cfg"]
    wav, wav_len = _inputs(2, 2.0, True)
    # realistic encoder activations: run the 1-layer oracle encoder
    with torch.inference_mode():
        enc, enc_len = orc.model_forward(wav, wav_len, sd, cfg)
        want = orc.rnnt_greedy(enc, enc_len, sd, 10)
    ids, frames, counts = eng.greedy(enc.transpose(1, 2).contiguous().cuda(), enc_len.cuda())
    torch.cuda.synchronize()
    ok = True
    for b in range(enc.shape[0]):
        n = int(counts[b])
        gi, gf = ids[b, :n].tolist(), frames[b, :n].tolist()
        good = gi == want[b][0] and gf == want[b][1]
        ok &= good
        print(f"rnnt b={b}: n={n} want={len(want[b][0])} first_ids={gi[:8]} want_ids={want[b][0][:8]} {'OK' if good else 'FAIL'}", flush=True)
    return ok


def step_e2e():
    import numpy as np
    import torch
    ck, eng = _engine("v2_ctc", n_layers=16)
    gold = np.load(ROOT / "tests/golden/v2_ctc_b2_2s.npz")
    wav, wav_len = _inputs(2, 2.0, True)
    mel = eng.logmel(wav.cuda())
    mel_len = torch.from_numpy(gold["mel_len"]).cuda()
    enc, enc_len = eng.encode(mel, mel_len)
    ids, frames, counts = eng.greedy(enc, enc_len)
    torch.cuda.synchronize()
    want_enc = torch.from_numpy(gold["enc"]).transpose(1, 2)
    valid = torch.arange(want_enc.shape[1])[None, :] < torch.from_numpy(gold["enc_len"])[:, None]
    err = rel(enc.cpu()[valid], want_enc[valid])
    print(f"e2e enc rel={err:.3e}", flush=True)
    ok = err < 5e-3
    margin = torch.from_numpy(gold["ctc_margin"])
    for b in range(2):
        n = int(counts[b])
        gi, gf = ids[b, :n].tolist(), frames[b, :n].tolist()
        wi, wf = gold[f"ids_{b}"].tolist(), gold[f"frames_{b}"].tolist()
        print(f"e2e b={b}: ids_equal={gi == wi} frames_equal={gf == wf} n={n}/{len(wi)} min_margin={float(margin[b][: int(gold['enc_len'][b])].min()):.4f}", flush=True)
    return ok


def step_timing():
    import torch
    ck, eng = _engine("v2_ctc", n_layers=16)
    wav, wav_len = _inputs(64, 10.0, False)
    wav = wav.cuda()
    mel_len = torch.full((64,), eng.logmel_frames(wav.shape[1]), dtype=torch.int64, device="cuda")

    def fwd():
        mel = eng.logmel(wav)
        enc, enc_len = eng.encode(mel, mel_len)
        return eng.greedy(enc, enc_len)

    for _ in range(3):
        fwd()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    ev[0].record()
    for _ in range(5):
        fwd()
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 5
    print(f"timing v2_ctc B=64x10s eager: {ms:.3f} ms/batch -> {64 / ms * 1e3:.1f} utt/s, RTFx {640 / ms * 1e3:.0f}", flush=True)
    # stage split
    mel = eng.logmel(wav)
    torch.cuda.synchronize()
    for name, fn in [("logmel", lambda: eng.logmel(wav)), ("encode", lambda: eng.encode(mel, mel_len)),
                     ("pre_encode_only", lambda: eng.encode(mel, mel_len, n_layers_run=0)),
                     ("one_layer", lambda: eng.encode(mel, mel_len, n_layers_run=1))]:
        fn()
        torch.cuda.synchronize()
        ev[2].record()
        for _ in range(3):
            fn()
        ev[3].record()
        torch.cuda.synchronize()
        print(f"  {name}: {ev[2].elapsed_time(ev[3]) / 3:.3f} ms", flush=True)
    return True


def main(argv):
    if len(argv) >= 2 and argv[0] == "--step":
        fn = globals()["step_" + argv[1]]
        ok = fn()
        print("STEP", argv[1], "PASS" if ok else "FAIL", flush=True)
        return 0 if ok else 1
    steps = argv if argv else STEPS
    OUT.mkdir(exist_ok=True)
    summary = {}
    for st in steps:
        t0 = time.time()
        log = OUT / f"check_{st}.log"
        with open(log, "w") as f:
            try:
                rc = subprocess.run([sys.executable, __file__, "--step", st], stdout=f, stderr=subprocess.STDOUT,
                                    timeout=int(os.environ.get("STEP_TIMEOUT", "420"))).returncode
            except subprocess.TimeoutExpired:
                rc = -999
        summary[st] = {"rc": rc, "sec": round(time.time() - t0, 1)}
        tail = log.read_text().splitlines()[-40:]
        print(f"===== {st}: rc={rc} ({summary[st]['sec']} s)")
        print("\n".join(tail), flush=True)
    (OUT / "check_summary.json").write_text(json.dumps(summary, indent=1))
    print(json.dumps(summary))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
