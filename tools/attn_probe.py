"""Isolated timing of the attention kernel at the BASELINE shapes (CUDA events, L2 flushed between launches).

    python tools/attn_probe.py            # c2 (64 x 251), c3 (32 x 376), c5 slice (32 x 626)
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

import torch  # noqa: E402

import gigaam_b200 as gigaam  # noqa: E402
from gigaam_b200.engine import Engine  # noqa: E402

dev = torch.device("cuda", 0)
ck = gigaam.synthetic_checkpoint("v2_ctc", n_layers=1)
eng = Engine(ck["cfg"], ck["state_dict"], dev)
stream = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(B, T, reps=10):
    d, H, dk = 768, 16, 48
    g = torch.Generator().manual_seed(B + T)
    qkv = torch.randn(B * T, 3 * d, generator=g).half().to(dev)
    out = torch.zeros(B * T, d, dtype=torch.float16, device=dev)
    klen = torch.full((B,), T, dtype=torch.int32, device=dev)
    cu = (torch.arange(B + 1, dtype=torch.int32) * T).to(dev)

    def launch():
        rc = eng.lib.gam_test_attention_varlen(eng.handle, qkv.data_ptr(), None, klen.data_ptr(), cu.data_ptr(), out.data_ptr(), B, T, B * T, stream)
        assert rc == 0
    launch()
    torch.cuda.synchronize()
    x = qkv[: 4 * T].float().view(4, T, 3, H, dk)
    q, k, v = (x[:, :, i].transpose(1, 2) for i in range(3))
    want = (torch.softmax(q @ k.transpose(-1, -2) / dk ** 0.5, -1) @ v).transpose(1, 2).reshape(4 * T, d)
    err = float((out[: 4 * T].float() - want).norm() / want.norm())
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launch()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], err


if __name__ == "__main__":
    for B, T in [(64, 251), (32, 376), (32, 626)]:
        us, err = run(B, T)
        print(f"B={B} T={T}: {us:8.1f} us per launch, rel err vs fp32 softmax {err:.2e}", flush=True)
