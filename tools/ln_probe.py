"""Timing probe of the fused-LayerNorm residual GEMM (development aid): plain residual epilogue vs LN modes, with the
wait / second pass switched off in turn.  python tools/ln_probe.py [M]"""
import ctypes as C
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from gigaam_b200 import synthetic  # noqa: E402
from gigaam_b200.engine import Engine  # noqa: E402

dev = torch.device("cuda", 0)
ck = synthetic.synthetic_checkpoint("v2_ctc", seed=0, n_layers=1)
eng = Engine(ck["cfg"], ck["state_dict"], dev)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 16064
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for K in (768, 3072):
    A = (torch.randn(M, K) * 0.5).half().to(dev)
    W = (torch.randn(768, K) / K ** 0.5).half().to(dev)
    bias = torch.randn(768).to(dev)
    x = torch.randn(M, 768).to(dev)
    g1 = torch.ones(768, device=dev)
    b1 = torch.zeros(768, device=dev)
    out16 = torch.zeros(M, 768, dtype=torch.float16, device=dev)
    rope16 = torch.zeros(M, 768, dtype=torch.float16, device=dev)
    ws = torch.empty(2 * ((M * 48 + 1023) // 1024 * 1024) + 2 * ((M + 255) // 256) * 32 + 1024, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def run(kind):
        if kind == "plain":
            return eng.lib.gam_test_gemm(eng.handle, 3, A.data_ptr(), W.data_ptr(), bias.data_ptr(), x.data_ptr(), x.data_ptr(), M, 768, K,
                                         768, 0.5, st)
        mode = {"ln1": 1, "ln2": 2, "ln3": 3, "ln3last": 3, "ln1_nowait": 1 | 16, "ln1_nopass2": 1 | 32, "ln1_nowait_nopass2": 1 | 48}[kind]
        two = kind == "ln3"
        return eng.lib.gam_test_gemm_ln(eng.handle, mode, A.data_ptr(), W.data_ptr(), bias.data_ptr(), x.data_ptr(), g1.data_ptr(),
                                        b1.data_ptr(), g1.data_ptr() if two else None, b1.data_ptr() if two else None, out16.data_ptr(),
                                        rope16.data_ptr(), x.data_ptr(), M, K, 251, 0.5, ws.data_ptr(), ws.numel(), st)

    for kind in ("plain", "ln1", "ln2", "ln3", "ln3last", "ln1_nowait", "ln1_nopass2", "ln1_nowait_nopass2"):
        ts = []
        for it in range(6):
            flush.fill_(it)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = run(kind)
            e1.record()
            torch.cuda.synchronize()
            assert rc == 0, kind
            ts.append(e0.elapsed_time(e1) * 1e3)
        print(f"M={M} K={K} {kind:22s} {min(ts[1:]):8.1f} us (median {sorted(ts[1:])[2]:.1f})", flush=True)
