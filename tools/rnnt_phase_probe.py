"""Phase timing of the RNN-T cluster kernel (development aid).  Needs a library built with -DGAM_RNNT_DBG:

    GIGAAM_B200_LIB=tools/_dbg/libgigaam_dbg.so python tools/rnnt_phase_probe.py v2_rnnt 32 15

Prints clock64() cycles accumulated by thread 0 of CTA 0 of cluster 0 per phase of rnnt_cluster_kernel."""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

import torch  # noqa: E402

import gigaam_b200 as gigaam  # noqa: E402

NAMES = ["gates (LSTM dot products)", "cell update + h broadcast", "cluster.sync after LSTM / pred", "pred projection + pg broadcast",
         "warp reduce + argmax push", "cluster.sync after joint", "setup / group init / tail", "rounds", "LSTM rounds",
         "loop head (+ first-round setup)", "read 16 partial argmax", "decision", "prefetch issue", "hid4 + sync", "class rows",
         "wbest sync"]


def main(model_name="v2_rnnt", B=32, sec=15.0):
    dev = torch.device("cuda", 0)
    model = gigaam.load_model(model_name, device=dev, synthetic=True)
    eng = model._get_engine()
    wav, wav_len = gigaam.synthetic_audio(int(B), float(sec), seed=1234)
    enc, enc_len = model(wav.to(dev), wav_len.to(dev))
    for _ in range(2):
        ids, frames, counts = model.decoding.decode_device(model.head, enc, enc_len)
    torch.cuda.synchronize()
    out = (C.c_longlong * 16)()
    assert eng.lib.gam_rnnt_debug_read(out) == 0
    cyc = list(out)[:16]
    rounds, lstm_rounds = cyc[7], cyc[8]
    total = sum(cyc[:7]) + sum(cyc[9:16])
    print(f"{model_name} B={B} {sec}s: tokens/frame {float(counts.sum()) / float(enc_len.sum()):.3f}; cluster 0: {rounds} rounds, "
          f"{lstm_rounds} with an LSTM step; {total} cycles total = {total / max(rounds, 1):.0f} per round")
    for i in (0, 1, 2, 3, 13, 14, 15, 4, 5, 10, 11, 12, 9, 6):
        print(f"  {NAMES[i]:45s} {cyc[i]:12d} cycles  {100.0 * cyc[i] / total:5.1f} %")


if __name__ == "__main__":
    main(*sys.argv[1:])
