"""One step of the bench workload between cudaProfilerStart/Stop, for ncu (--profile-from-start off):

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py
    ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_f16_tn -c 3 \
        -o gpurun_out/prof_gemm python tools/profile_step.py
"""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

import torch  # noqa: E402

import gigaam_b200 as gigaam  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="v2_ctc")
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--seconds", type=float, default=10.0)
ap.add_argument("--layers", type=int, default=None)
args = ap.parse_args()

dev = torch.device("cuda", 0)
ck = gigaam.synthetic_checkpoint(args.model, n_layers=args.layers)
model = gigaam.load_model(args.model, device=dev, checkpoint=ck)
wav, wav_len = gigaam.synthetic_audio(args.batch, args.seconds, seed=1234)
wav, wav_len = wav.to(dev), wav_len.to(dev)


def step():
    enc, enc_len = model(wav, wav_len)
    return model.decoding.decode_device(model.head, enc, enc_len) if hasattr(model, "head") else (enc, enc_len)


step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one step; launches so far:", model._get_engine().launch_count())
