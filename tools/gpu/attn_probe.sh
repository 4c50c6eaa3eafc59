#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "attention" 2>&1 | tail -5
PROBE_POLY=0,4,2,1 PROBE_STAGE=1,0 timeout -k 5 300 python tools/attn_probe.py 2>&1 | tee gpurun_out/r2n_attn_probe.txt | tail -30
