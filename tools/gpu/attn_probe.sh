#!/bin/bash
# attention kernel: unit tests (short timeout: a protocol bug would hang) and isolated timing at the BASELINE shapes
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "attention" 2>&1 | tail -5
timeout -k 5 200 python tools/attn_probe.py 2>&1 | tee gpurun_out/attn_probe.txt | tail -8
