#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/ln_probe.py > gpurun_out/r2b_probe.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2_f16_tn_kernel -s 10 -c 2 -o gpurun_out/r2b_ln python tools/ln_probe.py > gpurun_out/r2b_ncu.log 2>&1
cat gpurun_out/r2b_probe.log
