#!/bin/bash
# ncu --set full (with source counters) of the attention kernel at the c2 (T' = 251) and c5 (T' = 626) shapes, one layer each
mkdir -p gpurun_out
P=gpurun_out/r2m
timeout -k 5 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_kernel -c 1 \
    -o ${P}_attn_c2 -f python tools/profile_step.py --layers 1 > ${P}_ncu_c2.log 2>&1
tail -2 ${P}_ncu_c2.log
timeout -k 5 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_kernel -c 1 \
    -o ${P}_attn_c5 -f python tools/profile_step.py --layers 1 --model v2_ssl --batch 32 --seconds 25 > ${P}_ncu_c5.log 2>&1
tail -2 ${P}_ncu_c5.log
ls -la gpurun_out/*.ncu-rep
