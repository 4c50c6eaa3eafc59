#!/bin/bash
# single-sweep attention kernel bring-up: kernel-level tests under a short timeout, encoder-level parity, then the bench
mkdir -p gpurun_out
P=gpurun_out/r2l
timeout -k 5 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "attention" 2>&1 | tail -25 > ${P}_attention.log
tail -12 ${P}_attention.log
timeout -k 5 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "stagewise or varlen_ragged or config5 or 30s" 2>&1 | tail -40 > ${P}_encoder.log
grep -v "^$" ${P}_encoder.log | tail -25
timeout -k 5 300 python bench.py --gpus 1 --steps 20 --warmup 5 > ${P}_bench_n1.json 2> ${P}_bench_n1.err
timeout -k 5 300 python tools/bench_configs.py c3 c5 > ${P}_configs.jsonl 2> ${P}_configs.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2l_bench_n1.json'))
    print(d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'], d['gpu_launches'])
    print(d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['step_frac'], d['roofline']['classes_ms_per_step'])
except Exception as e:
    print("bench failed", e); print(open('gpurun_out/r2l_bench_n1.err').read()[-2000:])
for l in open('gpurun_out/r2l_configs.jsonl'):
    try:
        c=json.loads(l); print(c["config"], c["ms_per_batch"], c["utt_per_s"], c["classes_ms"])
    except Exception as e: print(l[:200])
PY
timeout -k 5 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_kernel -c 1 \
    -o gpurun_out/r2m_attn_c2 -f python tools/profile_step.py --layers 1 > gpurun_out/r2m_ncu_c2.log 2>&1
tail -1 gpurun_out/r2m_ncu_c2.log
