#!/bin/bash
# varlen (packed-row) bring-up: the kernel-level tests first (short timeout: a protocol bug would hang), then the whole GPU suite
mkdir -p gpurun_out
P=gpurun_out/r2k
timeout -k 5 240 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "attention" 2>&1 | tail -25 > ${P}_attention.log
cat ${P}_attention.log | tail -12
timeout -k 5 900 python -m pytest tests -q -m gpu -s 2>&1 | tail -80 > ${P}_pytest.log
grep -v "^$" ${P}_pytest.log | tail -45
timeout -k 5 300 python bench.py --gpus 1 --steps 20 --warmup 5 > ${P}_bench_n1.json 2> ${P}_bench_n1.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2k_bench_n1.json'))
    print(d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'], d['gpu_launches'])
    print(d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['step_frac'], d['roofline']['classes_ms_per_step'])
except Exception as e:
    print("bench failed", e); print(open('gpurun_out/r2k_bench_n1.err').read()[-2000:])
PY
