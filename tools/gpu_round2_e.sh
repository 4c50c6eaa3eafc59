#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/gemm_probe.py > gpurun_out/r2e_probe.log 2>&1
cat gpurun_out/r2e_probe.log | tail -8
timeout 1200 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/r2e_pytest.log
tail -4 gpurun_out/r2e_pytest.log
timeout 400 python bench.py --steps 30 --warmup 5 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2e_bench.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'])
print(d['roofline']['classes_ms_per_step'])
print(d.get('strong_scaling_c4'))
PY
tail -3 gpurun_out/r2e_bench.err
