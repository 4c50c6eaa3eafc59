#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/ln_probe.py > gpurun_out/r2c_probe.log 2>&1
cat gpurun_out/r2c_probe.log
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/r2c_pytest.log
tail -5 gpurun_out/r2c_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c_bench.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'])
print(d['roofline']['classes_ms_per_step'])
PY
