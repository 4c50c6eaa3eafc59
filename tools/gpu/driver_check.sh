#!/bin/bash
# what the driver runs at round end, in its order: GPU tests, smoke(), reference arm, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2q_pytest.log
cat gpurun_out/r2q_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2q_smoke.log 2>&1
tail -2 gpurun_out/r2q_smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 > gpurun_out/r2q_bench_reference.json 2> gpurun_out/r2q_bench_reference.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2q_bench_n1.json 2> gpurun_out/r2q_bench_n1.err
python - <<'PY'
import json
r=json.load(open('gpurun_out/r2q_bench_reference.json')); d=json.load(open('gpurun_out/r2q_bench_n1.json'))
print("reference", r['value'], r['cpu_baseline'])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'], d['gpu_launches'])
print(d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['step_frac'], d['roofline']['classes_ms_per_step'])
PY
