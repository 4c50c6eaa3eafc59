#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/ln_probe.py > gpurun_out/r2d_probe.log 2>&1
cat gpurun_out/r2d_probe.log | tail -20
if ! grep -q "K=3072 ln3 " gpurun_out/r2d_probe.log; then echo "probe incomplete: stopping"; nvidia-smi > gpurun_out/r2d_smi.txt 2>&1; exit 0; fi
timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gemm or encoder_stagewise or end_to_end or determinism or config2" 2>&1 | tail -15 > gpurun_out/r2d_pytest.log
tail -5 gpurun_out/r2d_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-c4 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2d_bench.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'])
print(d['roofline']['classes_ms_per_step'])
PY
tail -3 gpurun_out/r2d_bench.err
