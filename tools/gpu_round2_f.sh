#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
  GAM_ZIGZAG_OFF=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-c4 --no-cpu-baseline > gpurun_out/r2f_off_$rep.json 2> gpurun_out/r2f_off_$rep.err
  timeout 300 python bench.py --steps 30 --warmup 5 --no-c4 --no-cpu-baseline > gpurun_out/r2f_on_$rep.json 2> gpurun_out/r2f_on_$rep.err
done
python - <<'PY'
import json
for n in ("off_1","on_1","off_2","on_2"):
    try:
        d=json.load(open(f'gpurun_out/r2f_{n}.json'))
        print(n, round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['roofline']['classes_ms_per_step'])
    except Exception as e:
        print(n, "failed", e)
PY
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gemm or stagewise or end_to_end or determinism or batch_vs_single or v1_rel_pos or v3_frontend" 2>&1 | tail -4
