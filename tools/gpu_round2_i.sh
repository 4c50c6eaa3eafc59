#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "rnnt or config3 or config4 or v3" 2>&1 | tail -12 > gpurun_out/r2i_pytest.log
tail -6 gpurun_out/r2i_pytest.log
timeout 600 python tools/bench_configs.py c1 c2 c3 c4 c5 v1c2 > gpurun_out/r2i_configs.jsonl 2> gpurun_out/r2i_configs.err
python - <<'PY'
import json
for l in open('gpurun_out/r2i_configs.jsonl'):
    d=json.loads(l); print(d['config'], d['ms_per_batch'], d['utt_per_s'], d.get('tokens_per_frame'), d['clocks']['sm_mhz'], d['clocks']['reasons'], list(d['classes_ms'].items())[:4])
PY
