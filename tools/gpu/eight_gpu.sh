#!/bin/bash
# 8-GPU pass: the driver's launch line at N = 8 (weak c2 + strong c4 legs)
mkdir -p gpurun_out
NCCL_DEBUG=WARN timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/r2h_bench_n8.json 2> gpurun_out/r2h_bench_n8.err
tail -4 gpurun_out/r2h_bench_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2h_bench_n8.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'])
print(d.get('strong_scaling_c4'))
PY
