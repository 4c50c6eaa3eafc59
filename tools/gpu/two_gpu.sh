#!/bin/bash
# 2-GPU pass: NCCL test of the sharded path + bench at N = 2 (the driver's own launch line)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2g_gpus.txt
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -s 2>&1 | tail -15 > gpurun_out/r2g_pytest.log
tail -5 gpurun_out/r2g_pytest.log
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err
tail -5 gpurun_out/r2g_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2g_bench_n2.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'])
print(d.get('strong_scaling_c4'))
PY
