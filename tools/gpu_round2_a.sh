#!/bin/bash
# first GPU pass of round 2: unit + parity tests, then a bench line
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
nproc > gpurun_out/r2a_nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -80 > gpurun_out/r2a_pytest.log
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -3 gpurun_out/r2a_pytest.log
cat gpurun_out/r2a_bench.json | head -c 3000
