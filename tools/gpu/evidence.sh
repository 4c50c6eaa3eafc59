#!/bin/bash
# round-2 evidence pass on one B200: full GPU test suite, both bench arms, launch list, ncu --set full of the dominant kernels
mkdir -p gpurun_out
P=gpurun_out/r2f
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | tail -60 > ${P}_pytest.log
tail -3 ${P}_pytest.log
timeout 900 python bench.py --steps 40 --warmup 5 > ${P}_bench_n1.json 2> ${P}_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > ${P}_bench_reference.json 2> ${P}_bench_reference.err
timeout 600 python tools/bench_configs.py > ${P}_configs.jsonl 2> ${P}_configs.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file ${P}_launches.csv \
    python tools/profile_step.py > ${P}_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm2_f16_tn_kernel -c 24 \
    -o ${P}_gemm2 python tools/profile_step.py --layers 1 > ${P}_ncu_gemm2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"attention|dwconv|subsample_conv1|ctc_|ln_f16|frames_split|mel_log" -c 12 \
    -o ${P}_misc python tools/profile_step.py --layers 1 > ${P}_ncu_misc.log 2>&1
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench_n1.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'])
print(d['roofline']['classes_ms_per_step'])
print(d.get('strong_scaling_c4'))
PY
