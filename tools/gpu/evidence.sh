#!/bin/bash
# evidence pass on one B200: full GPU test suite, both bench arms, every BASELINE config (+ the ragged c2 batch), the ncu
# launch list of a step and `ncu --set full` captures of the attention kernel (c2 and c5 shapes)
mkdir -p gpurun_out
P=gpurun_out/r2p
timeout -k 5 1200 python -m pytest tests -m gpu -q -s 2>&1 | tail -70 > ${P}_pytest.log
tail -3 ${P}_pytest.log
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1; tail -1 ${P}_smoke.log
timeout -k 5 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > ${P}_bench_reference.json 2> ${P}_bench_reference.err
timeout -k 5 600 python bench.py --gpus 1 --steps 40 --warmup 5 > ${P}_bench_n1.json 2> ${P}_bench_n1.err
timeout -k 5 600 python tools/bench_configs.py > ${P}_configs.jsonl 2> ${P}_configs.err
timeout -k 5 200 python tools/attn_probe.py > ${P}_attn_probe.txt 2>&1
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file ${P}_launches.csv \
    python tools/profile_step.py > ${P}_ncu_list.log 2>&1
timeout -k 5 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"attention_kernel|subsample_conv1|unpack_rows|ln_rope" -c 4 \
    -o ${P}_attn_c2 -f python tools/profile_step.py --layers 1 > ${P}_ncu_c2.log 2>&1
timeout -k 5 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_kernel -c 1 \
    -o ${P}_attn_c5 -f python tools/profile_step.py --layers 1 --model v2_ssl --batch 32 --seconds 25 > ${P}_ncu_c5.log 2>&1
python - <<'PY'
import json
r=json.load(open('gpurun_out/r2p_bench_reference.json')); d=json.load(open('gpurun_out/r2p_bench_n1.json'))
print("reference", r['value'], r['cpu_baseline'])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'], d['gpu_launches'])
print(d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['step_frac'], d['roofline']['classes_ms_per_step'])
print(d.get('strong_scaling_c4'))
for l in open('gpurun_out/r2p_configs.jsonl'):
    c=json.loads(l); print(c['config'], c['ms_per_batch'], c['utt_per_s'], c.get('audio_fraction'), c['clocks'].get('sm_mhz'), c['classes_ms'])
print(open('gpurun_out/r2p_attn_probe.txt').read())
PY
