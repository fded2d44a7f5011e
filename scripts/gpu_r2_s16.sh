# Round 2, session 16 (evidence): full GPU suite, bench line with e2e + configs, launch list, full ncu of K1 / K4a / K4b at the bench size.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2s16_pytest.txt
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2s16_bench.json 2> gpurun_out/r2s16_bench.err; echo bench rc=$?
tail -5 gpurun_out/r2s16_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2s16_bench_reference_arm.json 2> gpurun_out/r2s16_bench_reference_arm.err; echo ref rc=$?
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2s16_launches.csv python bench.py --steps 3 --warmup 1 --no-cpu --no-configs --e2e-haps 2048 --e2e-steps 1 > gpurun_out/r2s16_ncu_launch.log 2>&1; echo launches rc=$?
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"lr_tc_kernel|gbt_rank_tile|gbt_smooth_tile" -c 6 -o gpurun_out/r2s16_k1k4 python bench.py --steps 1 --warmup 1 --no-cpu --no-configs --e2e-haps 2048 --e2e-steps 1 > gpurun_out/r2s16_ncu_full.log 2>&1; echo full rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2s16_bench.json'))
print(d['value'], d['ms_per_step'], {k:round(v['ms'],3) for k,v in d['kernels'].items()}, d['e2e']['value'], d['e2e']['packed_input']['value'], d['e2e']['plugin_pageable']['value'])
print({k:(v.get('haplotypes_per_s') or v.get('individuals_per_s') or v) for k,v in d['configs'].items()})
print(d['roofline']['frac'], d['cpu_baseline'])
PY
