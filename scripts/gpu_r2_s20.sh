# Round 2, session 20 (evidence on the last tree): full GPU suite, bench line, reference arm, smoke.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2s20_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2s20_smoke.txt
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2s20_bench.json 2> gpurun_out/r2s20_bench.err; echo bench rc=$?
tail -3 gpurun_out/r2s20_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2s20_bench_reference_arm.json 2> gpurun_out/r2s20_bench_reference_arm.err; echo ref rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2s20_bench.json'))
print(d['value'], d['ms_per_step'], {k:round(v['ms'],3) for k,v in d['kernels'].items()}, d['e2e']['value'], d['e2e']['packed_input']['value'], d['e2e']['plugin_pageable']['value'])
print({k:(v.get('haplotypes_per_s') or v.get('individuals_per_s') or v) for k,v in d['configs'].items()})
print(d['roofline']['frac'], d['cpu_baseline'])
PY
