# Round 2, session 4: K4a rewrite (coalesced tiles, float-key cells): parity suite, probe, bench, ncu of K4a.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2s4_pytest.txt
timeout 300 python scripts/k4_probe.py 50000 14,16 2>&1 | tee gpurun_out/r2s4_k4_probe.txt
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2s4_bench.json 2> gpurun_out/r2s4_bench.err; echo bench rc=$?
tail -5 gpurun_out/r2s4_bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gbt_rank_tile" -c 2 -o gpurun_out/r2s4_k4a python bench.py --steps 1 --warmup 1 --no-cpu --no-configs --no-e2e > gpurun_out/r2s4_ncu.log 2>&1; echo ncu rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2s4_bench.json'))
print(d['value'], d['ms_per_step'], {k:v['ms'] for k,v in d['kernels'].items()}, d['e2e']['value'], d['e2e']['packed_input']['value'])
PY
