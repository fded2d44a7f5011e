# Round 2, session 5: K4a v3 + CovRSK production kernel: parity suite, bench line.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/r2s5_pytest.txt
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2s5_bench.json 2> gpurun_out/r2s5_bench.err; echo bench rc=$?
tail -5 gpurun_out/r2s5_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2s5_bench.json'))
print(d['value'], d['ms_per_step'], {k:round(v['ms'],3) for k,v in d['kernels'].items()}, d['e2e']['value'])
print({k:(v.get('haplotypes_per_s') or v.get('individuals_per_s') or v) for k,v in d['configs'].items()})
PY
