# Round 2, session 14: K4b with pair words (root feature load shared by the two adjacent windows of a warp): parity, bench.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gbt_gpu.py tests/test_fullsize_gpu.py tests/test_edge_gpu.py tests/test_pipeline_gpu.py tests/test_cli_gpu.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2s14_pytest.txt
timeout 1200 python bench.py --steps 5 --warmup 3 --no-e2e > gpurun_out/r2s14_bench.json 2> gpurun_out/r2s14_bench.err; echo bench rc=$?
tail -5 gpurun_out/r2s14_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2s14_bench.json'))
print(d['value'], d['ms_per_step'], {k:round(v['ms'],3) for k,v in d['kernels'].items()})
print({k:(v.get('haplotypes_per_s') or v.get('individuals_per_s') or v) for k,v in d['configs'].items()})
print(d.get('parity'))
PY
