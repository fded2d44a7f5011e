# Round 2, session 7: K4a with per-cell records: parity, bench, ncu of K4a and of the CovRSK production kernel.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gbt_gpu.py tests/test_fullsize_gpu.py tests/test_edge_gpu.py tests/test_pipeline_gpu.py tests/test_gnofix_gpu.py tests/test_cli_gpu.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2s7_pytest.txt
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2s7_bench.json 2> gpurun_out/r2s7_bench.err; echo bench rc=$?
tail -5 gpurun_out/r2s7_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2s7_bench.json'))
print(d['value'], d['ms_per_step'], {k:round(v['ms'],3) for k,v in d['kernels'].items()}, d['e2e']['value'])
print({k:(v.get('haplotypes_per_s') or v.get('individuals_per_s') or v) for k,v in d['configs'].items()})
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gbt_rank_tile" -c 2 -o gpurun_out/r2s7_k4a python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-configs > gpurun_out/r2s7_ncu_a.log 2>&1; echo ncu rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"svc_kernel_csa|svc_pair|svc_couple" -c 6 -o gpurun_out/r2s7_k2 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --haps 8192 > gpurun_out/r2s7_ncu_b.log 2>&1; echo ncu rc=$?
