# Round 2, session 6: K4a without barriers, leaner CovRSK loop: parity of the touched kernels, bench, ncu of K4a + K2.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gbt_gpu.py tests/test_svc_gpu.py tests/test_fullsize_gpu.py tests/test_edge_gpu.py tests/test_pipeline_gpu.py tests/test_gnofix_gpu.py tests/test_lr_gpu.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2s6_pytest.txt
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2s6_bench.json 2> gpurun_out/r2s6_bench.err; echo bench rc=$?
tail -5 gpurun_out/r2s6_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2s6_bench.json'))
print(d['value'], d['ms_per_step'], {k:round(v['ms'],3) for k,v in d['kernels'].items()}, d['e2e']['value'])
print({k:(v.get('haplotypes_per_s') or v.get('individuals_per_s') or v) for k,v in d['configs'].items()})
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gbt_rank_tile|svc_kernel_csa" -c 4 -o gpurun_out/r2s6_k4a_k2 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --haps 50000 --covrsk-haps 8192 > gpurun_out/r2s6_ncu.log 2>&1; echo ncu rc=$?
