# Round-1 (session 3) GPU pass: new parity tests, K4 node-layout A/B, e2e packed-fraction sweep.
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; lscpu | grep -E "Model name|Socket|NUMA node\(s\)|^CPU\(s\)" >> gpurun_out/host.txt; free -g | head -2 >> gpurun_out/host.txt
timeout 600 python -m pytest tests/test_pack_gpu.py tests/test_gbt_gpu.py tests/test_fullsize_gpu.py -x -q 2>&1 | tail -8
for v in 1 2; do
  GNX_GBT_VARIANT=$v timeout 600 python bench.py --steps 5 --warmup 3 ${BENCH_EXTRA} > gpurun_out/bench_var$v.json 2> gpurun_out/bench_var$v.err; echo "variant $v rc=$?"
done
timeout 600 python scripts/e2e_sweep.py > gpurun_out/e2e_sweep.txt 2>&1; echo sweep rc=$?
python - <<'PY'
import json
for v in (1,2):
    try:
        d=json.load(open('gpurun_out/bench_var%d.json'%v))
        print('var',v,'value',round(d['value']),'ms',round(d['ms_per_step'],2),'K1',round(d['kernels']['K1_lr_tc_kernel']['ms'],2),'K4',round(d['kernels']['K4_gbt_smooth_kernel']['ms'],2),'e2e',d['e2e'])
        print('  cpu',d.get('cpu_baseline',{}).get('value'),d['clocks'])
    except Exception as e: print('var',v,'failed',e)
PY
cat gpurun_out/e2e_sweep.txt | tail -30
