# Round 2, session 1: parity of the new K4 tile kernel, A/B timing against the row kernel, ncu of the tile kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gbt_gpu.py tests/test_gnofix_gpu.py tests/test_fullsize_gpu.py tests/test_cli_gpu.py -x -q 2>&1 | tail -5 | tee gpurun_out/r2s1_pytest.txt
( timeout 300 python scripts/k4_probe.py 20000 14,16
  GNX_GBT_TOPW=3 timeout 300 python scripts/k4_probe.py 20000 16
  timeout 300 python scripts/k4_probe.py 50000 14,16 ) 2>&1 | tee gpurun_out/r2s1_k4_probe.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gbt_" --csv --log-file gpurun_out/r2s1_k4_launches.csv python scripts/k4_probe.py 20000 16 > gpurun_out/r2s1_ncu_launch.log 2>&1; echo launches rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gbt_smooth_tile|gbt_rank_tile" -c 2 -o gpurun_out/r2s1_k4_tile python scripts/k4_probe.py 20000 16 > gpurun_out/r2s1_ncu_full.log 2>&1; echo full rc=$?
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2s1_bench.json 2> gpurun_out/r2s1_bench.err; echo bench rc=$?
cut -c1-400 gpurun_out/r2s1_bench.json
