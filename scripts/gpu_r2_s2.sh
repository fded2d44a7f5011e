# Round 2, session 2: full GPU suite on the LUT rank pass, bench line with configs, launch list, full ncu of K1/K4a/K4b at the bench size.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2s2_pytest.txt
timeout 300 python scripts/k4_probe.py 50000 14,16 2>&1 | tee gpurun_out/r2s2_k4_probe.txt
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2s2_bench.json 2> gpurun_out/r2s2_bench.err; echo bench rc=$?
tail -5 gpurun_out/r2s2_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2s2_launches.csv python bench.py --steps 3 --warmup 1 --no-cpu --no-configs --e2e-haps 2048 --e2e-steps 1 > gpurun_out/r2s2_ncu_launch.log 2>&1; echo launches rc=$?
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"lr_tc_kernel|gbt_rank_tile|gbt_smooth_tile" -c 6 -o gpurun_out/r2s2_k1k4 python bench.py --steps 1 --warmup 1 --no-cpu --no-configs --e2e-haps 2048 --e2e-steps 1 > gpurun_out/r2s2_ncu_full.log 2>&1; echo full rc=$?
cut -c1-1500 gpurun_out/r2s2_bench.json
