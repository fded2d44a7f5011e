# Round-end evidence: bench line, launch list of the same command, full ncu capture of both kernels at the bench size.
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo bench rc=$?
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/clocks_after.csv
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lr_tc_kernel|gbt_smooth|unpack" --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 3 --warmup 1 --no-cpu > gpurun_out/ncu_launch_final.log 2>&1; echo launches rc=$?
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"lr_tc_kernel|gbt_smooth_rank" -c 2 -o gpurun_out/prof_final python bench.py --steps 1 --warmup 0 --no-cpu --e2e-haps 512 --e2e-steps 1 > gpurun_out/ncu_full_final.log 2>&1; echo full rc=$?
cat gpurun_out/bench_final.json | cut -c1-400
