set -x
free -g | head -2; nproc; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1_first.json 2> gpurun_out/bench_r1_first.err; echo rc=$?
tail -3 gpurun_out/bench_r1_first.err; cat gpurun_out/bench_r1_first.json
# launch list of a short run (smaller N to keep ncu replays short)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --haps 8192 --e2e-haps 1024 --e2e-steps 1 --no-cpu > gpurun_out/ncu_launch.log 2>&1; echo rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'lr_tc_kernel|gbt_smooth_kernel' -c 4 -o gpurun_out/prof_r1 python bench.py --steps 1 --warmup 1 --haps 8192 --e2e-haps 1024 --e2e-steps 1 --no-cpu > gpurun_out/ncu_full.log 2>&1; echo rc=$?
ls -la gpurun_out
