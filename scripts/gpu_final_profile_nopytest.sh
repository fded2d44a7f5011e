# Round-end evidence: full GPU test suite, bench line, reference-arm line, launch list of the same command,
# full ncu capture of the hot kernels at the bench size.
mkdir -p gpurun_out
echo "(pytest -m gpu run separately)"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo bench rc=$?
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err; echo ref rc=$?
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/clocks_after.csv
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lr_tc_kernel|gbt_smooth|unpack" --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 3 --warmup 1 --no-cpu --e2e-haps 2048 --e2e-steps 1 > gpurun_out/ncu_launch_final.log 2>&1; echo launches rc=$?
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"lr_tc_kernel|gbt_smooth_rank|unpack_kernel" -c 3 -o gpurun_out/prof_final python bench.py --steps 1 --warmup 0 --no-cpu --e2e-haps 2048 --e2e-steps 1 > gpurun_out/ncu_full_final.log 2>&1; echo full rc=$?
cat gpurun_out/bench_final.json | cut -c1-600
