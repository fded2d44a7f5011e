# Round 2, session 3: the generalised host pipeline (tests), bench line, e2e sweep with the measuring controller,
# ncu summaries of the secondary configs' kernels (K2/K3, K5, K6).
mkdir -p gpurun_out
lscpu | head -25 > gpurun_out/r2s3_lscpu.txt; nproc >> gpurun_out/r2s3_lscpu.txt; free -g >> gpurun_out/r2s3_lscpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2s3_pytest.txt
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2s3_bench.json 2> gpurun_out/r2s3_bench.err; echo bench rc=$?
tail -5 gpurun_out/r2s3_bench.err
timeout 600 python scripts/e2e_sweep.py 16384 2>&1 | tee gpurun_out/r2s3_e2e_sweep.txt | tail -30
for k in "svc_kernel_window|svc_proba_window|svc_pack" "crf_smooth" "gnofix_kernel|gnofix_rank|gnofix_diff|gnofix_apply"; do
  tag=$(echo $k | cut -c1-3)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$k" -c 6 -o gpurun_out/r2s3_$tag python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --haps 8192 --covrsk-haps 2048 > gpurun_out/r2s3_ncu_$tag.log 2>&1; echo ncu $tag rc=$?
done
cut -c1-800 gpurun_out/r2s3_bench.json
