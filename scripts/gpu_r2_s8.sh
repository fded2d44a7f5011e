# Round 2, session 8: K6 with the block walk + 4 teams, CovRSK 3 vs 4 CTAs per SM: parity of the touched kernels, A/B timings.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gnofix_gpu.py tests/test_svc_gpu.py tests/test_pipeline_gpu.py tests/test_cli_gpu.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2s8_pytest.txt
for t in 4 3 2; do echo "GNX_GNOFIX_TEAMS=$t"; GNX_GNOFIX_TEAMS=$t timeout 600 python scripts/gnofix_probe.py 10000 2>&1 | tail -1; done | tee gpurun_out/r2s8_gnofix_ab.txt
for c in 4 3; do GNX_SVC_CTAS=$c timeout 600 python scripts/svc_ab.py 8192 2>&1 | tail -1; done | tee gpurun_out/r2s8_svc_ab.txt
