# Round 2, session 12: K6 with class-uniform warp units: parity in both modes, A/B, ncu capture of the kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gnofix_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2s12_pytest.txt
GNX_GNOFIX_SPLIT=0 timeout 600 python -m pytest tests/test_gnofix_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee -a gpurun_out/r2s12_pytest.txt
timeout 800 python scripts/gnofix_ab.py 10000 3 1,4 0,4 1,3 1,4 2>&1 | tail -8 | tee gpurun_out/r2s12_gnofix_ab.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gnofix_kernel" -c 1 -o gpurun_out/r2s12_gno python scripts/gnofix_ab.py 4096 1 1,4 > gpurun_out/r2s12_ncu_gno.log 2>&1; echo ncu rc=$?
