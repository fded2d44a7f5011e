# Round 2, session 15: K6 after the offset-as-address walk: parity in both modes, A/B.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gnofix_gpu.py tests/test_gnofix_crf_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2s15_pytest.txt
GNX_GNOFIX_SPLIT=0 timeout 600 python -m pytest tests/test_gnofix_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee -a gpurun_out/r2s15_pytest.txt
timeout 800 python scripts/gnofix_ab.py 10000 3 1,4 0,4 1,4 2>&1 | tail -8 | tee gpurun_out/r2s15_gnofix_ab.txt
