# Round 2, session 11: K6 with 144-byte tree records, (row, class) chain tasks, 4-tree check: parity + A/B.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gnofix_gpu.py tests/test_pipeline_gpu.py tests/test_gbt_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2s11_pytest.txt
GNX_GNOFIX_SPLIT=0 timeout 600 python -m pytest tests/test_gnofix_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee -a gpurun_out/r2s11_pytest.txt
timeout 800 python scripts/gnofix_ab.py 10000 3 2>&1 | tail -8 | tee gpurun_out/r2s11_gnofix_ab.txt
