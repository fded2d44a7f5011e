# Round 2, session 9: K6 A/B -- first image vs block image with shared-memory tops, 3 vs 4 teams.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gnofix_gpu.py tests/test_svc_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2s9_pytest.txt
for b in 1 0; do for t in 4 3; do echo "GNX_GNOFIX_BLK=$b GNX_GNOFIX_TEAMS=$t"; GNX_GNOFIX_BLK=$b GNX_GNOFIX_TEAMS=$t timeout 600 python scripts/gnofix_probe.py 10000 2>&1 | tail -1 | grep -o "'K6_gnofix_ms': [0-9.]*"; done; done | tee gpurun_out/r2s9_gnofix_ab.txt
