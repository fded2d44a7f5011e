# Round 2 (final kernels): two GPUs -- NCCL scatter/gather parity test, bench line at N=2; smoke(); the new tile seam tests.
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_multigpu_gpu.py "tests/test_gbt_gpu.py::test_gbt_tile_pair_words_at_segment_seams" -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2n2b_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2n2b_smoke.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2n2b_bench.json 2> gpurun_out/r2n2b_bench.err; echo bench rc=$?
tail -5 gpurun_out/r2n2b_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2n2b_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['unpacked']['value'], d['e2e']['packed_input']['value'], d['e2e']['packed_fraction_of_rows'])
print(d.get('strong')); print(d.get('scatter_gather'))
PY
