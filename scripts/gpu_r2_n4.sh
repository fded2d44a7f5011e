# Round 2 (final kernels): eight GPUs -- bench line at N=8 (weak + strong + scatter_gather + e2e int8 / unpacked / packed input).
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2n4_bench.json 2> gpurun_out/r2n4_bench.err; echo bench rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2n4_bench.json'))
e=d['e2e']
print(d['value'], d['ms_per_step'], e['value'], e['unpacked']['value'], e['packed_input']['value'], e['packed_fraction_of_rows'], e.get('host_threads'))
print(d.get('strong')); print(d.get('scatter_gather'))
PY
