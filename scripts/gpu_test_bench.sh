# usage: bash scripts/gpu_test_bench.sh <tag> [bench args...]
tag=$1; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 900 python bench.py "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo bench rc=$?
tail -5 gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$tag.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
for k,v in d['kernels'].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items()})
print(d.get('cpu_baseline')); print(d['clocks'])
PY
