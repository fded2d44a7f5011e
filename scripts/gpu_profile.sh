# usage: bash scripts/gpu_profile.sh <tag> <kernel-regex>   (1 GPU; ncu full capture of the named kernels at N=8192)
tag=$1; pat=$2
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$pat" -c 2 -o gpurun_out/prof_$tag python bench.py --steps 1 --warmup 0 --haps 8192 --e2e-haps 512 --e2e-steps 1 --no-cpu > gpurun_out/ncu_$tag.log 2>&1; echo ncu rc=$?
tail -2 gpurun_out/ncu_$tag.log | cut -c1-300
