"""K6c probe: python scripts/crf_gnofix_probe.py [n_haplotypes] -- bench.py's config_crf_gnofix on its own."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from scripts import bench_configs as bc
from gnomix_b200 import synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
geom, base, smooth, fx, fpop = bc.models_for("chr1")
X = synth.admix_device(torch.from_numpy(fx).cuda(), N, geom[4], seed=3)
print(json.dumps(bench.config_crf_gnofix(X, X.stride(0), base, geom, N)))
