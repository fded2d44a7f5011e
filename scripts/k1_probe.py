"""K1 timing probe: python scripts/k1_probe.py [N]  (env GNX_LR_DBG selects profiling switches)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from gnomix_b200 import synth, _lib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
geom = synth.GEOMETRY["chr1"]
C, M, A, S, morgans = geom
W = C // M
base, smooth, (fx, fpop), _, _ = bench.build_models(geom)
ld = (C + 127) // 128 * 128
X = torch.randint(0, 2, (N, ld), dtype=torch.int8, device="cuda")
B = torch.empty((N, W, A), dtype=torch.float32, device="cuda")
lib = _lib.lib(); h = base.handle(); st = torch.cuda.current_stream().cuda_stream
for dbg in os.environ.get("DBGS", "0").split(","):
    os.environ["GNX_LR_DBG"] = dbg
    for _ in range(2):
        _lib.check(lib.gnx_lr_predict(h, X.data_ptr(), N, ld, B.data_ptr(), st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        _lib.check(lib.gnx_lr_predict(h, X.data_ptr(), N, ld, B.data_ptr(), st))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("dbg=%s N=%d: %.3f ms  %.0f GB/s (X bytes only)" % (dbg, N, ms, N * C / ms / 1e6), flush=True)
