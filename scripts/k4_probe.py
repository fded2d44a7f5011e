"""K4 timing probe: python scripts/k4_probe.py [N] -- times the rank-form flavours of gnx_gbt_smooth
(gnx_gbt_set_kernel 10 / 14 row kernel walks, 16 tile kernel) on the bench forest and checks they agree bit for bit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from gnomix_b200 import synth, _lib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
kinds = [int(k) for k in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["14", "16"])]
geom = synth.GEOMETRY["chr1"]
C, M, A, S, morgans = geom
W = C // M
base, smooth, (fx, fpop), _, _ = bench.build_models(geom)
ld = (C + 127) // 128 * 128
X = synth.admix_device(torch.from_numpy(fx).cuda(), N, morgans, seed=3, ld=ld)
B = torch.empty((N, W, A), dtype=torch.float32, device="cuda")
lib = _lib.lib(); st = torch.cuda.current_stream().cuda_stream
_lib.check(lib.gnx_lr_predict(base.handle(), X.data_ptr(), N, ld, B.data_ptr(), st))
del X
h = smooth.model.handle(S)
ref = None
for kind in kinds:
    _lib.check(lib.gnx_gbt_set_kernel(h, kind))
    P = torch.empty((N, W, A), dtype=torch.float32, device="cuda"); L = torch.empty((N, W), dtype=torch.int32, device="cuda")
    for _ in range(2):
        _lib.check(lib.gnx_gbt_smooth(h, B.data_ptr(), N, W, P.data_ptr(), L.data_ptr(), st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        _lib.check(lib.gnx_gbt_smooth(h, B.data_ptr(), N, W, P.data_ptr(), L.data_ptr(), st))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    same = "" if ref is None else (" same=%s" % bool(torch.equal(P, ref[0]) and torch.equal(L, ref[1])))
    if ref is None:
        ref = (P.clone(), L.clone())
    print("kernel %d N=%d: %.3f ms  %.0f G tree-rows/s%s" % (kind, N, ms, N * W * smooth.model.n_trees / ms / 1e6, same), flush=True)
