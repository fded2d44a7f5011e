"""K5 probe: python scripts/crf_probe.py -- gnx_crf_smooth on chr1-shaped float64 base probabilities (W=1430, A=7) at several N."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gnomix_b200 import _lib
from gnomix_b200.smooth import CRFModel
W, A = 1430, 7
rng = np.random.default_rng(3)
crf = CRFModel(np.eye(A) * 4.0 + rng.normal(0, 0.2, (A, A)), np.eye(A) * 3.0 + rng.normal(0, 0.2, (A, A)))
lib, st, h = _lib.lib(), torch.cuda.current_stream().cuda_stream, crf.handle()
for N in [int(a) for a in sys.argv[1:]] or [8192, 20000, 50000]:
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    B = torch.rand((N, W, A), dtype=torch.float64, device="cuda", generator=g)
    B /= B.sum(-1, keepdim=True)
    P = torch.empty_like(B); L = torch.empty((N, W), dtype=torch.int32, device="cuda")
    f = lambda: _lib.check(lib.gnx_crf_smooth(h, B.data_ptr(), N, W, P.data_ptr(), L.data_ptr(), st))
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): f()
    e1.record(); torch.cuda.synchronize()
    print("N=%d: %.3f ms  checksum %.12f %d" % (N, e0.elapsed_time(e1) / 3, float(P.sum().item()), int(L.sum().item())), flush=True)
    del B, P, L
