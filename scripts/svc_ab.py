"""CovRSK timing probe: python scripts/svc_ab.py [N] -- gnx_svc_predict on 64 chr1-shaped windows (M=857, ctx=428, 700 support
vectors per window); run under GNX_SVC_KERNEL=0 (first kernel) / GNX_SVC_CTAS=3|4 (production kernel builds) for an A/B."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gnomix_b200 import synth, _lib
from gnomix_b200.base import CovRSKBase
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
M, A, nsv_per_pop = 857, 7, 100
C2 = 64 * M + 13
rng = np.random.default_rng(7)
freqs = synth.population_frequencies(rng, C2, A)
tr, _ = synth.founders(rng, freqs, per_pop=nsv_per_pop)
fx2, _ = synth.founders(rng, freqs, per_pop=20)
cb = CovRSKBase(chm_len=C2, window_size=M, num_ancestry=A, context=M // 2)
P, W2, nsv = A * (A - 1) // 2, cb.W, len(tr)
trp = cb.pad(tr)
cb.set_window_svcs([trp[:, lo:hi] for lo, hi in cb.window_slices()], [np.full(A, nsv_per_pop, np.int32)] * W2,
                   [rng.normal(0, 1e-4, size=(A - 1, nsv))] * W2, [rng.normal(0, 0.1, P)] * W2, [np.full(P, -1.0)] * W2, [np.zeros(P)] * W2)
X = synth.admix_device(torch.from_numpy(fx2).cuda(), N, 0.15, seed=2)
ld = X.stride(0); h = cb.handle(); lib = _lib.lib(); st = torch.cuda.current_stream().cuda_stream
Bd = torch.empty((N, W2, A), dtype=torch.float64, device="cuda")
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
tp = t(lambda: _lib.check(lib.gnx_svc_predict(h, X.data_ptr(), N, ld, Bd.data_ptr(), st)))
lens = sum(hi - lo for lo, hi in cb.window_slices())
print("GNX_SVC_KERNEL=%s GNX_SVC_CTAS=%s N=%d: %.3f ms per window, %.3e SNP compares/s, checksum %.12f"
      % (os.environ.get("GNX_SVC_KERNEL", "-"), os.environ.get("GNX_SVC_CTAS", "-"), N, tp / W2, N * nsv * lens / (tp * 1e-3), float(Bd.sum().item())))
