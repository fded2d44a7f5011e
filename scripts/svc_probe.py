import sys, os, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/scripts")
import numpy as np, torch
import bench, bench_configs as bc
from gnomix_b200 import synth, _lib
from gnomix_b200.base import CovRSKBase
N=8192; nsv_per_pop=100
geom, base, smooth, fx, fpop = bc.models_for("chr1")
C, M, A, S, morgans = geom
W=C//M
# small model: only first 40 windows' worth by shrinking C? keep geometry but build only few windows -> use C2 = 40*M+13
C2 = 40*M + 13
rng=np.random.default_rng(7)
freqs = synth.population_frequencies(rng, C2, A)
tr, trpop = synth.founders(rng, freqs, per_pop=nsv_per_pop)
fx2, _ = synth.founders(rng, freqs, per_pop=20)
cb = CovRSKBase(chm_len=C2, window_size=M, num_ancestry=A, context=M//2)
P=A*(A-1)//2; W2=cb.W; nsv=len(tr)
trp=cb.pad(tr)
cb.set_window_svcs([trp[:, lo:hi] for lo,hi in cb.window_slices()], [np.full(A,nsv_per_pop,np.int32)]*W2, [rng.normal(0,1e-4,size=(A-1,nsv))]*W2, [rng.normal(0,0.1,P)]*W2, [np.full(P,-1.0)]*W2, [np.zeros(P)]*W2)
X = synth.admix_device(torch.from_numpy(fx2).cuda(), N, 0.08, seed=2)
ld=X.stride(0); h=cb.handle(); lib=_lib.lib(); st=torch.cuda.current_stream().cuda_stream
K=torch.empty((N,nsv),dtype=torch.int32,device="cuda")
Bd=torch.empty((N,W2,A),dtype=torch.float64,device="cuda")
def t(fn,reps=3):
    fn(); torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/reps
tk=t(lambda: _lib.check(lib.gnx_svc_kernel_window(h, 5, X.data_ptr(), N, ld, K.data_ptr(), st)))
tp=t(lambda: _lib.check(lib.gnx_svc_predict(h, X.data_ptr(), N, ld, Bd.data_ptr(), st)))
print("K2 one window: %.3f ms; predict per window: %.3f ms (W=%d)" % (tk, tp/W2, W2))
