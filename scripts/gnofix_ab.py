"""K6 A/B in one process: python scripts/gnofix_ab.py [n_individuals] [reps] -- chr1, 20 planted switch errors per individual;
every (GNX_GNOFIX_SPLIT, GNX_GNOFIX_TEAMS) setting is run `reps` times on fresh copies of the same pair block (the knobs are
read per call), labels compared across settings."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scripts import bench_configs as bc
from gnomix_b200 import synth, _lib
from gnomix_b200.gnofix import phase_device
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
N = 2 * n
geom, base, smooth, fx, fpop = bc.models_for("chr1")
C, M, A, S, morgans = geom
W = C // M
X = synth.admix_device(torch.from_numpy(fx).cuda(), N, morgans, seed=3)
ld = X.stride(0)
g = torch.Generator(device="cuda"); g.manual_seed(5)
ws = C // W
for _ in range(20):
    cut = torch.randint(1, W, (n,), device="cuda", generator=g) * ws
    cols = torch.arange(ld, device="cuda")[None, :]
    for i0 in range(0, n, 256):
        sl = slice(2 * i0, 2 * min(i0 + 256, n))
        pair = X[sl].view(-1, 2, ld)
        m = cols >= cut[i0:i0 + pair.shape[0], None]
        a, b = pair[:, 0].clone(), pair[:, 1].clone()
        pair[:, 0] = torch.where(m, b, a)
        pair[:, 1] = torch.where(m, a, b)
lib, st = _lib.lib(), torch.cuda.current_stream().cuda_stream
B = torch.empty((N, W, A), dtype=torch.float32, device="cuda")
_lib.check(lib.gnx_lr_predict(base.handle(), X.data_ptr(), N, ld, B.data_ptr(), st))
ref = None
settings = [s.split(",") for s in (sys.argv[3:] or ["1,4", "0,4", "1,3", "0,3", "1,4"])]
for blk, teams in settings:
    os.environ["GNX_GNOFIX_SPLIT"] = blk
    os.environ["GNX_GNOFIX_TEAMS"] = teams
    ts = []
    for r in range(reps):
        Bc = B.clone()
        Xc = X.clone()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        Y, trk = phase_device(smooth, Xc, ld, C, Bc, want_tracker=True)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
        del Xc
    same = True if ref is None else bool(torch.equal(Y, ref))
    if ref is None:
        ref = Y.clone()
    print("SPLIT=%s TEAMS=%s: %s ms (min %.1f)  labels_same=%s" % (blk, teams, " ".join("%.1f" % t for t in ts), min(ts), same), flush=True)
