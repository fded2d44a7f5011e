"""Timing of the other BASELINE.json configurations (not the bench.py headline line):
  cfg2  chr22, 1000-SNP windows, logistic + XGB, 10k haplotypes
  cfg4  chr1, CovRSK string-kernel base (+ XGB smoother), bounded haplotype count
  cfg5  chr1, logistic base + CRF smoother, and logistic + XGB + Gnofix on haplotype pairs
Prints one JSON object per configuration.  python scripts/bench_configs.py [cfg2 cfg4 cfg5]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from gnomix_b200 import synth, _lib
from gnomix_b200.base import CovRSKBase
from gnomix_b200.smooth import CRF_Smoother, CRFModel
from gnomix_b200.gnofix import phase_device


def ev_time(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def models_for(name):
    bench.WORKLOAD = name
    geom = synth.GEOMETRY[name]
    base, smooth, (fx, fpop), (coefs, icpts, ctx), kind = bench.build_models(geom)
    return geom, base, smooth, fx, fpop


def cfg2():
    geom, base, smooth, fx, fpop = models_for("chr22_m1000")
    C, M, A, S, morgans = geom
    W, N = C // M, 10_000
    X = synth.admix_device(torch.from_numpy(fx).cuda(), N, morgans, seed=1)
    ld = X.stride(0)
    B = torch.empty((N, W, A), dtype=torch.float32, device="cuda")
    P = torch.empty_like(B)
    L = torch.empty((N, W), dtype=torch.int32, device="cuda")
    lib, st = _lib.lib(), torch.cuda.current_stream().cuda_stream
    h1, h4 = base.handle(), smooth.model.handle(S)
    t1 = ev_time(lambda: _lib.check(lib.gnx_lr_predict(h1, X.data_ptr(), N, ld, B.data_ptr(), st)), reps=10, warm=3)
    t4 = ev_time(lambda: _lib.check(lib.gnx_gbt_smooth(h4, B.data_ptr(), N, W, P.data_ptr(), L.data_ptr(), st)), reps=10, warm=3)
    return {"config": "cfg2 chr22 C=%d M=%d W=%d, logistic+XGB, N=%d resident" % (C, M, W, N), "K1_ms": t1, "K4_ms": t4,
            "haps_per_s": N / ((t1 + t4) * 1e-3), "K1_GBs": N * (C + W * A * 4) / t1 / 1e6}


def cfg4(N=512, nsv_per_pop=100):
    geom, base, smooth, fx, fpop = models_for("chr1")
    C, M, A, S, morgans = geom
    W = C // M
    rng = np.random.default_rng(7)
    freqs = synth.population_frequencies(rng, C, A)
    tr, trpop = synth.founders(rng, freqs, per_pop=nsv_per_pop)            # training rows = support vectors, grouped by class
    ctx = int(M * 0.5)
    cb = CovRSKBase(chm_len=C, window_size=M, num_ancestry=A, context=ctx)
    P = A * (A - 1) // 2
    trp = cb.pad(tr)
    nsv = len(tr)
    t0 = time.time()
    cb.set_window_svcs([trp[:, lo:hi] for lo, hi in cb.window_slices()], [np.full(A, nsv_per_pop, np.int32)] * W,
                       [rng.normal(0, 1e-4, size=(A - 1, nsv))] * W, [rng.normal(0, 0.1, P)] * W, [np.full(P, -1.0)] * W, [np.zeros(P)] * W)
    X = synth.admix_device(torch.from_numpy(fx).cuda(), N, morgans, seed=2)
    ld = X.stride(0)
    h = cb.handle()
    pack_s = time.time() - t0
    Bd = torch.empty((N, W, A), dtype=torch.float64, device="cuda")
    lib, st = _lib.lib(), torch.cuda.current_stream().cuda_stream
    t = ev_time(lambda: _lib.check(lib.gnx_svc_predict(h, X.data_ptr(), N, ld, Bd.data_ptr(), st)), reps=1, warm=1)
    compares = float(N) * nsv * sum(hi - lo for lo, hi in cb.window_slices())
    return {"config": "cfg4 chr1 CovRSK base, %d support vectors per window, N=%d" % (nsv, N), "K2K3_ms": t, "haps_per_s": N / (t * 1e-3),
            "snp_compares_per_s": compares / (t * 1e-3), "model_pack_s": pack_s, "proba_rowsum_ok": bool(torch.allclose(Bd.sum(-1), torch.ones_like(Bd[..., 0])))}


def cfg5(N=20_000):
    geom, base, smooth, fx, fpop = models_for("chr1")
    C, M, A, S, morgans = geom
    W = C // M
    X = synth.admix_device(torch.from_numpy(fx).cuda(), N, morgans, seed=3)
    ld = X.stride(0)
    # plant phase switch errors: exchange the tails of each pair at ~20 random window boundaries
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    ws = C // W
    for _ in range(20):
        cut = torch.randint(1, W, (N // 2,), device="cuda", generator=g) * ws
        cols = torch.arange(ld, device="cuda")[None, :]
        for i0 in range(0, N // 2, 256):
            sl = slice(2 * i0, 2 * min(i0 + 256, N // 2))
            pair = X[sl].view(-1, 2, ld)
            m = cols >= cut[i0:i0 + pair.shape[0], None]
            a, b = pair[:, 0].clone(), pair[:, 1].clone()
            pair[:, 0] = torch.where(m, b, a)
            pair[:, 1] = torch.where(m, a, b)
    lib, st = _lib.lib(), torch.cuda.current_stream().cuda_stream
    h1 = base.handle()
    out = {"config": "cfg5 chr1, N=%d haplotypes (%d individuals), 20 planted switch errors each" % (N, N // 2)}
    # logistic (float64 out) + CRF
    Bd = torch.empty((N, W, A), dtype=torch.float64, device="cuda")
    crf = CRF_Smoother(n_windows=W, num_ancestry=A, smooth_window_size=S)
    rng = np.random.default_rng(3)
    crf.model = CRFModel(np.eye(A) * 4.0 + rng.normal(0, 0.2, (A, A)), np.eye(A) * 3.0 + rng.normal(0, 0.2, (A, A)))
    Pd = torch.empty((N, W, A), dtype=torch.float64, device="cuda")
    Ld = torch.empty((N, W), dtype=torch.int32, device="cuda")
    t1 = ev_time(lambda: _lib.check(lib.gnx_lr_predict_f64(h1, X.data_ptr(), N, ld, Bd.data_ptr(), st)))
    hc = crf.model.handle()
    t5 = ev_time(lambda: _lib.check(lib.gnx_crf_smooth(hc, Bd.data_ptr(), N, W, Pd.data_ptr(), Ld.data_ptr(), st)))
    out.update(K1_f64_ms=t1, K5_crf_ms=t5, lr_crf_haps_per_s=N / ((t1 + t5) * 1e-3))
    del Bd, Pd
    # logistic + XGB + Gnofix
    B = torch.empty((N, W, A), dtype=torch.float32, device="cuda")
    _lib.check(lib.gnx_lr_predict(h1, X.data_ptr(), N, ld, B.data_ptr(), st))
    Xc, Bc = X.clone(), B.clone()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    Y, trk = phase_device(smooth, Xc, ld, C, Bc, want_tracker=True)
    e1.record()
    torch.cuda.synchronize()
    t6 = e0.elapsed_time(e1)
    sw = (trk[0::2, 1:] != trk[0::2, :-1]).sum().item()
    stats = np.zeros(4, dtype=np.int64)
    lib.gnx_gnofix_last_stats(stats.ctypes.data)
    out.update(gnofix_iterations=int(stats[0]), gnofix_scans=int(stats[1]), gnofix_checks=int(stats[2]), gnofix_accepts=int(stats[3]))
    out.update(K6_gnofix_ms=t6, gnofix_individuals_per_s=(N // 2) / (t6 * 1e-3), accepted_switches_total=int(sw),
               switches_per_individual=sw / (N // 2))
    return out


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg2", "cfg5", "cfg4"]
    for w in which:
        print(json.dumps(globals()[w]()), flush=True)
