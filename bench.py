#!/usr/bin/env python
"""bench.py -- haplotypes/sec of the local-ancestry inference hot path
(Base.predict_proba -> Smoother.predict_proba + argmax) on B200, next to the CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (K1 logistic base for all windows, K4 tree-ensemble
smoother with fused argmax: a u16 rank pass + the tile walk) over one batch of synthetic admixed
haplotypes that is already resident in HBM as int8.  Workload = BASELINE.json configs[2]: chr1
geometry (C=1,226,139 SNPs, M=857, W=1430 windows, A=7, S=75), logistic base + XGB-semantics
smoother, 50,000 haplotypes PER GPU (weak scaling: every rank owns its own shard, no collective on
the data path).  `value` = haplotypes of all ranks / max-over-ranks step time.

Also on the one JSON line (rank 0):
  e2e        the same metric through the host-buffer C-ABI entry point gnx_infer_host (pinned host
             int8 in, labels out, H2D/D2H inside the timed region); e2e.plugin_pageable is the plugin
             call a gnomix user makes, Gnomix.predict_host(numpy int8 matrix in pageable memory);
             e2e.unpacked the same C call with host-side 2-bit packing off; e2e.packed_input the
             pipeline fed rows that are already 2-bit planes (what the repo's VCF reader hands the
             driver) -- an extra, the headline stays the reference's int8 interface.
  strong     BASELINE configs[2] as worded: 50,000 haplotypes in all, 50,000/N per rank.
  scatter_gather  (N > 1) rank 0's int8 block scattered over NCCL, hot path on every rank, labels
             gathered; compared with rank 0 running the whole block alone.
  parity     GPU labels / proba / float32 B of ALL haplotypes of the CPU sample against the float64
             CPU path (N = 1).
  configs    (N = 1) the other BASELINE configs on the same box: chr22 / M=1000 / 10k haplotypes,
             CovRSK base on chr1, logistic + CRF and logistic + XGB + Gnofix on chr1 pairs -- ms per
             kernel, haplotypes/s, algorithmic bytes, parity against the oracle on a sample.
  roofline   dominant kernel against the roofline that bounds it; ncu-derived fields come from the
             tracked profiles/ncu_facts.json (never literals in this file).
"""
from __future__ import annotations

import argparse
import ctypes as C_
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
_OUT = sys.stdout   # main() re-points it at the real stdout and sends file descriptor 1 to stderr
sys.path.insert(0, ROOT)

WORKLOAD = "chr1"
SEED = 94305  # reference default seed, config.yaml:2
N_SMS = 148


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _ncu_facts():
    """ncu-derived per-kernel facts (DRAM bytes per launch, issue / LSU utilisation) of the default
    workload, extracted from the tracked captures by scripts/ncu_facts.py."""
    p = os.path.join(ROOT, "profiles", "ncu_facts.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_models(geom, workload=None):
    """Synthetic model of the workload's architecture: discriminant LR weights per window +
    a 100-round x A-tree depth-4 forest (trained fixture if present, else seeded random)."""
    from gnomix_b200 import synth
    from gnomix_b200.base import LogisticRegressionBase, Base
    from gnomix_b200.smooth import XGB_Smoother
    from gnomix_b200.gbt import GBTForest
    workload = workload or WORKLOAD
    C, M, A, S, morgans = geom
    rng = np.random.default_rng(SEED)
    freqs = synth.population_frequencies(rng, C, A)
    fx, fpop = synth.founders(rng, freqs, per_pop=20)
    ctx = int(M * 0.5)
    coefs, icpts = synth.discriminant_lr_weights(freqs, C, M, ctx)
    base = LogisticRegressionBase.__new__(LogisticRegressionBase)
    Base.__init__(base, chm_len=C, window_size=M, num_ancestry=A, context=ctx)
    base.base_multithread = True
    base.set_window_weights(coefs, icpts)
    fixture = os.path.join(ROOT, "gnomix_b200", "data", "forest_%s.npz" % workload)
    if os.path.exists(fixture):
        forest = GBTForest.from_npz_dict(np.load(fixture))
        forest_kind = "HGB-trained fixture (xgboost hyper-parameters), gnomix_b200/data/forest_%s.npz" % workload
    else:
        forest = GBTForest.random(np.random.default_rng(SEED + 1), A, S, n_rounds=100, depth=4)
        forest_kind = "seeded random complete forest, 100 rounds x %d trees, depth 4" % A
    smooth = XGB_Smoother(n_windows=C // M, num_ancestry=A, smooth_window_size=S)
    smooth.model = forest
    return base, smooth, (fx, fpop), (coefs, icpts, ctx), forest_kind


def cpu_path(X, coefs, icpts, geom, ctx, forest, want_B=False):
    """The reference's CPU path restated (oracle/): per-window float64 GEMM + expit +
    normalise (what sklearn's predict_proba does under Base.predict_proba_vectorized),
    then slide_window + tree predictor + argmax with OpenMP (what xgboost does)."""
    from oracle import np_oracle as npo, c_oracle as co
    C, M, A, S, _ = geom
    B = npo.lr_base_predict_proba(X, coefs, icpts, C, M, ctx)
    proba, label = co.gbt_smooth(forest, B.astype(np.float32), S)
    return (proba, label, B) if want_B else (proba, label)


def time_cpu(X_sample, coefs, icpts, geom, ctx, forest, budget_s=15.0):
    """(haplotypes/s, n, seconds, (proba, label, B64) of those n haplotypes)."""
    n0 = min(64, len(X_sample))
    t = time.perf_counter()
    cpu_path(X_sample[:n0], coefs, icpts, geom, ctx, forest)
    dt = time.perf_counter() - t
    n = int(max(n0, min(len(X_sample), n0 * budget_s / max(dt, 1e-3))))
    t = time.perf_counter()
    res = cpu_path(X_sample[:n], coefs, icpts, geom, ctx, forest, want_B=True)
    dt = time.perf_counter() - t
    return n / dt, n, dt, res


def run_reference(args):
    """--impl reference: the CPU path on the host cores (oracle port; the reference's own
    stack -- sklearn 1.0.1 liblinear, xgboost 1.1.1 -- is not installable offline, and
    /root/reference does not exist on the GPU box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from gnomix_b200 import synth
    from oracle import c_oracle as co
    co.use_all_cores()
    geom = synth.GEOMETRY[WORKLOAD]
    C, M, A, S, morgans = geom
    _, smooth, (fx, fpop), (coefs, icpts, ctx), forest_kind = build_models(geom)
    rng = np.random.default_rng(SEED + 7)
    n = args.cpu_haps
    X, _ = synth.admix_host(rng, fx, fpop, n, morgans)
    for _ in range(args.warmup):
        cpu_path(X[:8], coefs, icpts, geom, ctx, smooth.model)
    t = time.perf_counter()
    for _ in range(args.steps):
        cpu_path(X, coefs, icpts, geom, ctx, smooth.model)
    dt = (time.perf_counter() - t) / args.steps
    v = n / dt
    cores = co.num_threads()
    sample = "%d of the workload's haplotypes per step (chr1 geometry, all %d windows)" % (n, C // M)
    print(json.dumps({
        "impl": "reference", "metric": "haplotypes/sec local-ancestry inference (Base->Smooth->argmax), chr1, 7-way",
        "value": v, "unit": "haplotypes/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "chr1 C=%d M=%d W=%d A=%d S=%d logistic+XGB; CPU sample of %d haplotypes" % (C, M, C // M, A, S, n),
                   "forest": forest_kind},
        "cpu_baseline": {"value": v, "unit": "haplotypes/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "haplotypes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), file=_OUT, flush=True)


# ------------------------------------------------------------------------------------------------
def _ev_time(fn, reps=3, warm=1):
    import torch
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


def _bits_equal(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    v = {4: np.uint32, 8: np.uint64}[a.dtype.itemsize]
    return bool(a.shape == b.shape and np.array_equal(a.view(v), b.view(v)))


def config_chr22(peak):
    """BASELINE configs[1]: chr22, 1000-SNP windows, logistic + XGB, 10k haplotypes, 1 GPU."""
    import torch
    from gnomix_b200 import synth, _lib
    from tests import util
    from oracle import c_oracle as co
    geom = synth.GEOMETRY["chr22_m1000"]
    C, M, A, S, morgans = geom
    W, N = C // M, 10_000
    base, smooth, (fx, fpop), (coefs, icpts, ctx), kind = build_models(geom, "chr22_m1000")
    X = synth.admix_device(torch.from_numpy(fx).cuda(), N, morgans, seed=1)
    ld = X.stride(0)
    B = torch.empty((N, W, A), dtype=torch.float32, device="cuda")
    P = torch.empty_like(B)
    L = torch.empty((N, W), dtype=torch.int32, device="cuda")
    lib, st = _lib.lib(), torch.cuda.current_stream().cuda_stream
    h1, h4 = base.handle(), smooth.model.handle(S)
    t1 = _ev_time(lambda: _lib.check(lib.gnx_lr_predict(h1, X.data_ptr(), N, ld, B.data_ptr(), st)), reps=10, warm=3)
    t4 = _ev_time(lambda: _lib.check(lib.gnx_gbt_smooth(h4, B.data_ptr(), N, W, P.data_ptr(), L.data_ptr(), st)), reps=10, warm=3)
    ns = 64
    Xs = X[:ns, :C].cpu().numpy()
    B_o, _ = util.oracle_lr_fixed(Xs, coefs, icpts, C, M, ctx, A)
    p_o, l_o = co.gbt_smooth(smooth.model, B_o, S)
    k1b, k4b = N * (C + W * A * 4), N * (2 * W * A * 4 + W * 4)
    return {"workload": "chr22 C=%d M=%d W=%d A=%d S=%d, logistic + XGB, %d haplotypes resident" % (C, M, W, A, S, N),
            "forest": kind, "K1_ms": t1, "K4_ms": t4, "haplotypes_per_s": N / ((t1 + t4) * 1e-3),
            "K1_algorithmic_bytes": k1b, "K1_frac_hbm": k1b / t1 / 1e6 / peak, "K4_algorithmic_bytes": k4b,
            "K4_tree_traversals_per_s": N * W * smooth.model.n_trees / (t4 * 1e-3),
            "parity_vs_oracle": {"haplotypes": ns, "B_f32_bit_exact": _bits_equal(B[:ns].cpu().numpy(), B_o),
                                 "proba_bit_exact": _bits_equal(P[:ns].cpu().numpy(), p_o),
                                 "labels_equal": bool(np.array_equal(L[:ns].cpu().numpy(), l_o))}}


def config_covrsk(X, ld, geom, n_haps, peak, nsv_per_pop=100):
    """BASELINE configs[3]: chr1, CovRSK string-kernel base (K2 + K3), on the first n_haps rows of the
    resident chr1 matrix; 700 support vectors per window (100 per population).  XGB smoother after it
    is K4 as in the headline step."""
    import torch
    from gnomix_b200 import synth, _lib
    from gnomix_b200.base import CovRSKBase
    from oracle import c_oracle as co, np_oracle as npo
    C, M, A, S, morgans = geom
    W = C // M
    rng = np.random.default_rng(7)
    freqs = synth.population_frequencies(rng, C, A)
    tr, _ = synth.founders(rng, freqs, per_pop=nsv_per_pop)       # training rows = support vectors, grouped by class
    ctx = int(M * 0.5)
    cb = CovRSKBase(chm_len=C, window_size=M, num_ancestry=A, context=ctx)
    P = A * (A - 1) // 2
    trp = cb.pad(tr)
    nsv = len(tr)
    sl = cb.window_slices()
    dual = rng.normal(0, 1e-4, size=(A - 1, nsv))
    icpt, pA, pB = rng.normal(0, 0.1, P), np.full(P, -1.0), np.zeros(P)
    nsup = np.full(A, nsv_per_pop, np.int32)
    t0 = time.perf_counter()
    cb.set_window_svcs([trp[:, lo:hi] for lo, hi in sl], [nsup] * W, [dual] * W, [icpt] * W, [pA] * W, [pB] * W)
    h = cb.handle()
    pack_s = time.perf_counter() - t0
    N = n_haps
    Bd = torch.empty((N, W, A), dtype=torch.float64, device="cuda")
    lib, st = _lib.lib(), torch.cuda.current_stream().cuda_stream
    t = _ev_time(lambda: _lib.check(lib.gnx_svc_predict(h, X.data_ptr(), N, ld, Bd.data_ptr(), st)), reps=1, warm=1)
    compares = float(N) * nsv * sum(hi - lo for lo, hi in sl)
    # parity: integer kernel values and float64 probabilities of 4 haplotypes x 3 windows against the oracle
    ns = 4
    Xs = X[:ns, :C].cpu().numpy()
    Xsp = npo.base_pad(Xs, ctx)
    k_ok, p_ok = True, True
    Bh = Bd[:ns].cpu().numpy()
    for w in (0, W // 2, W - 1):
        lo, hi = sl[w]
        Ko = co.covrsk(Xsp[:, lo:hi], np.ascontiguousarray(trp[:, lo:hi]), npo.cov_sample(hi - lo))
        k_ok &= bool(np.array_equal(cb.kernel_window(w, X[:ns, :C]).cpu().numpy(), Ko))
        p_ok &= _bits_equal(Bh[:, w, :], co.svc_proba(Ko, nsup, dual, icpt, pA, pB))
    return {"workload": "chr1 C=%d W=%d A=%d, CovRSK string-kernel base, %d support vectors per window, %d haplotypes resident"
                        % (C, W, A, nsv, N),
            "K2K3_ms": t, "haplotypes_per_s": N / (t * 1e-3), "snp_compares_per_s": compares / (t * 1e-3),
            "algorithmic_bytes": N * (C + W * A * 8), "model_pack_s": pack_s, "bound": "INT ALU (popcount / funnel-shift pipe)",
            "parity_vs_oracle": {"haplotypes": ns, "windows": 3, "kernel_values_equal": k_ok, "proba_f64_bit_exact": p_ok}}


def config_crf(X, ld, base, geom, n_haps):
    """BASELINE configs[4], first half: chr1, logistic base (float64 out) + CRF smoother."""
    import torch
    from gnomix_b200 import _lib
    from gnomix_b200.smooth import CRF_Smoother, CRFModel
    from oracle import c_oracle as co
    C, M, A, S, morgans = geom
    W, N = C // M, n_haps
    lib, st = _lib.lib(), torch.cuda.current_stream().cuda_stream
    Bd = torch.empty((N, W, A), dtype=torch.float64, device="cuda")
    Pd = torch.empty((N, W, A), dtype=torch.float64, device="cuda")
    Ld = torch.empty((N, W), dtype=torch.int32, device="cuda")
    crf = CRF_Smoother(n_windows=W, num_ancestry=A, smooth_window_size=S)
    rng = np.random.default_rng(3)
    crf.model = CRFModel(np.eye(A) * 4.0 + rng.normal(0, 0.2, (A, A)), np.eye(A) * 3.0 + rng.normal(0, 0.2, (A, A)))
    h1, hc = base.handle(), crf.model.handle()
    t1 = _ev_time(lambda: _lib.check(lib.gnx_lr_predict_f64(h1, X.data_ptr(), N, ld, Bd.data_ptr(), st)))
    t5 = _ev_time(lambda: _lib.check(lib.gnx_crf_smooth(hc, Bd.data_ptr(), N, W, Pd.data_ptr(), Ld.data_ptr(), st)))
    ns = 64
    p_o, l_o = co.crf_smooth(Bd[:ns].cpu().numpy(), crf.model.state_w, crf.model.trans_w)
    return {"workload": "chr1 W=%d A=%d, logistic base (float64) + CRF smoother, %d haplotypes resident" % (W, A, N),
            "K1_f64_ms": t1, "K5_ms": t5, "haplotypes_per_s": N / ((t1 + t5) * 1e-3),
            "K5_algorithmic_bytes": N * (2 * W * A * 8 + W * 4),
            "parity_vs_oracle": {"haplotypes": ns, "marginals_f64_bit_exact": _bits_equal(Pd[:ns].cpu().numpy(), p_o),
                                 "labels_equal": bool(np.array_equal(Ld[:ns].cpu().numpy(), l_o))}}


def plant_switches(X, ld, C, W, n_switch=20, seed=5):
    """Exchange the tails of each haplotype pair at n_switch random window boundaries (phase errors)."""
    import torch
    N = X.shape[0]
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    ws = C // W
    cols = torch.arange(ld, device="cuda")[None, :]
    for _ in range(n_switch):
        cut = torch.randint(1, W, (N // 2,), device="cuda", generator=g) * ws
        for i0 in range(0, N // 2, 256):
            sl = slice(2 * i0, 2 * min(i0 + 256, N // 2))
            pair = X[sl].view(-1, 2, ld)
            m = cols >= cut[i0:i0 + pair.shape[0], None]
            a, b = pair[:, 0].clone(), pair[:, 1].clone()
            pair[:, 0] = torch.where(m, b, a)
            pair[:, 1] = torch.where(m, a, b)


def config_gnofix(X, ld, base, smooth, geom, n_haps):
    """BASELINE configs[4], second half: chr1, logistic + XGB + Gnofix on haplotype pairs with 20 planted
    phase-switch errors each (the reference supports Gnofix with the XGB smoother only, src/model.py:194)."""
    import torch
    from gnomix_b200 import _lib
    from gnomix_b200.gnofix import phase_device
    from oracle import c_oracle as co, np_oracle as npo
    C, M, A, S, morgans = geom
    W, N = C // M, n_haps
    lib, st = _lib.lib(), torch.cuda.current_stream().cuda_stream
    Xc = X[:N].clone()
    plant_switches(Xc, ld, C, W)
    B = torch.empty((N, W, A), dtype=torch.float32, device="cuda")
    _lib.check(lib.gnx_lr_predict(base.handle(), Xc.data_ptr(), N, ld, B.data_ptr(), st))
    ns = 2   # individuals checked against the oracle restatement of the reference's gnofix
    X_in, B_in = Xc[:2 * ns, :C].cpu().numpy(), B[:2 * ns].cpu().numpy()
    nw_ = min(N, 512)   # warm-up on copies of a few pairs (module load, pool growth), then ONE timed pass over all
    phase_device(smooth, Xc[:nw_].clone(), ld, C, B[:nw_].clone())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    Y, trk = phase_device(smooth, Xc, ld, C, B, want_tracker=True)
    e1.record()
    torch.cuda.synchronize()
    t6 = e0.elapsed_time(e1)
    stats = np.zeros(4, dtype=np.int64)
    lib.gnx_gnofix_last_stats(stats.ctypes.data)
    sw = int((trk[0::2, 1:] != trk[0::2, :-1]).sum().item())
    forest = smooth.model
    rows_fn = lambda rows: co.gbt_rows(forest, rows)
    smooth_fn = lambda b: co.gbt_smooth(forest, b, S, want_proba=False)[1]
    ok_y = ok_x = ok_t = True
    Yh, Th, Xh = Y[:2 * ns].cpu().numpy(), trk[:2 * ns].cpu().numpy(), Xc[:2 * ns, :C].cpu().numpy()
    for i in range(ns):
        X_m, X_p, Y_m, Y_p, t = npo.gnofix_default(X_in[2 * i], X_in[2 * i + 1], B_in[2 * i:2 * i + 2], S, rows_fn, smooth_fn)
        ok_y &= bool(np.array_equal(Yh[2 * i:2 * i + 2], np.array([Y_m, Y_p])))
        ok_t &= bool(np.array_equal(Th[2 * i:2 * i + 2], t))
        ok_x &= bool(np.array_equal(Xh[2 * i:2 * i + 2], np.array([X_m, X_p])))
    return {"workload": "chr1 W=%d A=%d, logistic + XGB + Gnofix, %d individuals resident, 20 planted switch errors each" % (W, A, N // 2),
            "K6_ms": t6, "individuals_per_s": (N // 2) / (t6 * 1e-3), "iterations": int(stats[0]), "checks": int(stats[2]),
            "accepted_switches": int(stats[3]), "net_switches_per_individual": sw / (N // 2),
            "algorithmic_bytes": N * (2 * C + 2 * W * A * 4 + 2 * W * 4),
            "parity_vs_oracle": {"individuals": ns, "labels_equal": ok_y, "tracker_equal": ok_t, "X_phased_equal": ok_x}}


def config_crf_gnofix(X, ld, base, geom, n_haps):
    """BASELINE configs[4] as worded ("CRF smoother + Gnofix"): AN EXTENSION, the reference refuses the combination
    (src/model.py:194) -- gnx_gnofix_crf, defined in include/gnx.h; the checker is the oracle's restatement of the
    reference's gnofix control flow with the oracle's CRF plugged in (no reference oracle exists)."""
    import torch
    from gnomix_b200 import _lib
    from gnomix_b200.gnofix import phase_device_crf
    from gnomix_b200.smooth import CRF_Smoother, CRFModel
    from oracle import c_oracle as co, np_oracle as npo
    C, M, A, S, morgans = geom
    W, N = C // M, n_haps
    lib, st = _lib.lib(), torch.cuda.current_stream().cuda_stream
    Xc = X[:N].clone()
    plant_switches(Xc, ld, C, W)
    crf = CRF_Smoother(n_windows=W, num_ancestry=A, smooth_window_size=S)
    rng = np.random.default_rng(3)
    crf.model = CRFModel(np.eye(A) * 4.0 + rng.normal(0, 0.2, (A, A)), np.eye(A) * 3.0 + rng.normal(0, 0.2, (A, A)))
    B = torch.empty((N, W, A), dtype=torch.float64, device="cuda")
    _lib.check(lib.gnx_lr_predict_f64(base.handle(), Xc.data_ptr(), N, ld, B.data_ptr(), st))
    ns = 2
    X_in, B_in = Xc[:2 * ns, :C].cpu().numpy(), B[:2 * ns].cpu().numpy()
    nw_ = min(N, 128)
    phase_device_crf(crf, Xc[:nw_].clone(), ld, C, B[:nw_].clone())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    Y, trk = phase_device_crf(crf, Xc, ld, C, B, want_tracker=True)
    e1.record()
    torch.cuda.synchronize()
    t6 = e0.elapsed_time(e1)
    stats = np.zeros(4, dtype=np.int64)
    lib.gnx_gnofix_crf_last_stats(stats.ctypes.data)
    sw = int((trk[0::2, 1:] != trk[0::2, :-1]).sum().item())
    ok_y = ok_x = ok_t = True
    Yh, Th, Xh = Y[:2 * ns].cpu().numpy(), trk[:2 * ns].cpu().numpy(), Xc[:2 * ns, :C].cpu().numpy()
    for i in range(ns):
        X_m, X_p, Y_m, Y_p, t = npo.gnofix_crf_extension(X_in[2 * i], X_in[2 * i + 1], B_in[2 * i:2 * i + 2], S, crf.model.state_w,
                                                         crf.model.trans_w, crf_smooth_fn=co.crf_smooth)
        ok_y &= bool(np.array_equal(Yh[2 * i:2 * i + 2], np.array([Y_m, Y_p])))
        ok_t &= bool(np.array_equal(Th[2 * i:2 * i + 2], t))
        ok_x &= bool(np.array_equal(Xh[2 * i:2 * i + 2], np.array([X_m, X_p])))
    return {"workload": "chr1 W=%d A=%d, logistic (float64) + CRF + Gnofix, %d individuals resident, 20 planted switch errors each" % (W, A, N // 2),
            "extension": "no reference behaviour (src/model.py:194 refuses a CRF smoother): the reference's gnofix control flow with "
                         "smoother.predict := argmax CRF marginals, smoother.model.predict_proba(scope) := CRF marginal at the scope's centre",
            "K6c_ms": t6, "individuals_per_s": (N // 2) / (t6 * 1e-3), "rounds": int(stats[0]), "checks": int(stats[1]),
            "accepted_switches": int(stats[2]), "iterations": int(stats[3]), "net_switches_per_individual": sw / (N // 2),
            "parity_vs_oracle": {"individuals": ns, "labels_equal": ok_y, "tracker_equal": ok_t, "X_phased_equal": ok_x,
                                 "oracle": "oracle/np_oracle.py::gnofix_crf_extension (restated control flow + the oracle's CRF; not a reference output)"}}


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from gnomix_b200 import synth, _lib, parallel
    from gnomix_b200.model import Gnomix

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200; no CUDA device is visible (gnomix_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _lib.require_gpu()
    lib = _lib.lib()

    geom = synth.GEOMETRY[WORKLOAD]
    C, M, A, S, morgans = geom
    W = C // M
    N = args.haps
    base, smooth, (fx, fpop), (coefs, icpts, ctx), forest_kind = build_models(geom)
    T = smooth.model.n_trees

    # ---- resident inputs: this rank's shard, generated on the device -------------
    fdev = torch.from_numpy(fx).cuda()
    ld = (C + 127) // 128 * 128
    X = synth.admix_device(fdev, N, morgans, seed=SEED + 1000 * rank, ld=ld)
    B = torch.empty((N, W, A), dtype=torch.float32, device="cuda")
    P = torch.empty((N, W, A), dtype=torch.float32, device="cuda")
    L = torch.empty((N, W), dtype=torch.int32, device="cuda")
    hlr, hgbt = base.handle(), smooth.model.handle(S)
    st = torch.cuda.current_stream().cuda_stream

    def k1(n):
        _lib.check(lib.gnx_lr_predict(hlr, X.data_ptr(), n, ld, B.data_ptr(), st), "gnx_lr_predict")

    def k4(n):
        _lib.check(lib.gnx_gbt_smooth(hgbt, B.data_ptr(), n, W, P.data_ptr(), L.data_ptr(), st), "gnx_gbt_smooth")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(n, steps, warmup):
        """(ms per step max over ranks, K1 ms, K4 ms) of `steps` passes over the first n resident rows."""
        for _ in range(warmup):
            k1(n)
            k4(n)
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * steps + 1)]
        barrier()
        ev[0].record()
        for i in range(steps):
            k1(n)
            ev[2 * i + 1].record()
            k4(n)
            ev[2 * i + 2].record()
        barrier()
        total = ev[0].elapsed_time(ev[-1])
        a = float(np.mean([ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(steps)]))
        b = float(np.mean([ev[2 * i + 1].elapsed_time(ev[2 * i + 2]) for i in range(steps)]))
        return maxr(total) / steps, a, b

    sampler = ClockSampler(local)
    for _ in range(args.warmup):
        k1(N)
        k4(N)
    barrier()
    if rank == 0:
        sampler.start()
    ms_per_step, k1_ms, k4_ms = timed(N, args.steps, 0)
    clocks = sampler.stop() if rank == 0 else None
    value = N * world / (ms_per_step * 1e-3)

    # ---- strong scaling: BASELINE configs[2] as worded, 50,000 haplotypes in all ------------------
    n_strong = args.strong_haps // world
    if world == 1 and n_strong == N:
        strong = {"ms_per_step": ms_per_step, "K1_ms": k1_ms, "K4_ms": k4_ms}
    else:
        s_ms, s_k1, s_k4 = timed(min(n_strong, N), max(3, args.steps), 2)
        strong = {"ms_per_step": s_ms, "K1_ms": s_k1, "K4_ms": s_k4}
    strong.update({"total_haplotypes": min(n_strong, N) * world, "haplotypes_per_gpu": min(n_strong, N),
                   "value": min(n_strong, N) * world / (strong["ms_per_step"] * 1e-3), "unit": "haplotypes/s",
                   "limits": "per-GPU work shrinks to %d haplotypes: K1 runs %d CTAs of 256 haplotypes on 148 SMs and K4 %d tiles "
                             "of 32 haplotypes x <= 64 windows, so the tail wave and the fixed per-launch costs (forest staging, "
                             "schedule) weigh more; no collective is involved" % (min(n_strong, N), -(-min(n_strong, N) // 256),
                                                                                    -(-min(n_strong, N) // 32) * -(-W // 64))})
    # K4's two launches timed separately (CUDA events recorded by the library on the launching stream)
    _lib.check(lib.gnx_gbt_set_profile(hgbt, 1), "gnx_gbt_set_profile")
    k1(N)   # leaves B / P / L of the whole shard behind for the checks below
    ph = []
    for _ in range(3):
        k4(N)
        a_ms, b_ms = C_.c_float(0), C_.c_float(0)
        _lib.check(lib.gnx_gbt_last_phase_ms(hgbt, C_.byref(a_ms), C_.byref(b_ms)), "gnx_gbt_last_phase_ms")
        ph.append((a_ms.value, b_ms.value))
    _lib.check(lib.gnx_gbt_set_profile(hgbt, 0), "gnx_gbt_set_profile")
    k4a_ms, k4b_ms = float(np.mean([p[0] for p in ph])), float(np.mean([p[1] for p in ph]))
    torch.cuda.synchronize()

    # ---- scatter / gather over NCCL around the hot path (outside any timed region) ---------------
    sg = None
    if world > 1:
        n_sg = 2048 * world
        Xroot = X[:n_sg] if rank == 0 else None
        mine = parallel.scatter_rows(Xroot, n_sg, ld, torch.int8, torch.device("cuda", local))
        nl = mine.shape[0]
        Bl = torch.empty((nl, W, A), dtype=torch.float32, device="cuda")
        Ll = torch.empty((nl, W), dtype=torch.int32, device="cuda")
        _lib.check(lib.gnx_lr_predict(hlr, mine.data_ptr(), nl, ld, Bl.data_ptr(), st), "gnx_lr_predict")
        _lib.check(lib.gnx_gbt_smooth(hgbt, Bl.data_ptr(), nl, W, None, Ll.data_ptr(), st), "gnx_gbt_smooth")
        got = parallel.gather_rows(Ll, n_sg)
        if rank == 0:
            sg = {"haplotypes": n_sg, "ranks": world, "scatter_bytes": int(n_sg * ld), "gather_bytes": int(n_sg * W * 4),
                  "gather_labels_match": bool(torch.equal(got, L[:n_sg])),
                  "api": "gnomix_b200.parallel.scatter_rows / gather_rows (torch.distributed NCCL send/recv)"}
        del mine, Bl, Ll, got
        barrier()

    e2e = None
    if not args.no_e2e:
        # ---- e2e: host buffers through gnx_infer_host --------------------------------
        # 8 ranks of one box share the host: keep the pinned e2e batch at 10 GB per rank there
        n_e2e = min(N, args.e2e_haps if world == 1 else min(args.e2e_haps, 8192))
        Xh = torch.empty((n_e2e, ld), dtype=torch.int8, pin_memory=True)
        Xh.copy_(X[:n_e2e])
        Lh = torch.empty((n_e2e, W), dtype=torch.int32, pin_memory=True)
        torch.cuda.synchronize()

        def e2e_measure(fn, check, warm=1):
            for _ in range(warm):
                fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                fn()
            torch.cuda.synchronize()
            e2e_s = maxr((time.perf_counter() - t0) / args.e2e_steps)
            frac, h2d, d2h = C_.c_double(0), C_.c_int64(0), C_.c_int64(0)
            lib.gnx_infer_host_last_transfer(C_.byref(frac), C_.byref(h2d), C_.byref(d2h))
            return n_e2e * world / e2e_s, frac.value, h2d.value, d2h.value, check()

        def c_abi():
            _lib.check(lib.gnx_infer_host(hlr, hgbt, Xh.data_ptr(), n_e2e, ld, None, Lh.data_ptr(), 0), "gnx_infer_host")

        same_as_resident = lambda: bool(torch.equal(Lh.cuda(), L[:n_e2e]))
        # default path: part of every chunk crosses PCIe as 2-bit planes packed by the host cores
        e2e_value, e2e_frac, e2e_h2d, e2e_d2h, labels_match = e2e_measure(c_abi, same_as_resident, warm=3)
        pk, h2dr = C_.c_double(0), C_.c_double(0)
        lib.gnx_infer_host_rates(C_.byref(pk), C_.byref(h2dr))
        # for comparison: the same call with packing switched off (raw int8 over PCIe)
        os.environ["GNX_HOST_PACK"] = "0"
        raw_value, _, raw_h2d, _, raw_match = e2e_measure(c_abi, same_as_resident)
        del os.environ["GNX_HOST_PACK"]
        labels_match = labels_match and raw_match
        # the plugin call a gnomix user makes: Gnomix.predict_host on a numpy matrix in pageable memory
        gm = Gnomix.__new__(Gnomix)
        gm.C, gm.M, gm.A, gm.S, gm.W = C, M, A, S, W
        gm.base, gm.smooth, gm.calibrate = base, smooth, False
        Xnp = np.empty((n_e2e, C), dtype=np.int8)
        Xnp[:] = Xh.numpy()[:, :C]
        box = {}

        def plugin():
            box["labels"] = gm.predict_host(Xnp)

        plug_value, _, plug_h2d, plug_d2h, plug_match = e2e_measure(
            plugin, lambda: bool(np.array_equal(box["labels"], L[:n_e2e].cpu().numpy())))
        del Xnp, box
        # host rows that are ALREADY 2-bit planes (what gnomix_b200.io.vcf_to_packed hands the driver): a quarter of the
        # bytes cross PCIe and no host core packs anything inside the timed region
        from gnomix_b200.io import PackedHaplotypes
        Pk = PackedHaplotypes.from_numpy(Xh.numpy()[:, :C])
        pipe = _lib.Pipeline()
        pipe.lr, pipe.gbt, pipe.x_packed = hlr, hgbt, 1

        def c_abi_packed():
            _lib.check(lib.gnx_infer_host_ex(C_.byref(pipe), Pk.words.ctypes.data, n_e2e, Pk.pitch_words, None, Lh.data_ptr(), None, 0),
                       "gnx_infer_host_ex")

        Lh.zero_()
        pk_value, _, pk_h2d, pk_d2h, pk_match = e2e_measure(c_abi_packed, same_as_resident)
        del Pk
        e2e = {"value": e2e_value, "unit": "haplotypes/s", "h2d_bytes_per_step": int(e2e_h2d),
               "d2h_bytes_per_step": int(e2e_d2h), "haplotypes_per_step": n_e2e,
               "api": "gnx_infer_host (pinned host int8 in, int32 labels out)", "labels_match_resident_path": labels_match,
               "packed_fraction_of_rows": e2e_frac, "host_threads": int(lib.gnx_host_threads()),
               "calibrated_host_pack_gbs": pk.value, "calibrated_h2d_gbs": h2dr.value,
               "unpacked": {"value": raw_value, "h2d_bytes_per_step": int(raw_h2d)},
               "packed_input": {"value": pk_value, "api": "gnx_infer_host_ex(x_packed: pinned 2-bit-plane rows in, int32 labels out)",
                                "h2d_bytes_per_step": int(pk_h2d), "d2h_bytes_per_step": int(pk_d2h),
                                "labels_match_resident_path": pk_match,
                                "note": "not the reference's int8 interface: the form the repo's own VCF reader (vcf_to_packed) "
                                        "produces for the driver, reported beside the int8 number, not instead of it"},
               "plugin_pageable": {"value": plug_value, "api": "Gnomix.predict_host(numpy int8 [N, C], pageable) -> numpy labels",
                                   "h2d_bytes_per_step": int(plug_h2d), "d2h_bytes_per_step": int(plug_d2h),
                                   "labels_match_resident_path": plug_match}}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline + parity of every haplotype of the bounded sample (rank 0, N=1 only) --
    cpu = parity = None
    if world == 1 and not args.no_cpu:
        from oracle import c_oracle as co
        co.use_all_cores()
        ns = min(N, args.cpu_haps)
        Xs = X[:ns, :C].cpu().numpy()
        v, n_used, dt, (p_cpu, l_cpu, B64) = time_cpu(Xs, coefs, icpts, geom, ctx, smooth.model)
        l_gpu, p_gpu, b_gpu = L[:n_used].cpu().numpy(), P[:n_used].cpu().numpy(), B[:n_used].cpu().numpy()
        cpu = {"value": v, "unit": "haplotypes/s", "cores": co.num_threads(), "kind": "port",
               "sample": "%d haplotypes of the same workload in %.1f s (numpy float64 per-window GEMM + OpenMP tree predictor)" % (n_used, dt)}
        parity = {"haplotypes": int(n_used), "rows": int(n_used) * W,
                  "against": "float64 CPU path (sklearn-style per-window GEMM + expit + normalise, float32 hand-off, tree predictor)",
                  "label_mismatches": int((l_cpu != l_gpu).sum()),
                  "max_abs_proba_diff": float(np.max(np.abs(p_cpu - p_gpu))),
                  "proba_bit_mismatches": int((p_cpu.view(np.uint32) != p_gpu.view(np.uint32)).sum()),
                  "B_f32_values": int(b_gpu.size),
                  "B_f32_differ_from_f64_path_rounded": int((B64.astype(np.float32).view(np.uint32) != b_gpu.view(np.uint32)).sum()),
                  "max_abs_B_diff_vs_f64": float(np.max(np.abs(B64 - b_gpu.astype(np.float64))))}

    peak, peak_src = _peaks()
    facts = _ncu_facts()
    k1_bytes = N * (C + W * A * 4)
    k4_bytes = N * (2 * W * A * 4 + W * 4)
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    # K4's own roofline: shared-memory wavefronts.  A depth-4 tree costs a warp 6.5 wavefronts per window (3 feature
    # loads + the root's, which the two adjacent windows of a warp share through a pair word, 2 node loads, 1 leaf
    # load); an SM retires one wavefront per clock.
    k4_wavefronts = N * W * T / 32.0 * 6.5
    lsu_peak = N_SMS * sm_mhz * 1e6
    Wp = W + S - 1
    k4a_bytes = N * (W * A * 4 + Wp * A * 2)            # float32 B in, u16 rank tiles (reflect pad materialised) out
    k4b_bytes = N * (Wp * A * 2 + W * A * 4 + W * 4)    # rank tiles in, float32 proba + int32 labels out
    kernels = {
        "K1_lr_tc_kernel": {"ms": k1_ms, "algorithmic_bytes": k1_bytes, "gbs": k1_bytes / k1_ms / 1e6,
                            "frac_hbm": k1_bytes / k1_ms / 1e6 / peak},
        "K4_gbt_smooth": {"ms": k4_ms, "launches": "gbt_rank_tile_kernel (K4a) + gbt_smooth_tile_kernel (K4b)",
                          "algorithmic_bytes": k4_bytes, "gbs": k4_bytes / k4_ms / 1e6,
                          "frac_hbm": k4_bytes / k4_ms / 1e6 / peak,
                          "tree_traversals_per_s": N * W * T / (k4_ms * 1e-3)},
        "K4a_gbt_rank_tile": {"ms": k4a_ms, "algorithmic_bytes": k4a_bytes, "gbs": k4a_bytes / k4a_ms / 1e6,
                              "frac_hbm": k4a_bytes / k4a_ms / 1e6 / peak},
        "K4b_gbt_smooth_tile": {"ms": k4b_ms, "algorithmic_bytes": k4b_bytes, "gbs": k4b_bytes / k4b_ms / 1e6,
                                "frac_hbm": k4b_bytes / k4b_ms / 1e6 / peak,
                                "tree_traversals_per_s": N * W * T / (k4b_ms * 1e-3),
                                "algorithmic_lsu_wavefronts": k4_wavefronts,
                                "frac_lsu": k4_wavefronts / (k4b_ms * 1e-3) / lsu_peak},
    }
    dom = "K1_lr_tc_kernel" if k1_ms >= k4b_ms else "K4b_gbt_smooth_tile"
    at_default = (N == 50_000 and WORKLOAD == "chr1")
    f4, f1 = facts.get("K4b_gbt_smooth_tile", {}) if at_default else {}, facts.get("K1_lr_tc_kernel", {}) if at_default else {}
    roof_k4 = {"kernel": "K4b_gbt_smooth_tile", "bound": "lsu", "achieved": k4_wavefronts / (k4b_ms * 1e-3) / 1e9,
               "peak": lsu_peak / 1e9, "unit": "G shared-memory wavefronts/s", "frac": kernels["K4b_gbt_smooth_tile"]["frac_lsu"],
               "peak_source": "148 SMs x 1 wavefront per clock x the SM clock sampled during the timed region (%.0f MHz)" % sm_mhz,
               "algorithmic": "6.5 wavefronts per warp and tree (3.5 feature -- the root's load serves two adjacent windows --, 2 node, 1 leaf load) x N*W*T/32 warp-trees",
               "traffic": f4.get("dram_bytes"), "hbm": {"achieved": kernels["K4b_gbt_smooth_tile"]["gbs"], "peak": peak, "unit": "GB/s",
                                                        "frac": kernels["K4b_gbt_smooth_tile"]["frac_hbm"], "algorithmic_bytes": k4b_bytes},
               "ncu": f4 or None,
               "note": "the smoother dominates the step and is bound by shared-memory wavefronts and issue slots, not by HBM "
                       "(HBM fraction shown for completeness); the HBM-bound kernel of the path is K1, see roofline_base"}
    roof_k1 = {"kernel": "K1_lr_tc_kernel", "bound": "hbm", "achieved": kernels["K1_lr_tc_kernel"]["gbs"], "peak": peak,
               "unit": "GB/s", "frac": kernels["K1_lr_tc_kernel"]["frac_hbm"], "traffic": f1.get("dram_bytes"),
               "peak_source": peak_src, "ncu": f1 or None}
    out = {
        "metric": "haplotypes/sec local-ancestry inference (Base->Smooth->argmax), chr1, 7-way",
        "value": value, "unit": "haplotypes/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "s8 x s8 -> s32 (exact fixed point) + f64 epilogue; f32 trees",
        "data": "synthetic",
        "config": {"workload": "chr1 C=%d M=%d W=%d A=%d S=%d, logistic base + XGB smoother, %d haplotypes per GPU resident in HBM"
                               % (C, M, W, A, S, N),
                   "forest": forest_kind, "l2": "inputs (%.1f GB/GPU) exceed L2, no flush needed" % (N * ld / 1e9),
                   "parallelism": "haplotype shards, one rank per GPU, no collective on the data path"},
        "e2e": e2e,
        "strong": strong,
        "gpu_launches": 3 * args.steps,
        "roofline": roof_k4 if dom.startswith("K4") else roof_k1,
        "roofline_base": roof_k1,
        "kernels": kernels,
        "clocks": clocks,
    }
    if sg:
        out["scatter_gather"] = sg
    if cpu:
        out["cpu_baseline"] = cpu
        out["parity"] = parity
    if world == 1 and not args.no_configs:
        cfgs = {}
        n_c = min(N, 20_000)
        del B, P, L
        for name, fn in (("cfg2_chr22_m1000_lr_xgb", lambda: config_chr22(peak)),
                         ("cfg4_chr1_covrsk_base", lambda: config_covrsk(X, ld, geom, min(N, args.covrsk_haps), peak)),
                         ("cfg5_chr1_lr_crf", lambda: config_crf(X, ld, base, geom, n_c)),
                         ("cfg5_chr1_lr_xgb_gnofix", lambda: config_gnofix(X, ld, base, smooth, geom, n_c)),
                         ("cfg5_chr1_lr_crf_gnofix_extension", lambda: config_crf_gnofix(X, ld, base, geom, min(n_c, args.crf_gnofix_haps)))):
            torch.cuda.empty_cache()
            try:
                cfgs[name] = fn()
            except Exception as e:  # a failing secondary config must not lose the headline line
                cfgs[name] = {"error": "%s: %s" % (type(e).__name__, e)}
        out["configs"] = cfgs
    print(json.dumps(out), file=_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--crf-gnofix-haps", type=int, default=4000, help="haplotypes of the CRF + Gnofix extension config")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--haps", type=int, default=50_000, help="haplotypes per GPU")
    ap.add_argument("--strong-haps", type=int, default=50_000, help="haplotypes in all for the strong-scaling block")
    ap.add_argument("--e2e-haps", type=int, default=16384)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-haps", type=int, default=1024)
    ap.add_argument("--covrsk-haps", type=int, default=8192)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-buffer measurements")
    args = ap.parse_args()
    # stdout carries ONE JSON line: whatever a library writes to file descriptor 1 on the way ("NCCL version ..." at
    # communicator creation) is sent to stderr, the line itself goes to the real stdout
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    _OUT.flush()


if __name__ == "__main__":
    main()
