#!/usr/bin/env python
"""bench.py -- haplotypes/sec of the local-ancestry inference hot path
(Base.predict_proba -> Smoother.predict_proba + argmax) on B200, next to the CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (K1 logistic base for all windows, K4 tree-ensemble
smoother with fused argmax) over one batch of synthetic admixed haplotypes that is
already resident in HBM as int8.  Workload = BASELINE.json configs[2]: chr1 geometry
(C=1,226,139 SNPs, M=857, W=1430 windows, A=7, S=75), logistic base + XGB-semantics
smoother, 50,000 haplotypes PER GPU (weak scaling: every rank owns its own shard, no
collective on the data path).  `value` = haplotypes of all ranks / max-over-ranks step
time; `e2e` = the same metric through the host-buffer C-ABI entry point gnx_infer_host
(pinned host int8 in, labels out, H2D/D2H inside the timed region; by default the host
cores pack part of every chunk to 2 bits per SNP while the DMA engine moves the rest raw --
`e2e.unpacked` is the same call with packing off).

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "chr1"
SEED = 94305  # reference default seed, config.yaml:2


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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


def build_models(geom, want_device=True):
    """Synthetic model of the workload's architecture: discriminant LR weights per window +
    a 100-round x A-tree depth-4 forest (trained fixture if present, else seeded random)."""
    from gnomix_b200 import synth
    from gnomix_b200.base import LogisticRegressionBase, Base
    from gnomix_b200.smooth import XGB_Smoother
    from gnomix_b200.gbt import GBTForest
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
    fixture = os.path.join(ROOT, "gnomix_b200", "data", "forest_%s.npz" % WORKLOAD)
    if os.path.exists(fixture):
        forest = GBTForest.from_npz_dict(np.load(fixture))
        forest_kind = "HGB-trained fixture (xgboost hyper-parameters), gnomix_b200/data/forest_%s.npz" % WORKLOAD
    else:
        forest = GBTForest.random(np.random.default_rng(SEED + 1), A, S, n_rounds=100, depth=4)
        forest_kind = "seeded random complete forest, 100 rounds x %d trees, depth 4" % A
    smooth = XGB_Smoother(n_windows=C // M, num_ancestry=A, smooth_window_size=S)
    smooth.model = forest
    return base, smooth, (fx, fpop), (coefs, icpts, ctx), forest_kind


def cpu_path(X, coefs, icpts, geom, ctx, forest):
    """The reference's CPU path restated (oracle/): per-window float64 GEMM + expit +
    normalise (what sklearn's predict_proba does under Base.predict_proba_vectorized),
    then slide_window + tree predictor + argmax with OpenMP (what xgboost does)."""
    from oracle import np_oracle as npo, c_oracle as co
    C, M, A, S, _ = geom
    B = npo.lr_base_predict_proba(X, coefs, icpts, C, M, ctx)
    proba, label = co.gbt_smooth(forest, B.astype(np.float32), S)
    return proba, label


def time_cpu(X_sample, coefs, icpts, geom, ctx, forest, budget_s=15.0):
    n0 = min(64, len(X_sample))
    t = time.perf_counter()
    cpu_path(X_sample[:n0], coefs, icpts, geom, ctx, forest)
    dt = time.perf_counter() - t
    n = int(max(n0, min(len(X_sample), n0 * budget_s / max(dt, 1e-3))))
    t = time.perf_counter()
    cpu_path(X_sample[:n], coefs, icpts, geom, ctx, forest)
    dt = time.perf_counter() - t
    return n / dt, n, dt


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
    }))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from gnomix_b200 import synth, _lib

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

    # ---- resident inputs: this rank's shard, generated on the device -------------
    fdev = torch.from_numpy(fx).cuda()
    ld = (C + 127) // 128 * 128
    X = synth.admix_device(fdev, N, morgans, seed=SEED + 1000 * rank, ld=ld)
    B = torch.empty((N, W, A), dtype=torch.float32, device="cuda")
    P = torch.empty((N, W, A), dtype=torch.float32, device="cuda")
    L = torch.empty((N, W), dtype=torch.int32, device="cuda")
    hlr, hgbt = base.handle(), smooth.model.handle(S)
    st = torch.cuda.current_stream().cuda_stream

    def step():
        _lib.check(lib.gnx_lr_predict(hlr, X.data_ptr(), N, ld, B.data_ptr(), st), "gnx_lr_predict")
        _lib.check(lib.gnx_gbt_smooth(hgbt, B.data_ptr(), N, W, P.data_ptr(), L.data_ptr(), st), "gnx_gbt_smooth")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 1)]
    barrier()
    ev[0].record()
    for i in range(args.steps):
        _lib.check(lib.gnx_lr_predict(hlr, X.data_ptr(), N, ld, B.data_ptr(), st), "gnx_lr_predict")
        ev[2 * i + 1].record()
        _lib.check(lib.gnx_gbt_smooth(hgbt, B.data_ptr(), N, W, P.data_ptr(), L.data_ptr(), st), "gnx_gbt_smooth")
        ev[2 * i + 2].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = ev[0].elapsed_time(ev[-1])
    k1_ms = float(np.mean([ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(args.steps)]))
    k4_ms = float(np.mean([ev[2 * i + 1].elapsed_time(ev[2 * i + 2]) for i in range(args.steps)]))
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = N * world / (ms_per_step * 1e-3)

    # ---- e2e: host buffers through gnx_infer_host --------------------------------
    # 8 ranks of one box share the host: keep the pinned e2e batch at 10 GB per rank there
    n_e2e = min(N, args.e2e_haps if world == 1 else min(args.e2e_haps, 8192))
    Xh = torch.empty((n_e2e, ld), dtype=torch.int8, pin_memory=True)
    Xh.copy_(X[:n_e2e])
    Lh = torch.empty((n_e2e, W), dtype=torch.int32, pin_memory=True)
    torch.cuda.synchronize()

    import ctypes as C_

    def e2e_step():
        _lib.check(lib.gnx_infer_host(hlr, hgbt, Xh.data_ptr(), n_e2e, ld, None, Lh.data_ptr(), 0), "gnx_infer_host")

    def e2e_measure():
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / args.e2e_steps
        te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        frac, h2d, d2h = C_.c_double(0), C_.c_int64(0), C_.c_int64(0)
        lib.gnx_infer_host_last_transfer(C_.byref(frac), C_.byref(h2d), C_.byref(d2h))
        return n_e2e * world / float(te.item()), frac.value, h2d.value, d2h.value, bool(torch.equal(Lh.cuda(), L[:n_e2e]))

    # default path: part of every chunk crosses PCIe as 2-bit planes packed by the host cores
    e2e_value, e2e_frac, e2e_h2d, e2e_d2h, labels_match = e2e_measure()
    pk, h2dr = C_.c_double(0), C_.c_double(0)
    lib.gnx_infer_host_rates(C_.byref(pk), C_.byref(h2dr))
    # for comparison: the same call with packing switched off (raw int8 over PCIe)
    os.environ["GNX_HOST_PACK"] = "0"
    raw_value, _, raw_h2d, _, raw_match = e2e_measure()
    del os.environ["GNX_HOST_PACK"]
    labels_match = labels_match and raw_match

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline on a bounded sample of the same haplotypes (rank 0, N=1 only) --
    cpu = None
    if world == 1 and not args.no_cpu:
        from oracle import c_oracle as co
        co.use_all_cores()
        ns = min(N, args.cpu_haps)
        Xs = X[:ns, :C].cpu().numpy()
        v, n_used, dt = time_cpu(Xs, coefs, icpts, geom, ctx, smooth.model)
        # parity of the bounded sample while we are here: labels identical, proba within 1e-5
        p_cpu, l_cpu = cpu_path(Xs[:64], coefs, icpts, geom, ctx, smooth.model)
        lab_ok = bool(np.array_equal(l_cpu, L[:64].cpu().numpy()))
        p_err = float(np.max(np.abs(p_cpu - P[:64].cpu().numpy())))
        cpu = {"value": v, "unit": "haplotypes/s", "cores": co.num_threads(), "kind": "port",
               "sample": "%d haplotypes of the same workload in %.1f s (numpy float64 per-window GEMM + OpenMP tree predictor)" % (n_used, dt),
               "labels_match_gpu_on_64": lab_ok, "max_abs_proba_diff_on_64": p_err}

    peak, peak_src = _peaks()
    k1_bytes = N * (C + W * A * 4)
    k4_bytes = N * (2 * W * A * 4 + W * 4)
    kernels = {
        "K1_lr_tc_kernel": {"ms": k1_ms, "algorithmic_bytes": k1_bytes, "gbs": k1_bytes / k1_ms / 1e6,
                            "frac_hbm": k1_bytes / k1_ms / 1e6 / peak},
        "K4_gbt_smooth_kernel": {"ms": k4_ms, "algorithmic_bytes": k4_bytes, "gbs": k4_bytes / k4_ms / 1e6,
                                 "frac_hbm": k4_bytes / k4_ms / 1e6 / peak,
                                 "tree_traversals_per_s": N * W * smooth.model.n_trees / (k4_ms * 1e-3)},
    }
    dom = "K1_lr_tc_kernel" if k1_ms >= k4_ms else "K4_gbt_smooth_kernel"
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture of this
    # exact command (profiles/r1_ncu_final_summary.txt); only valid for the default workload size
    traffic = {"K1_lr_tc_kernel": 68.05e9, "K4_gbt_smooth_kernel": 4.25e9} if (N == 50_000 and WORKLOAD == "chr1") else {}
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
        "e2e": {"value": e2e_value, "unit": "haplotypes/s", "h2d_bytes_per_step": int(e2e_h2d),
                "d2h_bytes_per_step": int(e2e_d2h), "haplotypes_per_step": n_e2e,
                "api": "gnx_infer_host (pinned host int8 in, int32 labels out)", "labels_match_resident_path": labels_match,
                "packed_fraction_of_rows": e2e_frac, "host_threads": int(lib.gnx_host_threads()),
                "calibrated_host_pack_gbs": pk.value, "calibrated_h2d_gbs": h2dr.value,
                "unpacked": {"value": raw_value, "h2d_bytes_per_step": int(raw_h2d)}},
        "gpu_launches": 2 * args.steps,
        "roofline": {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["gbs"], "peak": peak, "unit": "GB/s",
                     "frac": kernels[dom]["frac_hbm"], "traffic": traffic.get(dom), "peak_source": peak_src,
                     # what actually bounds K4 (ncu --set full of this command, profiles/r1_ncu_final_summary.txt)
                     "issue_slot_frac": 0.75 if dom.startswith("K4") else None,
                     "shared_wavefront_frac": 0.84 if dom.startswith("K4") else None,
                     "note": "the smoother (K4) dominates the step and is bound by issue slots / shared-memory wavefronts "
                             "(ncu: LSU shared wavefronts 84 % of peak, issue-active 75 %), not by HBM; the HBM-bound kernel "
                             "of the path is K1, see roofline_base"},
        "roofline_base": {"kernel": "K1_lr_tc_kernel", "bound": "hbm", "achieved": kernels["K1_lr_tc_kernel"]["gbs"], "peak": peak,
                          "unit": "GB/s", "frac": kernels["K1_lr_tc_kernel"]["frac_hbm"], "traffic": traffic.get("K1_lr_tc_kernel"),
                          "peak_source": peak_src},
        "kernels": kernels,
        "clocks": clocks,
    }
    if cpu:
        out["cpu_baseline"] = cpu
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--haps", type=int, default=50_000, help="haplotypes per GPU")
    ap.add_argument("--e2e-haps", type=int, default=16384)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-haps", type=int, default=1024)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
