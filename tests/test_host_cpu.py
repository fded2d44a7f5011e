"""CPU tests of the host side: C-ABI exports, plugin surface, geometry, pickling, the
no-CPU-fallback rule and the multi-rank sharding helpers (gloo, world_size 2)."""
import os
import pickle
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol(libgnx):
    hdr = open(os.path.join(ROOT, "include", "gnx.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(gnx_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    from gnomix_b200 import _lib
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (gnx_[a-z0-9_]+)", out))
    assert declared <= exported, declared - exported
    assert libgnx.gnx_version() == 120
    assert libgnx.gnx_release_workspace() == 0      # nothing allocated: a no-op that must not need a device


def test_no_cpu_fallback():
    """Without a GPU every compute entry point must fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gnomix_b200 import _lib
    from tests import util
    rng = np.random.default_rng(0)
    coefs, icpts, ctx = util.random_lr(rng, 1000, 100, 3)
    base = util.make_lr_base(1000, 100, 3, coefs, icpts)
    with pytest.raises(_lib.GnxError):
        base.predict_proba(np.zeros((2, 1000), dtype=np.int8))
    assert _lib.lib().gnx_device_count() == 0
    import ctypes as C
    out = C.c_void_p()
    rc = _lib.lib().gnx_lr_model_create(C.byref(out), 3, 1000, 100, 50, None, None, 0)
    assert rc != 0 and _lib.lib().gnx_last_error()


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under gnomix_b200/ may import, include, link or
    load it (comments may cite it)."""
    pat_py = re.compile(r"^\s*(from\s+oracle|import\s+oracle|from\s+\.\.?oracle)|libgnx_oracle|oracle/_", re.M)
    pat_c = re.compile(r"#\s*include\s*[<\"][^>\"]*oracle", re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gnomix_b200")):
        for f in files:
            path = os.path.join(dirpath, f)
            if f.endswith(".py"):
                assert not pat_py.search(open(path).read()), path
            elif f.endswith((".cu", ".cuh", ".h")):
                assert not pat_c.search(open(path).read()), path


def test_plugin_surface_matches_reference_signatures():
    import inspect
    from gnomix_b200 import Gnomix, Base, Smoother, LogisticRegressionBase, XGB_Smoother, CRF_Smoother, CovRSKBase
    assert list(inspect.signature(Base.__init__).parameters) == [
        "self", "chm_len", "window_size", "num_ancestry", "missing_encoding", "context", "train_admix", "n_jobs", "seed", "verbose"]
    assert list(inspect.signature(Smoother.__init__).parameters) == [
        "self", "n_windows", "num_ancestry", "smooth_window_size", "model", "calibrate", "n_jobs", "seed", "mode_filter", "verbose"]
    assert list(inspect.signature(Gnomix.__init__).parameters)[:5] == ["self", "C", "M", "A", "S"]
    for cls, names in ((Base, ["train", "predict_proba", "predict", "evaluate", "pad", "init_base_models"]),
                       (Smoother, ["train", "predict_proba", "predict", "evaluate", "process_base_proba"]),
                       (Gnomix, ["train", "predict", "predict_proba", "phase", "save", "write_config", "write_gen_map_df"])):
        for n in names:
            assert callable(getattr(cls, n)), (cls, n)
    m = Gnomix(C=5000, M=200, A=3, S=9)
    assert m.W == 25 and m.context == 100 and isinstance(m.base, LogisticRegressionBase) and isinstance(m.smooth, XGB_Smoother)
    assert m.smooth.gnofix is True and m.smooth.S == 9 and len(m.base.models) == 25
    assert isinstance(Gnomix(C=5000, M=200, A=3, S=9, mode="fast").smooth, CRF_Smoother)
    assert isinstance(Gnomix(C=5000, M=200, A=3, S=9, mode="best").base, CovRSKBase)
    assert Smoother(10, 3, smooth_window_size=8).S == 7  # S forced odd (src/Smooth/smooth.py:14)


def test_window_geometry_known_answers():
    """SURVEY.md 8(a) row A probe values for chr22 / M=857."""
    from gnomix_b200.base import Base
    from oracle import np_oracle as npo
    C, M = 317_408, 857
    b = Base(C, M, 7, context=428)
    sl = b.window_slices()
    assert len(sl) == 370 and sl == npo.base_window_ranges(C, M, 428)
    lo0, hi0 = sl[0]
    assert npo.padded_to_orig(np.array([lo0, hi0 - 1]), C, 428).tolist() == [427, 1284]
    lo, hi = sl[-1]
    assert hi - lo == 2031
    o = npo.padded_to_orig(np.array([lo, hi - 1]), C, 428)
    assert o.tolist() == [315_805, 316_980]
    X = np.arange(20, dtype=np.int8).reshape(1, 20)
    b2 = Base(20, 6, 2, context=3)
    assert np.array_equal(b2.pad(X), npo.base_pad(X, 3))


def test_model_pickle_roundtrip_drops_device_handles(tmp_path):
    from gnomix_b200 import Gnomix, GBTForest
    from tests import util
    rng = np.random.default_rng(0)
    C, M, A, S = 3000, 250, 3, 5
    m = Gnomix(C, M, A, S)
    coefs, icpts, ctx = util.random_lr(rng, C, M, A)
    m.base.set_window_weights(coefs, icpts)
    m.smooth.model = GBTForest.random(rng, A, m.smooth.S, n_rounds=3)
    m.base._handles["fake"] = object()
    blob = pickle.dumps(m)
    m2 = pickle.loads(blob)
    assert m2.base._handles == {} and m2.smooth.model._handles == {}
    c1, b1 = m.base.packed_weights()
    c2, b2 = m2.base.packed_weights()
    assert np.array_equal(c1, c2) and np.array_equal(b1, b2)
    assert np.array_equal(m2.smooth.model.thr, m.smooth.model.thr)


def test_partition_keeps_individuals_together():
    from gnomix_b200.parallel import partition
    for n in (0, 1, 2, 7, 16, 100_000, 99_999):
        for world in (1, 2, 3, 8):
            parts = partition(n, world)
            assert parts[0][0] == 0 and parts[-1][1] == n
            for (a, b), (c, d) in zip(parts, parts[1:]):
                assert b == c
            for a, b in parts:
                assert b == a or a % 2 == 0
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 3  # one individual + a trailing odd haplotype


_WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from gnomix_b200 import parallel
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% sys.argv[2], rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
N, C, W = 37, 501, 9
rng = np.random.default_rng(5)
X = torch.from_numpy(rng.integers(0, 3, size=(N, C), dtype=np.int8))
mine = parallel.scatter_rows(X if rank == 0 else None, N, C, torch.int8, "cpu")
lo, hi = parallel.local_rows(N, 2, rank)
assert mine.shape[0] == hi - lo and torch.equal(mine, X[lo:hi])
# stand-in for the per-rank hot path: a row-wise function of the shard
lab = torch.stack([mine[:, w::W].to(torch.int32).sum(1) for w in range(W)], 1)
full = parallel.gather_rows(lab, N)
if rank == 0:
    want = torch.stack([X[:, w::W].to(torch.int32).sum(1) for w in range(W)], 1)
    assert torch.equal(full, want)
    print("OK")
dist.destroy_process_group()
'''


def test_two_rank_scatter_gather_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER % ROOT)
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ps = [subprocess.Popen([sys.executable, str(script), str(r), str(port)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
          for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in ps]
    assert all(p.returncode == 0 for p in ps), outs
    assert "OK" in outs[0]


def test_calibrator_fit_matches_reference_semantics():
    """Calibrator.fit (src/Smooth/Calibration.py:41-55): one IsotonicRegression(out_of_bounds="clip") per
    class on the one-hot labels; the fitted thresholds are what the device transform (K7) consumes.
    transform itself is device-only (tests/test_calibrator_gpu.py); untrained -> input returned."""
    from sklearn.isotonic import IsotonicRegression
    from gnomix_b200.calibration import Calibrator
    rng = np.random.default_rng(0)
    A = 4
    y = rng.integers(0, A, 3000)
    proba = rng.dirichlet(np.ones(A), 3000).astype(np.float32)
    proba[np.arange(3000), y] += 0.5
    proba /= proba.sum(1, keepdims=True)
    cal = Calibrator(A)
    cal.fit(proba, y)
    for i, (xt, yt) in enumerate(cal.thresholds()):
        ref = IsotonicRegression(out_of_bounds="clip").fit(proba[:, i], (y == i).astype(float))
        assert xt.dtype == np.float32 and np.array_equal(xt, ref.X_thresholds_) and np.array_equal(yt, ref.y_thresholds_)
    test = rng.dirichlet(np.ones(3), (5, 40)).astype(np.float32)
    assert Calibrator(3).transform(test) is test  # untrained: returns the input with a warning
    import pickle
    cal2 = pickle.loads(pickle.dumps(cal))
    assert all(np.array_equal(a[0], b[0]) for a, b in zip(cal.thresholds(), cal2.thresholds()))


def test_crf_trainer_learns_sticky_chain():
    from gnomix_b200.crf_train import fit_crf, crf_nll_grad
    from oracle import np_oracle as npo
    rng = np.random.default_rng(1)
    A, N, W = 3, 60, 40
    y = np.zeros((N, W), dtype=int)
    for n in range(N):
        cur = rng.integers(A)
        for t in range(W):
            if rng.random() < 0.05:
                cur = rng.integers(A)
            y[n, t] = cur
    B = rng.dirichlet(np.ones(A) * 0.8, (N, W))
    B[np.arange(N)[:, None], np.arange(W)[None, :], y] += 0.4
    B /= B.sum(-1, keepdims=True)
    sw, tw = fit_crf(B, y, A, max_iterations=80)
    assert np.all(np.diag(tw) > tw.max(axis=1) - 1e-9)           # staying is the preferred transition
    proba, lab = npo.crf_smooth(B, sw, tw)
    assert (lab == y).mean() > (np.argmax(B, -1) == y).mean()    # smoothing beats the raw base argmax


def test_reference_import_paths_resolve_to_this_package():
    """gnomix.py and the reference's pickles name these modules (gnomix.py:11-21)."""
    import importlib
    import sys
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        if "/root/reference" in (getattr(sys.modules[k], "__file__", "") or ""):
            del sys.modules[k]
    saved = list(sys.path)
    sys.path[:] = [p for p in sys.path if "/root/reference" not in p]
    try:
        import gnomix_b200 as g
        assert importlib.import_module("src.model").Gnomix is g.Gnomix
        assert importlib.import_module("src.Base.models").LogisticRegressionBase is g.LogisticRegressionBase
        assert importlib.import_module("src.Base.models").CovRSKBase is g.CovRSKBase
        assert importlib.import_module("src.Smooth.models").XGB_Smoother is g.XGB_Smoother
        assert importlib.import_module("src.Smooth.models").CRF_Smoother is g.CRF_Smoother
        sw = importlib.import_module("src.Smooth.utils").slide_window
        from oracle import np_oracle as npo
        B = np.random.default_rng(0).dirichlet(np.ones(3), (2, 20))
        assert np.array_equal(sw(B, 5)[0], npo.slide_window(B, 5))
        assert callable(importlib.import_module("src.utils").vcf_to_npy) and callable(importlib.import_module("src.postprocess").write_msp)
        assert callable(importlib.import_module("src.Gnofix.gnofix").gnofix)
    finally:
        sys.path[:] = saved
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]


def test_xgboost_json_roundtrip_and_exported_model(tmp_path):
    """convert.py: a forest written in xgboost's JSON model schema parses back to the same arrays, and an
    exported-model .npz (the format scripts/export_reference_model.py writes) loads into a Gnomix."""
    import json
    from gnomix_b200 import GBTForest
    from gnomix_b200.convert import forest_from_xgboost_json, forest_to_xgboost_json, load_exported_model
    from tests import util
    rng = np.random.default_rng(3)
    A, S = 3, 5
    f = GBTForest.random(rng, A, S, n_rounds=4, depth=3)
    # make it ragged: turn one internal node of tree 0 into a leaf by rewriting the arrays through JSON
    js = forest_to_xgboost_json(f)
    txt = json.dumps(js)
    g = forest_from_xgboost_json(txt)
    for name in ("feat", "thr", "left", "right", "default_left", "leaf", "tree_offsets", "base_margin"):
        assert np.array_equal(getattr(f, name), getattr(g, name)), name
    p = tmp_path / "m.json"
    p.write_text(txt)
    assert forest_from_xgboost_json(str(p)).n_trees == f.n_trees
    # exported model file
    C, M = 3000, 250
    coefs, icpts, ctx = util.random_lr(rng, C, M, A)
    np.savez_compressed(tmp_path / "exp.npz", C=C, M=M, A=A, S=S, context=ctx, context_ratio=0.5, snp_pos=np.arange(C) * 10 + 5,
                        snp_ref=np.array(["A"] * C), snp_alt=np.array(["G"] * C), population_order=np.array(["X", "Y", "Z"]),
                        lr_coef=np.concatenate([c.ravel() for c in coefs]), lr_intercept=np.stack(icpts), xgb_json=np.array(txt),
                        gen_map_chm=np.array(["22", "22"]), gen_map_pos=np.array([1, 40000]), gen_map_cm=np.array([0.0, 1.5]))
    m = load_exported_model(str(tmp_path / "exp.npz"))
    assert m.W == C // M and m.population_order == ["X", "Y", "Z"] and m.smooth.model.n_trees == f.n_trees
    c2, b2 = m.base.packed_weights()
    assert np.array_equal(c2, np.concatenate([c.ravel() for c in coefs])) and np.array_equal(b2, np.stack(icpts))
    assert list(m.gen_map_df["pos"]) == [1, 40000]


def test_bench_reference_arm_prints_one_contract_line():
    """bench.py --impl reference runs without a GPU (the CPU port of the reference path) and prints exactly ONE JSON
    line on stdout carrying the keys the driver reads."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--cpu-haps", "8"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "haplotypes/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "haplotypes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["data"] == "synthetic"
