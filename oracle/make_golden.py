"""Generates tests/golden/*.npz by running the REFERENCE's own Python (read-only tree at
/root/reference, imported with stub modules -- oracle/refimport.py) on small seeded
inputs.  TEST INFRASTRUCTURE ONLY.  Run in the build container:

    python -m oracle.make_golden

The GPU box has no /root/reference; the committed vectors are what travels.  Every
fixture stores its inputs next to the reference's outputs so that the oracle
restatement (tests -m "not gpu") and the CUDA path (tests -m gpu) are checked against
the same numbers.

What each file pins (reference file:line):
  base_lr_*.npz     Base.pad + Base.predict_proba_vectorized windowing (src/Base/base.py:
                    41-44,146-180) with scikit-learn liblinear logistic models trained
                    through the reference's own Base.train_vectorized (99-127).
  slide_window.npz  slide_window (src/Smooth/utils.py:4-29).
  covrsk.npz        CovSample / CovRSK_DP_triangular_numbers (src/Base/string_kernel.py:
                    80-111) and CovRSKBase -> sklearn SVC(probability=True) (src/Base/
                    models.py:195-215) through Base.predict_proba_vectorized.
  gnofix.npz        gnofix + track_switch + correct_phase_error (src/Gnofix/gnofix.py:
                    58-208, src/Gnofix/phasing.py:182-198) and Smoother.predict
                    (src/Smooth/smooth.py:40-65) driven with the oracle's tree predictor as
                    smoother.model (xgboost itself is not installable here).
  meta.npz          get_meta_data (src/postprocess.py:25-67).
  lai_bed.npz       msp_to_lai / msp_to_bed (src/postprocess.py:128-210) on the .msp of meta.npz.
  calibrator.npz    Calibrator.fit / transform (src/Smooth/Calibration.py:19-69) through
                    scikit-learn IsotonicRegression.
"""
from __future__ import annotations

import os
import types
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")


def _structured_haplotypes(rng, n_per_pop, A, C):
    freqs = np.clip(rng.beta(0.5, 0.5, size=(A, C)), 0.05, 0.95)
    X = np.concatenate([(rng.random((n_per_pop, C)) < freqs[a]).astype(np.int8) for a in range(A)])
    pop = np.repeat(np.arange(A), n_per_pop)
    return X, pop, freqs


def golden_base_lr(src, name, C, M, A, seed, n_per_pop=24, n_query=20):
    from sklearn.linear_model import LogisticRegression
    from sklearn.multiclass import OneVsRestClassifier
    from src.Base.base import Base
    rng = np.random.default_rng(seed)
    ctx = int(M * 0.5)
    Xt, pop, freqs = _structured_haplotypes(rng, n_per_pop, A, C)
    W = C // M
    yt = np.repeat(pop[:, None], W, axis=1)
    base = Base(chm_len=C, window_size=M, num_ancestry=A, context=ctx)
    # src/Base/models.py:19-21 with the one change scikit-learn >= 1.8 forces for A > 2:
    # liblinear is one-vs-rest by construction, spelled explicitly (SURVEY.md headline 3)
    if A == 2:
        factory = lambda: LogisticRegression(penalty="l2", C=3., solver="liblinear", max_iter=1000)
    else:
        factory = lambda: OneVsRestClassifier(LogisticRegression(penalty="l2", C=3., solver="liblinear", max_iter=1000))
    base.init_base_models(factory)
    base.base_multithread = False
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        base.train(Xt, yt)
        # queries: mosaics of training haplotypes + missing calls
        Xq = Xt[rng.integers(0, len(Xt), n_query)].copy()
        cut = rng.integers(1, C - 1, n_query)
        other = Xt[rng.integers(0, len(Xt), n_query)]
        for i in range(n_query):
            Xq[i, cut[i]:] = other[i, cut[i]:]
        Xq[rng.random(Xq.shape) < 0.02] = 2
        B = base.predict_proba(Xq)
    coefs, icpts = [], []
    for mdl in base.models:
        if hasattr(mdl, "estimators_"):
            coefs.append(np.concatenate([e.coef_ for e in mdl.estimators_], axis=0))
            icpts.append(np.concatenate([e.intercept_ for e in mdl.estimators_]))
        else:
            coefs.append(mdl.coef_)
            icpts.append(mdl.intercept_)
    flat = np.concatenate([c.ravel() for c in coefs])
    np.savez_compressed(os.path.join(OUT, name), C=C, M=M, A=A, ctx=ctx, X=Xq, B=B, coef_flat=flat,
                        intercepts=np.stack(icpts), padded=base.pad(Xq)[:2])
    print(name, "B", B.shape, B.dtype)


def golden_slide_window(src):
    from src.Smooth.utils import slide_window
    rng = np.random.default_rng(5)
    out = {}
    for tag, (N, W, A, S) in {"small": (3, 12, 4, 5), "s75": (2, 160, 7, 75), "s9": (2, 20, 3, 9)}.items():
        B = rng.dirichlet(np.ones(A), size=(N, W))
        Xs, _ = slide_window(B, S)
        out["B_" + tag] = B
        out["S_" + tag] = S
        out["X_" + tag] = Xs
    np.savez_compressed(os.path.join(OUT, "slide_window.npz"), **out)
    print("slide_window", {k: v.shape for k, v in out.items() if k.startswith("X_")})


def golden_covrsk(src):
    from src.Base import string_kernel as sk
    from src.Base.models import CovRSKBase
    rng = np.random.default_rng(11)
    out = {"Ms_2500": np.array(sk.CovSample(2500, 0.6, 1.0, 37)), "Ms_300": np.array(sk.CovSample(300, 0.6, 1.0, 37))}
    # raw kernel values incl. missing calls and long identical stretches
    Mlen = 450
    Y = rng.integers(0, 2, size=(12, Mlen)).astype(np.int8)
    X = Y[rng.integers(0, 12, 9)].copy()
    X[rng.random(X.shape) < 0.03] ^= 1
    X[rng.random(X.shape) < 0.02] = 2
    Y[rng.random(Y.shape) < 0.02] = 2
    X[0] = Y[0]
    out["K_X"], out["K_Y"] = X, Y
    out["K"] = sk.CovRSK_DP_triangular_numbers(X, Y)
    # full base: CovRSKBase (M < 500 -> single-process kernel), sklearn SVC(probability=True)
    C, M, A = 1130, 200, 3
    ctx = int(M * 0.5)
    W = C // M
    Xt, pop, _ = _structured_haplotypes(rng, 14, A, C)
    yt = np.repeat(pop[:, None], W, axis=1)
    # CovRSKBase.__init__ asserts `int(np.__version__.split(".")[1]) >= 20`, which numpy 2.x
    # fails on its minor number; the rest of the constructor (src/Base/models.py:202-215) is
    # replayed verbatim on an instance made without it.
    from sklearn import svm
    from src.Base.base import Base
    base = CovRSKBase.__new__(CovRSKBase)
    Base.__init__(base, chm_len=C, window_size=M, num_ancestry=A, context=ctx)
    base.train_admix = False
    base.kernel = sk.CovRSK_DP_triangular_numbers          # M < 500 branch
    base.init_base_models(lambda: svm.SVC(kernel=base.kernel, probability=True))
    base.base_multithread = False
    base.log_inference = False
    np.random.seed(1)
    base.train(Xt, yt)
    Xq = Xt[rng.integers(0, len(Xt), 10)].copy()
    Xq[rng.random(Xq.shape) < 0.05] ^= 1
    Xq[rng.random(Xq.shape) < 0.02] = 2
    B = base.predict_proba(Xq)
    out.update(svc_C=C, svc_M=M, svc_A=A, svc_ctx=ctx, svc_X=Xq, svc_B=B, svc_Xtrain=Xt)
    for w, mdl in enumerate(base.models):
        out["svc_w%d_support" % w] = mdl.support_
        out["svc_w%d_n_support" % w] = mdl.n_support_
        out["svc_w%d_dual_coef" % w] = mdl._dual_coef_
        out["svc_w%d_intercept" % w] = mdl._intercept_
        out["svc_w%d_probA" % w] = mdl.probA_ if hasattr(mdl, "probA_") else mdl._probA
        out["svc_w%d_probB" % w] = mdl.probB_ if hasattr(mdl, "probB_") else mdl._probB
    np.savez_compressed(os.path.join(OUT, "covrsk.npz"), **out)
    print("covrsk K", out["K"].shape, "svc B", B.shape, "Ms", out["Ms_2500"])


class _OracleForestModel:
    """smoother.model stand-in: predict_proba = the oracle's xgboost-semantics predictor."""

    def __init__(self, forest):
        self.forest = forest

    def predict_proba(self, rows):
        from oracle import c_oracle as co
        return co.gbt_rows(self.forest, np.asarray(rows, dtype=np.float32))


def _random_forest(rng, A, S, rounds, depth=4):
    from oracle.np_oracle import GBTModel
    F = S * A
    feat, thr, left, right, dl, leaf, offs = [], [], [], [], [], [], [0]
    n_split = 2 ** depth - 1
    for _ in range(rounds * A):
        for i in range(n_split):
            feat.append(int(rng.integers(0, F))); thr.append(float(rng.random() * 0.6)); left.append(2 * i + 1)
            right.append(2 * i + 2); dl.append(0); leaf.append(0.0)
        for i in range(2 ** depth):
            feat.append(-1); thr.append(0.0); left.append(0); right.append(0); dl.append(0)
            leaf.append(float(rng.normal() * 0.3))
        offs.append(len(feat))
    return GBTModel(A, F, feat, thr, left, right, dl, leaf, offs, np.full(A, 0.5, dtype=np.float32))


def golden_gnofix(src):
    from src.Smooth.models import XGB_Smoother
    from src.Gnofix.gnofix import gnofix
    rng = np.random.default_rng(23)
    W, A, S, C = 60, 3, 11, 60 * 37 + 13
    forest = _random_forest(rng, A, S, rounds=12)
    smoother = XGB_Smoother(n_windows=W, num_ancestry=A, smooth_window_size=S)
    smoother.model = _OracleForestModel(forest)
    out = dict(W=W, A=A, S=S, C=C, feat=forest.feat, thr=forest.thr, left=forest.left, right=forest.right,
               default_left=forest.default_left, leaf=forest.leaf, tree_offsets=forest.tree_offsets,
               base_margin=forest.base_margin)
    n_ind = 6
    Xs, Bs, Xo, Yo, Tr = [], [], [], [], []
    for i in range(n_ind):
        # two ancestry tracks with phase switch errors planted in B and X
        anc = np.zeros((2, W), dtype=int)
        for h in range(2):
            cuts = np.sort(rng.integers(1, W, 2))
            anc[h, cuts[0]:cuts[1]] = rng.integers(0, A)
            anc[h, cuts[1]:] = rng.integers(0, A)
        B = np.full((2, W, A), 0.1) + rng.random((2, W, A)) * 0.15
        for h in range(2):
            B[h, np.arange(W), anc[h]] += 0.7
        for sw in np.sort(rng.integers(2, W - 2, 3)):
            B[:, sw:] = B[::-1, sw:].copy()
        B /= B.sum(-1, keepdims=True)
        X = rng.integers(0, 2, size=(2, C)).astype(np.int8)
        X_m, X_p, Y_m, Y_p, history, tracker = gnofix(X[0], X[1], B, smoother)
        Xs.append(X); Bs.append(B); Xo.append(np.array([X_m, X_p])); Yo.append(np.array([Y_m, Y_p]))
        Tr.append(np.array(tracker))
    out.update(X=np.array(Xs), B=np.array(Bs), X_out=np.array(Xo), Y_out=np.array(Yo), tracker=np.array(Tr))
    # the smoother's own predict on the stacked inputs (Smoother.predict -> slide_window -> model)
    out["Y_smooth"] = smoother.predict(np.array(Bs).reshape(-1, W, A))
    out["P_smooth"] = smoother.predict_proba(np.array(Bs).reshape(-1, W, A))
    np.savez_compressed(os.path.join(OUT, "gnofix.npz"), **out)
    n_sw = [int((np.array(t)[0][:-1] != np.array(t)[0][1:]).sum()) for t in Tr]
    print("gnofix: switches per individual", n_sw)


def golden_meta(src):
    """get_meta_data + write_msp + write_fb (src/postprocess.py:25-126) on a small seeded case;
    the written files are stored byte for byte."""
    import tempfile
    import pandas as pd
    from src.postprocess import get_meta_data, write_msp, write_fb
    rng = np.random.default_rng(3)
    C, M, A = 5231, 400, 3
    W = C // M
    n_ind = 4
    pos = np.sort(rng.choice(np.arange(16_000_000, 51_000_000), size=C, replace=False))
    # query has a subset of the model SNPs plus a few extra positions
    qpos = np.sort(np.concatenate([rng.choice(pos, size=C - 300, replace=False), rng.choice(np.arange(16_000_000, 51_000_000), 50)]))
    gm_pos = np.sort(rng.choice(np.arange(17_000_000, 50_000_000), 300, replace=False))
    gm_cm = np.sort(rng.random(300) * 70.0)
    gen_map_df = pd.DataFrame({"chm": ["22"] * 300, "pos": gm_pos, "pos_cm": gm_cm})
    meta = get_meta_data("22", pos, qpos, W, M, gen_map_df)
    labels = rng.integers(0, A, size=(2 * n_ind, W))
    proba = rng.dirichlet(np.ones(A), size=(2 * n_ind, W)).astype(np.float32)
    pops = ["AFR", "EUR", "EAS"]
    samples = ["S%d" % i for i in range(n_ind)]
    with tempfile.TemporaryDirectory() as td:
        write_msp(os.path.join(td, "q"), meta, labels, pops, samples)
        write_fb(os.path.join(td, "q"), meta, proba, pops, samples)
        msp_bytes = open(os.path.join(td, "q.msp"), "rb").read()
        fb_bytes = open(os.path.join(td, "q.fb"), "rb").read()
    np.savez_compressed(os.path.join(OUT, "meta.npz"), C=C, M=M, W=W, A=A, pos=pos, qpos=qpos, gm_pos=gm_pos, gm_cm=gm_cm,
                        columns=np.array(list(meta.columns)), table=np.asarray(meta.values).astype(str),
                        labels=labels, proba=proba, pops=np.array(pops), samples=np.array(samples),
                        msp=np.frombuffer(msp_bytes, dtype=np.uint8), fb=np.frombuffer(fb_bytes, dtype=np.uint8))
    print("meta", meta.shape, list(meta.columns), len(msp_bytes), len(fb_bytes))


def golden_vcf_to_npy(src):
    """vcf_to_npy / snp_intersection (src/utils.py:83-159): pure numpy, driven with a
    read_vcf-shaped dict (scikit-allel itself is not installable here)."""
    from src.utils import vcf_to_npy
    rng = np.random.default_rng(8)
    n_ind, Cm = 5, 400
    model_pos = np.sort(rng.choice(np.arange(1000, 90000), Cm, replace=False))
    model_ref = rng.choice(np.array(["A", "C", "G", "T"]), Cm)
    keep = np.sort(rng.choice(Cm, 330, replace=False))
    extra = np.setdiff1d(rng.choice(np.arange(1000, 90000), 40), model_pos)
    vpos = np.concatenate([model_pos[keep], extra])
    vref = np.concatenate([model_ref[keep], rng.choice(np.array(["A", "C", "G", "T"]), len(extra))])
    flip = rng.random(len(keep)) < 0.1
    vref[:len(keep)][flip] = "N"
    order = np.argsort(vpos)
    vpos, vref = vpos[order], vref[order]
    gt = rng.integers(0, 2, size=(len(vpos), n_ind, 2)).astype(np.int8)
    gt[rng.random(gt.shape) < 0.02] = -1
    vcf = {"calldata/GT": gt, "variants/POS": vpos, "variants/REF": vref}
    X = vcf_to_npy(vcf, model_pos, model_ref, verbose=False)
    X0 = vcf_to_npy({"calldata/GT": gt.copy(), "variants/POS": vpos, "variants/REF": vref}, None, None, verbose=False)
    np.savez_compressed(os.path.join(OUT, "vcf_to_npy.npz"), gt=gt, vpos=vpos, vref=vref, model_pos=model_pos, model_ref=model_ref,
                        X=X, X_noformat=X0)
    print("vcf_to_npy", X.shape, X.dtype, np.bincount(X.ravel()))


def golden_lai_bed(src):
    """msp_to_lai + msp_to_bed (src/postprocess.py:128-210) on the .msp of meta.npz; the written
    files are stored byte for byte."""
    import tempfile
    from src.postprocess import msp_to_lai, msp_to_bed
    d = np.load(os.path.join(OUT, "meta.npz"))
    out = {}
    with tempfile.TemporaryDirectory() as td:
        msp = os.path.join(td, "q.msp")
        open(msp, "wb").write(d["msp"].tobytes())
        msp_to_lai(msp_file=msp, positions=d["qpos"], lai_file=os.path.join(td, "q.lai"))
        out["lai"] = np.frombuffer(open(os.path.join(td, "q.lai"), "rb").read(), dtype=np.uint8)
        for tag, pop_order in (("num", None), ("pop", d["pops"].tolist())):
            root = os.path.join(td, "bed_" + tag)
            os.makedirs(root)
            msp_to_bed(msp_file=msp, root=root, pop_order=pop_order)
            names = sorted(os.listdir(root))
            out["bed_%s_names" % tag] = np.array(names)
            for i, nm in enumerate(names):
                out["bed_%s_%d" % (tag, i)] = np.frombuffer(open(os.path.join(root, nm), "rb").read(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "lai_bed.npz"), **out)
    print("lai_bed", len(out["lai"]), list(out["bed_num_names"]))


def golden_calibrator(src):
    """Calibrator.fit / transform (src/Smooth/Calibration.py:19-69) with scikit-learn's
    IsotonicRegression: float32 probabilities (what the XGB smoother returns), float64
    (the CRF smoother), and the binary case; queries include exact threshold hits, values outside
    the fitted range (clip) and rows whose calibrated probabilities are all zero (NaN -> 1/A)."""
    from src.Smooth.Calibration import Calibrator
    out = {}
    for tag, A, dt, seed in [("a7_f32", 7, np.float32, 11), ("a3_f64", 3, np.float64, 12), ("a2_f32", 2, np.float32, 13)]:
        rng = np.random.default_rng(seed)
        n = 4000
        y = rng.integers(0, A, n)
        p = rng.dirichlet(np.full(A, 0.4), n)
        p[np.arange(n), y] += rng.random(n) * 1.5
        p = (p / p.sum(1, keepdims=True)).astype(dt)
        cal = Calibrator(A)
        cal.fit(p, y)
        q = rng.dirichlet(np.full(A, 0.3), (6, 50)).astype(dt)
        thr = [(m.X_thresholds_, m.y_thresholds_) for m in cal.models]
        for i in range(A):                       # exact hits, one ulp either side, out of range
            xt = thr[i][0]
            q[0, :8, i] = xt[rng.integers(0, len(xt), 8)]
            q[1, :8, i] = np.nextafter(xt[rng.integers(0, len(xt), 8)], dt(2))
            q[1, 8:16, i] = np.nextafter(xt[rng.integers(0, len(xt), 8)], dt(-1))
        q[2, 0, :] = 0.0
        q[2, 1, :] = 1.0
        q[2, 2, :] = -0.5
        q[2, 3, :] = 1.5
        got = cal.transform(q.copy())
        out["q_" + tag] = q
        out["out_" + tag] = got
        for i in range(A):
            out["x_%s_%d" % (tag, i)] = thr[i][0]
            out["y_%s_%d" % (tag, i)] = thr[i][1]
        print("calibrator", tag, got.dtype, [len(t[0]) for t in thr], float(np.nanmin(got)), float(np.nanmax(got)))
    np.savez_compressed(os.path.join(OUT, "calibrator.npz"), **out)


def golden_mode_filter(src):
    """The reference's own mode_filter (src/Smooth/utils.py:31-46) on random label rows.  scipy >= 1.11 returns
    scalars from stats.mode; the reference indexes `[0][0]` (scipy 1.5.3), so the call is given keepdims=True."""
    from scipy import stats
    su = sys.modules["src.Smooth.utils"]
    orig = stats.mode
    su.stats = types.SimpleNamespace(mode=lambda a: orig(a, keepdims=True))
    rng = np.random.default_rng(11)
    out = {}
    for tag, (n, W, A, size) in {"a": (6, 40, 3, 5), "b": (4, 33, 7, 9), "c": (3, 20, 2, 4), "d": (2, 12, 4, 1), "e": (2, 7, 3, 11)}.items():
        y = rng.integers(0, A, (n, W))
        # runs of equal labels, as real predictions have them
        y = np.where(rng.random((n, W)) < 0.6, np.roll(y, 1, axis=1), y)
        got = np.apply_along_axis(func1d=su.mode_filter, axis=1, arr=y, size=size)
        out["y_" + tag], out["size_" + tag], out["A_" + tag], out["out_" + tag] = y, size, A, got
        print("mode_filter", tag, int((got != y).sum()), "changed")
    su.stats = stats
    np.savez_compressed(os.path.join(OUT, "mode_filter.npz"), **out)


def main():
    from oracle import refimport
    src = refimport.import_reference()
    os.makedirs(OUT, exist_ok=True)
    golden_base_lr(src, "base_lr_a3.npz", C=2311, M=200, A=3, seed=1)
    golden_base_lr(src, "base_lr_a7.npz", C=4205, M=300, A=7, seed=2, n_per_pop=12)
    golden_base_lr(src, "base_lr_a2.npz", C=1501, M=250, A=2, seed=3)
    golden_slide_window(src)
    golden_covrsk(src)
    golden_gnofix(src)
    golden_meta(src)
    golden_vcf_to_npy(src)
    golden_calibrator(src)
    golden_lai_bed(src)
    golden_mode_filter(src)


if __name__ == "__main__":
    main()
