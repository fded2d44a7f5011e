"""CPU tests: the oracle restatement (oracle/np_oracle.py, oracle/gnx_oracle.c) against the
golden vectors produced by the reference's own Python (oracle/make_golden.py) and against
independent third-party implementations available offline (scikit-learn)."""
import math
import os

import numpy as np
import pytest

from oracle import np_oracle as npo, c_oracle as co
from tests import util

G = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return np.load(os.path.join(G, name), allow_pickle=False)


def _split_coefs(d):
    C, M, A, ctx = int(d["C"]), int(d["M"]), int(d["A"]), int(d["ctx"])
    rows = 1 if A == 2 else A
    coefs, o = [], 0
    for lo, hi in npo.base_window_ranges(C, M, ctx):
        n = rows * (hi - lo)
        coefs.append(d["coef_flat"][o:o + n].reshape(rows, hi - lo))
        o += n
    assert o == len(d["coef_flat"])
    return C, M, A, ctx, coefs, [b for b in d["intercepts"]]


@pytest.mark.parametrize("name", ["base_lr_a3.npz", "base_lr_a7.npz", "base_lr_a2.npz"])
def test_lr_base_against_reference(name):
    d = _load(name)
    C, M, A, ctx, coefs, icpts = _split_coefs(d)
    X, B_ref = d["X"], d["B"]
    # Base.pad index map
    assert np.array_equal(npo.base_pad(X[:2], ctx), d["padded"])
    assert np.array_equal(X[:2][:, npo.padded_to_orig(np.arange(C + 2 * ctx), C, ctx)], d["padded"])
    # float64 restatements (numpy and C) of the reference's windowing + sklearn predict_proba
    B_np = npo.lr_base_predict_proba(X, coefs, icpts, C, M, ctx)
    assert B_np.shape == B_ref.shape
    assert np.max(np.abs(B_np - B_ref)) < 1e-13
    B_c = co.lr_f64(X, coefs, np.stack(icpts), C, M, ctx, A)
    assert np.max(np.abs(B_c - B_ref)) < 1e-12
    # the exact fixed-point form the CUDA kernel evaluates
    (Bf, Bd), s = util.oracle_lr_fixed(X, coefs, icpts, C, M, ctx, A, want_f64=True)
    assert np.max(np.abs(Bd - B_ref)) < 1e-11
    assert np.mean(Bf == B_ref.astype(np.float32)) > 0.9999
    assert np.array_equal(np.argmax(Bf, -1), np.argmax(B_ref, -1))
    # numpy twin of the fixed-point logits agrees with the C one
    qf = npo.lr_quantize_fold(coefs, C, M, ctx, s)
    logits = npo.lr_fixed_logits(X, qf, icpts, C, M, ctx, s)
    p = npo.expit(logits)
    p = np.concatenate([1 - p, p], -1) if A == 2 else p / p.sum(-1, keepdims=True)
    assert np.max(np.abs(p - Bd)) < 1e-15


def test_slide_window_against_reference():
    d = _load("slide_window.npz")
    for tag in ("small", "s75", "s9"):
        B, S, Xs = d["B_" + tag], int(d["S_" + tag]), d["X_" + tag]
        assert np.array_equal(npo.slide_window(B, S), Xs)
        assert np.array_equal(co.slide_window(B.astype(np.float32), S), Xs)
    # SURVEY.md 8(c) known answers: W=12, S=5 -> row 0 sees windows [2,1,0,0,1], row 11 [8,9,10,11,11]
    W, S = 12, 5
    B = np.arange(W, dtype=np.float64).reshape(1, W, 1)
    rows = npo.slide_window(B, S)
    assert rows[0].tolist() == [2, 1, 0, 0, 1] and rows[11].tolist() == [8, 9, 10, 11, 11]


def test_covrsk_against_reference():
    d = _load("covrsk.npz")
    assert d["Ms_2500"].tolist() == [1, 4, 8, 39, 42, 117, 376, 866]
    assert npo.cov_sample(2500) == d["Ms_2500"].tolist()
    assert npo.cov_sample(300) == d["Ms_300"].tolist()
    X, Y, K = d["K_X"], d["K_Y"], d["K"]
    Ms = npo.cov_sample(X.shape[1])
    assert np.array_equal(npo.covrsk_kernel(X, Y, Ms), K)
    assert np.array_equal(co.covrsk(X, Y, Ms), K)


def test_svc_proba_against_reference():
    """CovRSKBase -> sklearn SVC(kernel=callable, probability=True) -> libsvm, as driven by the
    reference's Base.predict_proba_vectorized; restated from the fitted attributes."""
    d = _load("covrsk.npz")
    C, M, A, ctx = int(d["svc_C"]), int(d["svc_M"]), int(d["svc_A"]), int(d["svc_ctx"])
    Xq, Xt, B_ref = d["svc_X"], d["svc_Xtrain"], d["svc_B"]
    Xqp, Xtp = npo.base_pad(Xq, ctx), npo.base_pad(Xt, ctx)
    for w, (lo, hi) in enumerate(npo.base_window_ranges(C, M, ctx)):
        sv = Xtp[d["svc_w%d_support" % w], lo:hi]
        Ms = npo.cov_sample(hi - lo)
        K = co.covrsk(Xqp[:, lo:hi], sv, Ms)
        args = (d["svc_w%d_n_support" % w], d["svc_w%d_dual_coef" % w], d["svc_w%d_intercept" % w],
                d["svc_w%d_probA" % w], d["svc_w%d_probB" % w])
        p_c = co.svc_proba(K, *args)
        assert np.max(np.abs(p_c - B_ref[:, w, :])) < 1e-12, "window %d" % w
        p_np = npo.svc_predict_proba(K[:3], *args)
        assert np.max(np.abs(p_np - B_ref[:3, w, :])) < 1e-12


def _golden_forest(d):
    return npo.GBTModel(int(d["A"]), int(d["S"]) * int(d["A"]), d["feat"], d["thr"], d["left"], d["right"],
                        d["default_left"], d["leaf"], d["tree_offsets"], d["base_margin"])


def test_gnofix_against_reference():
    d = _load("gnofix.npz")
    forest = _golden_forest(d)
    S, W, A = int(d["S"]), int(d["W"]), int(d["A"])
    rows_fn = lambda rows: co.gbt_rows(forest, rows)
    smooth_fn = lambda B: co.gbt_smooth(forest, B, S, want_proba=False)[1]
    # the reference's Smoother.predict / predict_proba on the stacked inputs
    Ball = d["B"].reshape(-1, W, A)
    p, y = co.gbt_smooth(forest, Ball, S)
    assert np.array_equal(y, d["Y_smooth"])
    assert np.array_equal(p, d["P_smooth"].astype(np.float32))
    n_sw = 0
    for i in range(len(d["X"])):
        X_m, X_p, Y_m, Y_p, trk = npo.gnofix_default(d["X"][i, 0], d["X"][i, 1], d["B"][i], S, rows_fn, smooth_fn)
        assert np.array_equal(np.array([X_m, X_p]), d["X_out"][i])
        assert np.array_equal(np.array([Y_m, Y_p]), d["Y_out"][i])
        assert np.array_equal(trk, d["tracker"][i])
        n_sw += int((trk[0][:-1] != trk[0][1:]).sum())
    assert n_sw > 10  # the fixture exercises real switches


def test_gbt_traversal_against_sklearn_hgb():
    """Independent pin of the tree predictor: sklearn HistGradientBoosting with the reference's
    hyper-parameters, exported to xgboost form (x <= t64  ->  x < nextafter32(floor32(t64)))."""
    from sklearn.ensemble import HistGradientBoostingClassifier
    from gnomix_b200.gbt import GBTForest
    rng = np.random.default_rng(0)
    A, S = 4, 7
    F = S * A
    Xtr = rng.random((1500, F)).astype(np.float32)
    ytr = (Xtr[:, 3] * 2 + Xtr[:, 10] + 0.3 * rng.normal(size=1500) > 1.4).astype(int) + 2 * (Xtr[:, 20] > 0.5)
    hgb = HistGradientBoostingClassifier(max_iter=25, max_depth=4, learning_rate=0.1, l2_regularization=1.0,
                                         max_leaf_nodes=None, early_stopping=False, random_state=0).fit(Xtr, ytr)
    forest = GBTForest.from_hgb(hgb, F)
    assert forest.n_trees == 25 * A
    Xte = rng.random((400, F)).astype(np.float32)
    Xte[:50] = Xtr[:50]
    p_o = co.gbt_rows(forest, Xte)
    p_ref = hgb.predict_proba(Xte)
    assert np.max(np.abs(p_o - p_ref)) < 2e-6
    assert np.array_equal(np.argmax(p_o, 1), np.argmax(p_ref, 1))
    # numpy twin == C oracle, bit for bit except libm exp (float32 softmax within 1 ulp)
    p_np = npo.gbt_predict_proba(forest, Xte)
    assert np.max(np.abs(p_np - p_o)) < 2e-7


def test_crf_marginals_bruteforce():
    rng = np.random.default_rng(1)
    A = L = 3
    sw, tw = rng.normal(size=(A, L)), rng.normal(size=(L, L))
    B = rng.dirichlet(np.ones(A), size=(2, 6))
    for b in B:
        assert np.max(np.abs(npo.crf_marginals(b, sw, tw) - npo.crf_bruteforce_marginals(b, sw, tw))) < 1e-12
    Bl = rng.dirichlet(np.ones(A), size=(5, 300))
    p_c, l_c = co.crf_smooth(Bl, sw, tw)
    p_np, l_np = npo.crf_smooth(Bl, sw, tw)
    assert np.max(np.abs(p_c - p_np)) < 1e-12
    assert np.array_equal(l_c, l_np)
    assert np.allclose(p_c.sum(-1), 1.0, atol=1e-12)


def test_gnx_math_exp():
    lib = co.lib()
    xs = np.concatenate([np.linspace(-745, 709, 20001), np.random.default_rng(0).normal(0, 5, 20000), [0.0, -0.0, 1e-300]])
    worst = 0.0
    for x in xs:
        got, want = lib.orc_exp(float(x)), math.exp(float(x))
        if want == 0.0 or math.isinf(want):
            continue
        worst = max(worst, abs(got - want) / math.ulp(want))
    assert worst <= 1.0, worst
    assert lib.orc_exp(float("inf")) == float("inf") and lib.orc_exp(-float("inf")) == 0.0
    assert math.isnan(lib.orc_exp(float("nan")))
    xf = np.random.default_rng(1).uniform(-30, 0, 20000).astype(np.float32)
    got = np.array([lib.orc_expf_cr(float(v)) for v in xf], dtype=np.float32)
    want = np.exp(xf.astype(np.float64)).astype(np.float32)
    assert np.array_equal(got, want)


def test_calibrator_restatement_matches_reference_calibrator():
    """oracle calibrator_transform == the reference's Calibrator.transform (scikit-learn isotonic
    models) bit for bit: float32 models (XGB smoother), float64 models (CRF smoother), binary."""
    g = np.load(os.path.join(G, "calibrator.npz"))
    for tag, A in [("a7_f32", 7), ("a3_f64", 3), ("a2_f32", 2)]:
        thr = [(g["x_%s_%d" % (tag, i)], g["y_%s_%d" % (tag, i)]) for i in range(A)]
        got = npo.calibrator_transform(g["q_" + tag], thr)
        assert got.dtype == np.float64 and np.array_equal(got, g["out_" + tag]), tag
    # numpy.interp restatement against numpy itself on awkward inputs
    rng = np.random.default_rng(1)
    xp = np.sort(rng.random(40)); fp = np.sort(rng.random(40))
    x = np.concatenate([rng.random(500) * (xp[-1] - xp[0]) + xp[0], xp, [xp[0], xp[-1]]])
    assert np.array_equal(npo.np_interp_restated(x, xp, fp), np.interp(x, xp, fp))


def test_mode_filter_restatement_and_plugin_against_reference():
    """mode_filter of src/Smooth/utils.py:31-46 (the reference's own function, tests/golden/mode_filter.npz):
    the oracle restatement and the plugin's torch statement (gnomix_b200.smooth.mode_filter_device, run on CPU
    tensors here and on the device in tests/test_pipeline_gpu.py)."""
    import torch
    from gnomix_b200.smooth import mode_filter_device
    d = np.load(os.path.join(G, "mode_filter.npz"))
    for t in "abcde":
        y, size, A = d["y_" + t], int(d["size_" + t]), int(d["A_" + t])
        assert np.array_equal(np.stack([npo.mode_filter(r, size) for r in y]), d["out_" + t]), t
        assert np.array_equal(mode_filter_device(torch.from_numpy(y.astype(np.int32)), size, A).numpy(), d["out_" + t]), t


def test_gnofix_crf_extension_oracle_is_self_consistent():
    """The CRF + Gnofix extension has no reference behaviour (src/model.py:194): its checker is the pinned gnofix control
    flow with the oracle's CRF plugged in.  Here: the NumPy and the C restatement of the CRF give the same phasing, the
    scorer is the centre marginal of the scope run as its own chain, and planted switch errors get undone."""
    from oracle import c_oracle as co, np_oracle as npo
    rng = np.random.default_rng(4)
    W, A, S = 60, 3, 9
    C = W * 7 + 2
    sw = np.eye(A) * 4.0 + rng.normal(0, 0.3, (A, A))
    tw = np.eye(A) * 2.0 + rng.normal(0, 0.3, (A, A))
    anc = np.zeros((2, W), dtype=int)
    anc[0, 25:] = 1
    anc[1, :40] = 2
    b = 0.05 + rng.random((2, W, A)) * 0.2
    for h in range(2):
        b[h, np.arange(W), anc[h]] += 0.7
    b /= b.sum(-1, keepdims=True)
    clean = b.copy()
    b[:, 30:] = b[::-1, 30:].copy()            # one planted switch error at window 30
    X = rng.integers(0, 2, size=(2, C)).astype(np.int8)
    r_np = npo.gnofix_crf_extension(X[0], X[1], b, S, sw, tw)
    r_c = npo.gnofix_crf_extension(X[0], X[1], b, S, sw, tw, crf_smooth_fn=co.crf_smooth)
    for a_, c_ in zip(r_np, r_c):
        assert np.array_equal(a_, c_)
    X_m, X_p, Y_m, Y_p, trk = r_np
    assert (trk[0][1:] != trk[0][:-1]).sum() >= 1                       # it switched
    # the scorer: centre marginal of a scope as its own chain
    rows = b[0, 10:10 + S].reshape(1, -1)
    marg, _ = npo.crf_smooth(rows.reshape(1, S, A), sw, tw)
    assert marg.shape == (1, S, A) and abs(marg[0, (S - 1) // 2].sum() - 1.0) < 1e-12
    # after phasing, each haplotype's labels follow one of the clean tracks again (up to which row is called m)
    clean_lab = npo.crf_smooth(clean, sw, tw)[1]
    got = np.array([Y_m, Y_p])
    same = (got == clean_lab).mean()
    swapped = (got == clean_lab[::-1]).mean()
    assert max(same, swapped) > 0.9
