"""gnx_infer_host_ex / Gnomix.predict_host: every plugin combination of run_inference (gnomix.py:48-72) as one
host-buffer pipeline, checked against the stage-by-stage plugin calls (which the other GPU tests check against the
oracle and the reference goldens); packed (2-bit) host input; a reference-style pickle loaded and run."""
import numpy as np
import pytest

from tests import util, refpickle

pytestmark = pytest.mark.gpu


def _lr_xgb_model(rng, C=30_011, M=500, A=7, S=15, rounds=20):
    from gnomix_b200 import Gnomix, GBTForest
    model = Gnomix(C, M, A, S)
    coefs, icpts, _ = util.random_lr(rng, C, M, A)
    model.base.set_window_weights(coefs, icpts)
    model.smooth.model = GBTForest.random(rng, A, model.smooth.S, n_rounds=rounds, depth=4)
    return model, coefs, icpts


def _bits(a):
    return np.ascontiguousarray(a).view({4: np.uint32, 8: np.uint64}[a.dtype.itemsize])


def test_logistic_xgb_int8_and_packed_input():
    from gnomix_b200.io import PackedHaplotypes
    rng = np.random.default_rng(21)
    model, _, _ = _lr_xgb_model(rng)
    X = util.random_haplotypes(rng, 777, model.C)
    y_ref, p_ref = model.predict(X), model.predict_proba(X)
    y, p = model.predict_host(X, want_proba=True, chunk_haps=256)
    assert y.dtype == np.int32 and p.dtype == np.float32
    assert np.array_equal(y, y_ref) and np.array_equal(_bits(p), _bits(p_ref))
    for pinned in (True, False):
        P = PackedHaplotypes.from_numpy(X, pinned=pinned)
        assert np.array_equal(P.to_numpy(), X)
        y2, p2 = model.predict_host(P, want_proba=True, chunk_haps=200)
        assert np.array_equal(y2, y_ref) and np.array_equal(_bits(p2), _bits(p_ref))
    assert np.array_equal(model.predict_host(PackedHaplotypes.from_numpy(X)), y_ref)   # default chunking, labels only


def test_logistic_crf_and_calibrator():
    from gnomix_b200.smooth import CRF_Smoother, CRFModel
    rng = np.random.default_rng(22)
    model, _, _ = _lr_xgb_model(rng)
    C, M, A, S = model.C, model.M, model.A, model.S
    X = util.random_haplotypes(rng, 515, C)
    # tree smoother + calibrator: float64 probabilities, labels from the calibrated ones
    B = model.base.predict_proba(X)
    np.random.seed(1)
    model.smooth.calibrate = True
    model.smooth.train_calibrator(B[:200], np.argmax(B[:200], axis=-1), frac=0.5)
    y, p = model.predict_host(X, want_proba=True, chunk_haps=128)
    assert p.dtype == np.float64
    assert np.array_equal(_bits(p), _bits(model.predict_proba(X))) and np.array_equal(y, model.predict(X))
    assert np.array_equal(model.predict_host(X, chunk_haps=128), model.predict(X))
    # CRF smoother (float64 hand-off), without and with calibrator
    crf = CRF_Smoother(n_windows=C // M, num_ancestry=A, smooth_window_size=S)
    crf.model = CRFModel(rng.normal(0, 1.5, (A, A)), rng.normal(0, 1.0, (A, A)))
    model.smooth = crf
    y, p = model.predict_host(X, want_proba=True, chunk_haps=128)
    assert p.dtype == np.float64
    assert np.array_equal(_bits(p), _bits(model.predict_proba(X))) and np.array_equal(y, model.predict(X))
    np.random.seed(2)
    crf.calibrate = True
    crf.train_calibrator(B[:200], np.argmax(B[:200], axis=-1), frac=0.5)
    y, p = model.predict_host(X, want_proba=True, chunk_haps=128)
    assert np.array_equal(_bits(p), _bits(model.predict_proba(X))) and np.array_equal(y, model.predict(X))


def test_covrsk_xgb():
    from gnomix_b200 import Gnomix, GBTForest
    rng = np.random.default_rng(23)
    C, M, A, S = 4100, 200, 3, 9
    model = Gnomix(C, M, A, S, mode="best")
    b = model.base
    W, P = C // M, A * (A - 1) // 2
    tr = util.random_haplotypes(rng, 18, C, missing=0.0)
    trp = b.pad(tr)
    sl = b.window_slices()
    nsup = np.array([6, 6, 6], np.int32)
    b.set_window_svcs([trp[:, lo:hi] for lo, hi in sl], [nsup] * W, [rng.normal(0, 1e-3, (A - 1, 18)) for _ in range(W)],
                      [rng.normal(0, 0.1, P) for _ in range(W)], [np.full(P, -1.0)] * W, [np.zeros(P)] * W)
    model.smooth.model = GBTForest.random(rng, A, model.smooth.S, n_rounds=10, depth=4)
    X = util.random_haplotypes(rng, 140, C)
    y, p = model.predict_host(X, want_proba=True, chunk_haps=64)
    assert np.array_equal(y, model.predict(X)) and np.array_equal(_bits(p), _bits(model.predict_proba(X)))


def test_phase_pipeline_equals_phase_then_predict():
    rng = np.random.default_rng(24)
    model, _, _ = _lr_xgb_model(rng, C=24_000, M=400, A=4, S=9, rounds=15)
    from gnomix_b200 import synth
    freqs = synth.population_frequencies(rng, model.C, model.A, fst=0.3)
    fx, fpop = synth.founders(rng, freqs, per_pop=6)
    X, _ = synth.admix_host(rng, fx, fpop, 61, morgans=1.0)     # odd: the trailing haplotype is dropped
    Xp_ref, y_ref = model.phase(X)
    p_ref = model.predict_proba(Xp_ref.astype(np.int8))
    y, p, Xp = model.predict_host(X, want_proba=True, phase=True, want_phased=True, chunk_haps=16)
    assert y.shape == (60, model.W) and np.array_equal(y, y_ref)
    assert np.array_equal(Xp, Xp_ref) and np.array_equal(_bits(p), _bits(p_ref))
    assert np.array_equal(model.predict_host(X, phase=True), y_ref)


def test_reference_pickle_runs(tmp_path):
    """A pickle with the reference's object graph (sklearn LogisticRegression + xgboost Booster buffer in the
    1.0-1.5 serialisation envelope) loads without those libraries and predicts what the same weights installed
    directly predict."""
    import gnomix_b200
    from gnomix_b200 import xgb_io
    rng = np.random.default_rng(25)
    model, coefs, icpts = _lr_xgb_model(rng, C=12_345, M=300, A=5, S=9, rounds=12)
    data = refpickle.reference_pickle(model.C, model.M, model.A, model.S, model.context, lr=(coefs, icpts),
                                      booster_bytes=xgb_io.wrap_serialized(xgb_io.write_legacy_binary(model.smooth.model)))
    path = tmp_path / "model_chm_22.pkl"
    path.write_bytes(data)
    m = gnomix_b200.load_model(str(path))
    X = util.random_haplotypes(rng, 100, model.C)
    assert np.array_equal(m.predict(X), model.predict(X))
    assert np.array_equal(_bits(m.predict_proba(X)), _bits(model.predict_proba(X)))
    assert np.array_equal(m.predict_host(X), model.predict(X))


def test_binary_smoother_trained_here_runs():
    """A = 2 through train -> predict (HGB fits one tree per iteration; exported as two classes)."""
    import warnings
    from gnomix_b200 import Gnomix, synth
    from oracle import c_oracle as co
    rng = np.random.default_rng(26)
    C, M, A, S = 4000, 200, 2, 5
    freqs = synth.population_frequencies(rng, C, A, fst=0.3)
    fx, fpop = synth.founders(rng, freqs, per_pop=40)
    W = C // M
    y = np.repeat(fpop[:, None], W, axis=1)
    idx = rng.permutation(len(fx))
    t1, t2 = idx[:40], idx[40:]
    model = Gnomix(C, M, A, S)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model.train(((fx[t1], y[t1]), (fx[t2], y[t2]), (None, None)), retrain_base=False, evaluate=False, verbose=False)
    assert model.smooth.model.A == 2
    B = model.base.predict_proba(fx[t2]).astype(np.float32)
    p_o, l_o = co.gbt_smooth(model.smooth.model, B, model.smooth.S)
    assert np.array_equal(model.smooth.predict(B), l_o) and np.array_equal(_bits(model.smooth.predict_proba(B)), _bits(p_o))
    assert (model.predict(fx[t2]) == y[t2]).mean() > 0.8


def test_mode_filter_on_device_equals_reference_golden():
    import os
    import torch
    from gnomix_b200.smooth import mode_filter_device
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "mode_filter.npz"))
    for t in "abcde":
        got = mode_filter_device(torch.from_numpy(d["y_" + t].astype(np.int32)).cuda(), int(d["size_" + t]), int(d["A_" + t]))
        assert np.array_equal(got.cpu().numpy(), d["out_" + t]), t
