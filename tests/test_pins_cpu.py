"""Pins of the two restatements that cannot be pinned offline: xgboost's predictor and CRFsuite's
marginals.  The golden files are written by scripts/make_pin_goldens.py in an environment that has xgboost /
sklearn-crfsuite (neither is installable here); until they are committed these tests are SKIPPED and the parity of
rows E / F of SURVEY.md section 8 stays "unpinned"."""
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(__file__), "golden")


def _need(name):
    p = os.path.join(G, name)
    if not os.path.exists(p):
        pytest.skip("%s not present: run scripts/make_pin_goldens.py where xgboost / sklearn-crfsuite are installed" % name)
    return np.load(p, allow_pickle=False)


def test_xgboost_predictor_pin():
    """cpu_predictor.cc PredValue (float32 leaf sums in tree order from zero, base margin added last, `fvalue < split_cond`,
    default direction for NaN) + common/math.h Softmax, restated in oracle/np_oracle.py gbt_margins / softmax_xgb and
    oracle/gnx_oracle.c: float32 probabilities bit for bit."""
    from gnomix_b200 import convert, xgb_io
    from oracle import c_oracle as co, np_oracle as npo
    d = _need("pin_xgboost.npz")
    A, S = int(d["A"]), int(d["S"])
    f = convert.forest_from_xgboost_json(str(d["model_json"]), num_class=A, n_features=A * S)
    rows = d["rows"].astype(np.float32)
    want = d["proba_bits"]
    got = co.gbt_rows(f, rows)
    assert np.array_equal(got.view(np.uint32), want), "C oracle differs from xgboost %s" % d["xgboost_version"]
    model = npo.GBTModel(f.A, f.n_features, f.feat, f.thr, f.left, f.right, f.default_left, f.leaf, f.tree_offsets, f.base_margin)
    assert np.array_equal(npo.gbt_predict_proba(model, rows[:64]).view(np.uint32), want[:64])
    # every buffer format the package parses gives the same forest as the JSON model
    for key in ("save_raw", "pickled_handle"):
        buf = d[key].tobytes()
        if not buf:
            continue
        g = xgb_io.forest_from_booster_bytes(buf, num_class=A, n_features=A * S)
        assert np.array_equal(co.gbt_rows(g, rows).view(np.uint32), want), key
    from gnomix_b200 import pickle_compat as pc
    import io
    m = pc.ReferenceUnpickler(io.BytesIO(d["pickled_model"].tobytes())).load()
    g = pc.adopt_smoother_model(m, A, S)
    assert np.array_equal(co.gbt_rows(g, rows).view(np.uint32), want)


def test_crfsuite_marginals_pin():
    """crf1d_context.c crf1dc_alpha_score / crf1dc_beta_score / crf1dc_marginal_point restated in oracle/np_oracle.py
    crf_marginals (scaled forward-backward, float64)."""
    from gnomix_b200 import pickle_compat as pc
    from oracle import np_oracle as npo
    d = _need("pin_crfsuite.npz")
    A = int(d["A"])
    sw, tw = d["state_w"], d["trans_w"]
    got = np.stack([npo.crf_marginals(x, sw, tw) for x in d["X"]])
    assert np.max(np.abs(got - d["marginals"])) <= 1e-12
    labels, attrs, sw2, tw2 = pc.parse_crfsuite_model(d["model_file"].tobytes())
    lo, ao = np.argsort([int(s) for s in labels]), np.argsort([int(s) for s in attrs])
    assert np.array_equal(sw2[np.ix_(ao, lo)], sw) and np.array_equal(tw2[np.ix_(lo, lo)], tw)
    import io
    m = pc.ReferenceUnpickler(io.BytesIO(d["pickled_model"].tobytes())).load()
    cm = pc.crf_from_foreign(m, A)
    assert np.array_equal(cm.state_w, sw) and np.array_equal(cm.trans_w, tw)
