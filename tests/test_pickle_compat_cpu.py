"""Reference `.pkl` ingestion without scikit-learn / xgboost / sklearn-crfsuite state being imported
(gnomix_b200/pickle_compat.py, gnomix_b200/xgb_io.py; reference gnomix.py:26-35).  CPU only: byte-level
round trips of every booster buffer format and of the whole reference object graph."""
import gzip
import json
import pickle
import struct

import numpy as np
import pytest

from gnomix_b200 import GBTForest, Gnomix, XGB_Smoother, CRF_Smoother, LogisticRegressionBase, CovRSKBase
from gnomix_b200 import pickle_compat as pc, xgb_io, convert
from gnomix_b200.smooth import CRFModel
from tests import refpickle


def _forest(seed=0, A=3, S=5, rounds=4, ragged=True):
    rng = np.random.default_rng(seed)
    f = GBTForest.random(rng, A, S, n_rounds=rounds, depth=3)
    if not ragged:
        return f
    # make it look like a real booster: some trees shallower (a leaf where a split was), default directions mixed
    feat, thr, left, right, dl, leaf, offs = [], [], [], [], [], [], [0]
    for t in range(f.n_trees):
        o, e = f.tree_offsets[t], f.tree_offsets[t + 1]
        if t % 3 == 1:      # a stump: root + two leaves
            feat += [int(f.feat[o]), -1, -1]; thr += [float(f.thr[o]), 0, 0]; left += [1, 0, 0]; right += [2, 0, 0]
            dl += [int(f.default_left[o]), 0, 0]; leaf += [0, float(rng.normal()), float(rng.normal())]
        elif t % 3 == 2:    # a single leaf
            feat += [-1]; thr += [0]; left += [0]; right += [0]; dl += [0]; leaf += [float(rng.normal())]
        else:
            feat += list(f.feat[o:e]); thr += list(f.thr[o:e]); left += list(f.left[o:e]); right += list(f.right[o:e])
            dl += list(f.default_left[o:e]); leaf += list(f.leaf[o:e])
        offs.append(len(feat))
    return GBTForest(A, S * A, feat, thr, left, right, dl, leaf, offs, np.full(A, 0.5, np.float32))


def _same_forest(a, b):
    assert a.A == b.A and a.n_features == b.n_features and a.n_trees == b.n_trees
    for k in ("feat", "left", "right", "default_left", "tree_offsets"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    for k in ("thr", "leaf", "base_margin"):
        assert np.array_equal(getattr(a, k).view(np.uint32), getattr(b, k).view(np.uint32)), k


@pytest.mark.parametrize("binf", [True, False])
def test_legacy_binary_round_trip(binf):
    f = _forest()
    raw = xgb_io.write_legacy_binary(f, binf=binf, attributes=[("objective", '{"name":"multi:softprob"}')])
    # layout facts of the format (src/learner.cc, gbtree_model.h, tree_model.h)
    off = 4 if binf else 0
    assert struct.unpack_from("<f", raw, off)[0] == 0.5 and struct.unpack_from("<i", raw, off + 8)[0] == f.A
    p = off + xgb_io.LEARNER_PARAM_BYTES
    assert raw[p:p + 8] == struct.pack("<Q", 14) and raw[p + 8:p + 22] == b"multi:softprob"
    _same_forest(xgb_io.parse_legacy_binary(raw), f)
    _same_forest(xgb_io.forest_from_booster_bytes(raw), f)
    _same_forest(xgb_io.forest_from_booster_bytes(xgb_io.wrap_serialized(raw)), f)   # what a pickled 1.0-1.5 Booster holds


def test_legacy_binary_rejects_garbage():
    f = _forest()
    raw = bytearray(xgb_io.write_legacy_binary(f))
    with pytest.raises(ValueError):
        xgb_io.parse_legacy_binary(raw[:200])
    bad = bytearray(raw)
    struct.pack_into("<i", bad, 4 + 8, 1)            # num_class = 1
    with pytest.raises(ValueError):
        xgb_io.parse_legacy_binary(bytes(bad))
    bad = bytearray(raw)
    bad[-4 * f.n_trees:] = struct.pack("<%di" % f.n_trees, *([0] * f.n_trees))   # tree_info all class 0
    with pytest.raises(ValueError):
        xgb_io.parse_legacy_binary(bytes(bad))


def test_ubjson_round_trip_and_model():
    obj = {"a": [1, 2, 300, -70000], "b": [0.5, 0.25], "c": {"d": "text", "e": None, "f": True, "g": False, "h": 2 ** 40, "i": 0.1},
           "mixed": [1, "x", 2.5, [1, 2]], "empty": [], "emptyd": {}}
    for typed in (True, False):
        got = xgb_io.ubjson_loads(xgb_io.ubjson_dumps(obj, typed_arrays=typed))
        assert got == obj
    # hand-assembled optimised containers: [$U#i3 1 2 3], {#i1 i1 k i5}
    assert xgb_io.ubjson_loads(b"[$U#i\x03\x01\x02\x03") == [1, 2, 3]
    assert xgb_io.ubjson_loads(b"{#i\x01i\x01ki\x05") == {"k": 5}
    assert xgb_io.ubjson_loads(b"[$d#i\x02" + struct.pack(">ff", 0.5, -2.0)) == [0.5, -2.0]
    f = _forest(1)
    model = convert.forest_to_xgboost_json(f)
    _same_forest(xgb_io.forest_from_booster_bytes(json.dumps(model).encode()), f)
    _same_forest(xgb_io.forest_from_booster_bytes(xgb_io.ubjson_dumps(model)), f)
    snap = {"Model": model, "Config": {"learner": {}}}     # pickled Booster of xgboost >= 1.6
    _same_forest(xgb_io.forest_from_booster_bytes(xgb_io.ubjson_dumps(snap)), f)
    _same_forest(xgb_io.forest_from_booster_bytes(json.dumps(snap).encode()), f)
    # xgboost >= 2 writes base_score as a string in scientific notation, 3.x as a bracketed vector
    model["learner"]["learner_model_param"]["base_score"] = "[5E-1]"
    _same_forest(xgb_io.forest_from_model_dict(model), f)


def _lr_weights(rng, C, M, A, ctx):
    W = C // M
    rem = C - M * W
    rows = 1 if A == 2 else A
    lens = [M + 2 * ctx] * (W - 1) + [M + 2 * ctx + rem]
    return [rng.normal(0, 0.1, (rows, n)) for n in lens], [rng.normal(0, 0.1, rows) for _ in lens]


def test_reference_pickle_logistic_xgb(tmp_path):
    rng = np.random.default_rng(5)
    C, M, A, S, ctx = 1000, 90, 3, 5, 45
    coefs, icpts = _lr_weights(rng, C, M, A, ctx)
    f = _forest(2, A=A, S=S)
    cal = [(np.sort(rng.random(6)).astype(np.float32), np.sort(rng.random(6)).astype(np.float32)) for _ in range(A)]
    for booster in (xgb_io.wrap_serialized(xgb_io.write_legacy_binary(f)), xgb_io.write_legacy_binary(f, binf=False),
                    xgb_io.ubjson_dumps({"Model": convert.forest_to_xgboost_json(f), "Config": {}})):
        data = refpickle.reference_pickle(C, M, A, S, ctx, lr=(coefs, icpts), booster_bytes=booster, calibrator=cal, mode_filter=5)
        assert b"xgboost.core" in data and b"sklearn.linear_model._logistic" in data and b"src.model" in data
        m = pc.loads(data)
        assert type(m) is Gnomix and type(m.base) is LogisticRegressionBase and type(m.smooth) is XGB_Smoother
        assert (m.C, m.M, m.A, m.S, m.W, m.context) == (C, M, A, S, C // M, ctx)
        c, b = m.base.packed_weights()
        assert np.array_equal(c, np.concatenate([x.ravel() for x in coefs])) and np.array_equal(b, np.stack(icpts))
        assert isinstance(m.smooth.model, GBTForest)
        _same_forest(m.smooth.model, f)
        assert m.smooth.mode_filter == 5 and m.smooth.gnofix and m.smooth.S == S
        thr = m.smooth.calibrator.thresholds()
        assert all(np.array_equal(t[0], c0[0]) and np.array_equal(t[1], c0[1]) for t, c0 in zip(thr, cal))
        assert list(m.population_order) == ["P0", "P1", "P2"] and len(m.snp_pos) == C
    # through a file, gzip as gnomix.py:29-31 accepts, and re-pickled by this package
    p = tmp_path / "model_chm_22.pkl.gz"
    with gzip.open(p, "wb") as fh:
        fh.write(data)
    import gnomix_b200
    m2 = gnomix_b200.load_model(str(p))
    _same_forest(m2.smooth.model, f)
    m3 = pc.loads(pickle.dumps(m2))
    _same_forest(m3.smooth.model, f)
    from gnomix_b200.cli import load_model as cli_load
    _same_forest(cli_load(str(p), verbose=False).smooth.model, f)


def test_reference_pickle_covrsk_and_crf():
    rng = np.random.default_rng(6)
    C, M, A, S, ctx = 600, 100, 3, 5, 50
    W = C // M
    P = A * (A - 1) // 2
    lens = [M + 2 * ctx] * W
    n = 12
    Xfit = [rng.integers(0, 2, (n, L)).astype(np.int8) for L in lens]
    support = [np.sort(rng.choice(n, 9, replace=False)) for _ in range(W)]
    n_support = [np.array([3, 3, 3]) for _ in range(W)]
    dual = [rng.normal(0, 1, (A - 1, 9)) for _ in range(W)]
    icpt, pA, pB = ([rng.normal(0, 1, P) for _ in range(W)] for _ in range(3))
    sw, tw = rng.normal(0, 1, (A, A)), rng.normal(0, 1, (A, A))
    # CRFsuite numbers labels / attributes in order of first appearance: permute them to exercise the mapping
    lab_order, att_order = [2, 0, 1], [1, 2, 0]
    crf_bytes = pc.write_crfsuite_model([str(i) for i in lab_order], [str(i) for i in att_order],
                                        sw[np.ix_(att_order, lab_order)], tw[np.ix_(lab_order, lab_order)])
    labels, attrs, sw_f, tw_f = pc.parse_crfsuite_model(crf_bytes)
    assert labels == ["2", "0", "1"] and attrs == ["1", "2", "0"]
    data = refpickle.reference_pickle(C, M, A, S, ctx, svc=(Xfit, support, n_support, dual, icpt, pA, pB), crfsuite_bytes=crf_bytes)
    assert b"sklearn_crfsuite" in data and b"sklearn.svm._classes" in data
    m = pc.loads(data)
    assert type(m.base) is CovRSKBase and type(m.smooth) is CRF_Smoother
    assert all(np.array_equal(m.base.sv_rows[w], Xfit[w][support[w]]) for w in range(W))
    f = m.base._fitted
    assert all(np.array_equal(f["dual_coef"][w], dual[w]) and np.array_equal(f["intercept"][w], icpt[w]) and
               np.array_equal(f["probA"][w], pA[w]) and np.array_equal(f["probB"][w], pB[w]) and
               np.array_equal(f["n_support"][w], n_support[w]) for w in range(W))
    assert isinstance(m.smooth.model, CRFModel)
    assert np.array_equal(m.smooth.model.state_w, sw) and np.array_equal(m.smooth.model.trans_w, tw)


def test_unsupported_plugins_fail_loudly():
    K = refpickle._make_classes()
    other = refpickle._cls("src.Base.models", "XGBBase")
    K["XGBBase"] = other
    model = refpickle._obj(K["Gnomix"], C=100, M=10, A=2, S=5, base=refpickle._obj(other, models=[]), smooth=None)
    with refpickle._registered(K):
        data = pickle.dumps(model)
    with pytest.raises(NotImplementedError, match="XGBBase"):
        pc.loads(data)
    with pytest.raises(TypeError):
        pc.loads(pickle.dumps({"not": "a model"}))


def test_binary_hgb_export_has_two_classes():
    """a 2-ancestry smoother trained here must give an A=2 forest (HGB fits one tree per iteration)."""
    from sklearn.ensemble import HistGradientBoostingClassifier
    from oracle import c_oracle as co
    rng = np.random.default_rng(0)
    X = rng.random((1500, 10)).astype(np.float32)
    y = (X[:, 0] + X[:, 3] > 1).astype(int)
    h = HistGradientBoostingClassifier(max_iter=10, max_depth=4, early_stopping=False).fit(X, y)
    f = GBTForest.from_hgb(h, 10)
    assert f.A == 2 and f.n_trees == 20 and f.n_features == 10
    assert np.abs(co.gbt_rows(f, X) - h.predict_proba(X)).max() < 2e-6
