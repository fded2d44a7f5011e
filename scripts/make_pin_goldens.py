#!/usr/bin/env python
"""Run this where the reference's third-party stack is installed (requirements.txt: xgboost==1.1.1,
sklearn-crfsuite==0.3.6; any later xgboost works too and is recorded) to produce the two golden files that pin the
oracle's restatement of their arithmetic -- the pins that cannot be produced offline:

    python scripts/make_pin_goldens.py tests/golden/

writes
  pin_xgboost.npz    a small multi:softprob booster trained with the reference's hyper-parameters
                     (src/Smooth/models.py:14-20) on seeded data: the rows, XGBClassifier.predict_proba(rows) as
                     float32 bits, the model as JSON (save_model), as the legacy binary buffer (save_raw) and as the
                     pickled Booster state (what a reference .pkl holds), rows containing NaN included;
  pin_crfsuite.npz   a small CRF trained as src/Smooth/crf.py:5-67 does: the sequences, predict_marginals as float64,
                     state_features_ / transition_features_ and the CRFsuite model file bytes.
tests/test_pins_cpu.py consumes them when present (skipped otherwise) and then demands bit-exact float32
probabilities from the oracle's tree predictor and <= 1e-12 from its CRF marginals, plus identical forests from
every buffer parser of gnomix_b200/xgb_io.py.  Only numpy + the two libraries are needed; nothing of this repository
is imported."""
import json
import os
import pickle
import sys
import tempfile

import numpy as np


def pin_xgboost(out_dir):
    import xgboost
    from xgboost import XGBClassifier
    rng = np.random.default_rng(94305)
    A, S, n = 5, 7, 6000
    F = A * S
    centers = rng.dirichlet(np.full(A, 0.4), size=(A, S))                  # class c: its own mean simplex point per slot
    y = rng.integers(0, A, n)
    X = np.stack([rng.dirichlet(centers[c].mean(0) * 6 + 0.3, size=S).reshape(-1) for c in y]).astype(np.float32)
    kw = dict(n_estimators=30, max_depth=4, learning_rate=0.1, reg_lambda=1, reg_alpha=0, nthread=1, random_state=94305,
              objective="multi:softprob")
    try:
        m = XGBClassifier(num_class=A, use_label_encoder=False, eval_metric="mlogloss", **kw)
        m.fit(X, y)
    except TypeError:
        m = XGBClassifier(**kw)
        m.fit(X, y)
    rows = np.concatenate([X[:400], rng.random((200, F)).astype(np.float32)])
    rows[-50:, ::3] = np.nan                                                 # default-direction semantics
    thr_hits = rows[:100].copy()                                             # exact threshold hits: x == split condition
    booster = m.get_booster()
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "m.json")
        booster.save_model(p)
        model_json = open(p).read()
    conds = sorted({float(c) for t in json.loads(model_json)["learner"]["gradient_booster"]["model"]["trees"]
                    for c, l in zip(t["split_conditions"], t["left_children"]) if l != -1})
    if conds:
        thr_hits[:, :] = np.asarray(conds, dtype=np.float32)[rng.integers(0, len(conds), thr_hits.shape)]
    rows = np.concatenate([rows, thr_hits])
    proba = np.asarray(m.predict_proba(rows), dtype=np.float32)
    try:
        raw = bytes(booster.save_raw("deprecated"))     # xgboost >= 1.6: ask for the legacy binary explicitly
    except TypeError:
        raw = bytes(booster.save_raw())
    state = booster.__getstate__()
    handle = bytes(state["handle"]) if state.get("handle") is not None else b""
    np.savez_compressed(os.path.join(out_dir, "pin_xgboost.npz"), A=A, S=S, rows=rows, proba_bits=proba.view(np.uint32),
                        model_json=np.array(model_json), save_raw=np.frombuffer(raw, dtype=np.uint8),
                        pickled_handle=np.frombuffer(handle, dtype=np.uint8), pickled_model=np.frombuffer(pickle.dumps(m), dtype=np.uint8),
                        xgboost_version=np.array(xgboost.__version__))
    print("pin_xgboost.npz: xgboost", xgboost.__version__, rows.shape, "rows")


def pin_crfsuite(out_dir):
    import sklearn_crfsuite
    rng = np.random.default_rng(94305)
    A, W, n = 4, 30, 300
    Y = np.zeros((n, W), dtype=int)
    for i in range(n):
        y = rng.integers(0, A)
        for w in range(W):
            if rng.random() < 0.1:
                y = rng.integers(0, A)
            Y[i, w] = y
    X = rng.dirichlet(np.full(A, 0.5), size=(n, W))
    X = 0.5 * X + 0.5 * np.eye(A)[Y]
    to_crf = lambda Xs: [[{str(a): Xs[i, b, a] for a in range(A)} for b in range(Xs.shape[1])] for i in range(len(Xs))]   # crf.py:17-33
    crf = sklearn_crfsuite.CRF(algorithm="lbfgs", max_iterations=200, all_possible_transitions=True, all_possible_states=True)
    crf.fit(to_crf(X), [[str(v) for v in row] for row in Y])
    Xq = rng.dirichlet(np.full(A, 0.5), size=(40, W))
    marg = crf.predict_marginals(to_crf(Xq))
    M = np.array([[[marg[i][b][str(a)] for a in range(A)] for b in range(W)] for i in range(len(Xq))], dtype=np.float64)
    sw, tw = np.zeros((A, A)), np.zeros((A, A))
    for (attr, lab), v in crf.state_features_.items():
        sw[int(attr), int(lab)] = v
    for (i, j), v in crf.transition_features_.items():
        tw[int(i), int(j)] = v
    model_bytes = open(crf.modelfile.name, "rb").read()
    np.savez_compressed(os.path.join(out_dir, "pin_crfsuite.npz"), A=A, X=Xq, marginals=M, state_w=sw, trans_w=tw,
                        model_file=np.frombuffer(model_bytes, dtype=np.uint8), pickled_model=np.frombuffer(pickle.dumps(crf), dtype=np.uint8))
    print("pin_crfsuite.npz:", Xq.shape)


if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else "tests/golden"
    os.makedirs(out, exist_ok=True)
    for fn in (pin_xgboost, pin_crfsuite):
        try:
            fn(out)
        except ImportError as e:
            print("skipped", fn.__name__, "-", e)
