"""Bringing a model trained with the reference (scikit-learn 1.0.1 + xgboost 1.1.1) over to this
package without unpickling it here.

xgboost cannot be imported in this environment, so the reference's `.pkl` (which embeds an
`xgboost.sklearn.XGBClassifier`) is exported ONCE, inside the environment it was trained in, by
`scripts/export_reference_model.py` (plain numpy + json, no dependency on this package) into a
single `.npz`; `load_exported_model` turns that file into a `gnomix_b200.Gnomix` whose stages run
on the GPU.  The smoother travels as xgboost's own JSON model (`Booster.save_model("*.json")`,
available since xgboost 1.0), parsed by `forest_from_xgboost_json`."""
from __future__ import annotations

import json

import numpy as np

from .gbt import GBTForest


def forest_from_xgboost_json(obj, num_class=None, n_features=None) -> GBTForest:
    """xgboost JSON model (dict, JSON string or path) -> GBTForest (schema: xgb_io.forest_from_model_dict)."""
    from .xgb_io import forest_from_model_dict
    if isinstance(obj, (str, bytes)):
        txt = obj if isinstance(obj, str) else obj.decode()
        obj = json.loads(txt) if txt.lstrip().startswith("{") else json.load(open(txt))
    return forest_from_model_dict(obj, num_class, n_features)


def forest_to_xgboost_json(forest: GBTForest) -> dict:
    """The inverse (used by the tests and to hand a forest trained here to xgboost users)."""
    trees = []
    for t in range(forest.n_trees):
        o, e = int(forest.tree_offsets[t]), int(forest.tree_offsets[t + 1])
        is_leaf = forest.feat[o:e] < 0
        trees.append({
            "id": t,
            "left_children": [-1 if l else int(v) for l, v in zip(is_leaf, forest.left[o:e])],
            "right_children": [-1 if l else int(v) for l, v in zip(is_leaf, forest.right[o:e])],
            "split_indices": [0 if l else int(v) for l, v in zip(is_leaf, forest.feat[o:e])],
            "split_conditions": [float(lv) if l else float(tv) for l, lv, tv in zip(is_leaf, forest.leaf[o:e], forest.thr[o:e])],
            "default_left": [int(v) for v in forest.default_left[o:e]],
            "tree_param": {"num_nodes": str(e - o), "num_feature": str(forest.n_features)},
        })
    return {"learner": {"learner_model_param": {"base_score": repr(float(forest.base_margin[0])), "num_class": str(forest.A),
                                                "num_feature": str(forest.n_features)},
                        "gradient_booster": {"name": "gbtree", "model": {"gbtree_model_param": {"num_trees": str(forest.n_trees)},
                                                                       "tree_info": [t % forest.A for t in range(forest.n_trees)],
                                                                       "trees": trees}},
                        "objective": {"name": "multi:softprob"}},
            "version": [1, 1, 1]}


def load_exported_model(path):
    """`.npz` written by scripts/export_reference_model.py -> gnomix_b200.Gnomix (logistic base + XGB
    smoother), ready for predict / predict_proba / phase."""
    import pandas as pd
    from .model import Gnomix
    d = np.load(path, allow_pickle=False)
    C, M, A, S = int(d["C"]), int(d["M"]), int(d["A"]), int(d["S"])
    model = Gnomix(C, M, A, S, snp_pos=d["snp_pos"], snp_ref=d["snp_ref"], snp_alt=d["snp_alt"],
                   population_order=[str(p) for p in d["population_order"]], context_ratio=float(d["context_ratio"]))
    model.context = model.base.context = int(d["context"])   # the fitted value, not int(M * ratio) re-derived
    rows = 1 if A == 2 else A
    coefs, o = [], 0
    for lo, hi in model.base.window_slices():
        n = rows * (hi - lo)
        coefs.append(d["lr_coef"][o:o + n].reshape(rows, hi - lo))
        o += n
    assert o == len(d["lr_coef"]), "logistic weights do not match the window geometry"
    model.base.set_window_weights(coefs, list(d["lr_intercept"]))
    model.smooth.model = forest_from_xgboost_json(str(d["xgb_json"]), num_class=A, n_features=model.smooth.S * A)
    if "gen_map_pos" in d.files:
        model.write_gen_map_df(pd.DataFrame({"chm": [str(c) for c in d["gen_map_chm"]], "pos": d["gen_map_pos"], "pos_cm": d["gen_map_cm"]}))
    return model
