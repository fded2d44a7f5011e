#!/usr/bin/env python
"""Run this INSIDE the environment the reference model was trained in (AI-sandbox/gnomix checkout on
PYTHONPATH, scikit-learn and xgboost installed):

    python export_reference_model.py path/to/model_chm_22.pkl[.gz] exported_chm_22.npz

It writes everything the GPU engine needs into one .npz (numpy + json only; no dependency on
gnomix_b200): geometry, SNP annotation, the per-window logistic weights in the padded-window feature
order they were fitted on, the xgboost smoother as xgboost's own JSON model, and the genetic map.
`gnomix_b200.convert.load_exported_model` reads it."""
import gzip
import os
import pickle
import sys
import tempfile

import numpy as np


def main(pkl, out):
    opener = gzip.open if pkl.endswith(".gz") else open
    with opener(pkl, "rb") as f:
        m = pickle.load(f)
    assert type(m.base).__name__ == "LogisticRegressionBase" and type(m.smooth).__name__ == "XGB_Smoother", \
        "only the default logistic + XGB model is exported (got %s + %s)" % (type(m.base).__name__, type(m.smooth).__name__)
    coefs, icpts = [], []
    for mdl in m.base.models:
        coefs.append(np.asarray(mdl.coef_, dtype=np.float64).ravel())
        icpts.append(np.atleast_1d(np.asarray(mdl.intercept_, dtype=np.float64)))
    booster = m.smooth.model.get_booster()
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "smoother.json")
        booster.save_model(p)
        xgb_json = open(p).read()
    g = m.gen_map_df
    extra = {}
    if g is not None and len(g):
        extra = dict(gen_map_chm=np.asarray(g["chm"]).astype(str), gen_map_pos=np.asarray(g["pos"]).astype(np.int64),
                     gen_map_cm=np.asarray(g["pos_cm"]).astype(np.float64))
    np.savez_compressed(out, C=m.C, M=m.M, A=m.A, S=m.smooth.S, context=m.context, context_ratio=m.context / m.M,
                        snp_pos=np.asarray(m.snp_pos), snp_ref=np.asarray(m.snp_ref).astype(str), snp_alt=np.asarray(m.snp_alt).astype(str),
                        population_order=np.asarray(m.population_order).astype(str), lr_coef=np.concatenate(coefs),
                        lr_intercept=np.stack(icpts), xgb_json=np.array(xgb_json), **extra)
    print("wrote", out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
