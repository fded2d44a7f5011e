"""Trains the smoother forest used by bench.py (gnomix_b200/data/forest_<workload>.npz).

Same founders and per-window logistic weights as bench.build_models (seeded), admixed training
haplotypes from those founders, base probabilities from the float64 CPU restatement, window
labels = per-window mode of the SNP-level ancestry (reference src/preprocess.py:37-59), then the
reference's smoother hyper-parameters (100 rounds x A trees, depth 4, lr 0.1, lambda 1 --
src/Smooth/models.py:14-20) through scikit-learn's HistGradientBoosting (xgboost is not
installable offline), exported to xgboost form.  Runs on CPU in a few minutes:

    python scripts/make_bench_forest.py chr1
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench
from gnomix_b200 import synth
from gnomix_b200.gbt import GBTForest
from gnomix_b200.smooth import host_slide_window
from oracle import np_oracle as npo


def main(workload="chr1", n_train=160, max_rows=150_000):
    bench.WORKLOAD = workload
    geom = synth.GEOMETRY[workload]
    C, M, A, S, morgans = geom
    W = C // M
    fixture = os.path.join(bench.ROOT, "gnomix_b200", "data", "forest_%s.npz" % workload)
    if os.path.exists(fixture):
        os.remove(fixture)
    base, smooth, (fx, fpop), (coefs, icpts, ctx), _ = bench.build_models(geom)
    rng = np.random.default_rng(bench.SEED + 99)
    t = time.time()
    X, y_snp = synth.admix_host(rng, fx, fpop, n_train, morgans)
    B = npo.lr_base_predict_proba(X, coefs, icpts, C, M, ctx).astype(np.float32)
    # window label = mode over the window's SNPs; last window absorbs the remainder
    yw = np.empty((n_train, W), dtype=np.int64)
    for w in range(W):
        lo, hi = w * M, (C if w == W - 1 else (w + 1) * M)
        seg = y_snp[:, lo:hi]
        yw[:, w] = np.array([np.bincount(r, minlength=A).argmax() for r in seg])
    print("base accuracy (argmax B vs truth): %.3f  [%.0fs]" % ((np.argmax(B, -1) == yw).mean(), time.time() - t))
    Xs = host_slide_window(B, S)
    ys = yw.reshape(-1)
    if len(Xs) > max_rows:
        idx = rng.choice(len(Xs), max_rows, replace=False)
        Xs, ys = Xs[idx], ys[idx]
    from sklearn.ensemble import HistGradientBoostingClassifier
    t = time.time()
    hgb = HistGradientBoostingClassifier(max_iter=100, max_depth=4, learning_rate=0.1, l2_regularization=1.0, max_leaf_nodes=None,
                                         early_stopping=False, random_state=bench.SEED).fit(Xs, ys)
    forest = GBTForest.from_hgb(hgb, S * A)
    print("trained %d trees in %.0fs; train accuracy %.4f" % (forest.n_trees, time.time() - t, (hgb.predict(Xs) == ys).mean()))
    np.savez_compressed(fixture, **forest.to_npz_dict())
    print("wrote", fixture, os.path.getsize(fixture), "bytes")


if __name__ == "__main__":
    main(*(sys.argv[1:2] or ["chr1"]))
