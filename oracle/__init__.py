"""CPU oracle for the gnomix hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  Nothing under ``gnomix_b200/``
imports it, and the product path raises if the CUDA library is missing.

Pinning status (see DESIGN.md "Oracle"):
  * windowing, reflect pad, slide_window, CovRSK kernel, CovSample, gnofix control
    flow, get_meta_data: pinned against the reference's own Python, imported from
    /root/reference with stub modules (oracle/refimport.py) -> tests/golden/*.npz
    (generator: oracle/make_golden.py).
  * LR base arithmetic: pinned against scikit-learn 1.9 OneVsRest(LogisticRegression
    (liblinear)) run through the reference's Base.predict_proba_vectorized.
  * SVC probability: pinned against sklearn.svm.SVC (libsvm) with the reference's
    callable kernel.
  * GBT predictor (xgboost semantics): xgboost is not installable here -> the
    traversal is pinned against sklearn HistGradientBoosting (independent trees
    implementation); accumulation order / softmax follow xgboost's published
    predictor and are otherwise PARITY UNPINNED.
  * CRF marginals (CRFsuite semantics): sklearn_crfsuite not installable -> pinned
    only against brute-force path enumeration; PARITY UNPINNED w.r.t. CRFsuite.
"""
