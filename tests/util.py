"""Shared helpers for the parity tests: seeded synthetic models and inputs.
The oracle (oracle/) is the checker; gnomix_b200 is the thing checked."""
from __future__ import annotations

import numpy as np

from oracle import np_oracle as npo


def geometry(C, M, ctx_ratio=0.5):
    W = C // M
    ctx = int(M * ctx_ratio)
    return W, ctx, npo.base_window_ranges(C, M, ctx)


def random_lr(rng, C, M, A, ctx_ratio=0.5, scale=0.05):
    """Per-window (coef [A_rows, M_w], intercept [A_rows]) float64, sklearn layout."""
    W, ctx, pr = geometry(C, M, ctx_ratio)
    rows = 1 if A == 2 else A
    coefs = [rng.normal(0, scale, size=(rows, hi - lo)) for lo, hi in pr]
    icpts = [rng.normal(0, 0.5, size=rows) for _ in pr]
    return coefs, icpts, ctx


def random_haplotypes(rng, N, C, missing=0.01):
    X = rng.integers(0, 2, size=(N, C), dtype=np.int8)
    if missing > 0:
        X[rng.random((N, C)) < missing] = 2
    return X


def make_lr_base(C, M, A, coefs, icpts, ctx_ratio=0.5):
    from gnomix_b200.base import LogisticRegressionBase, Base
    b = LogisticRegressionBase.__new__(LogisticRegressionBase)
    Base.__init__(b, chm_len=C, window_size=M, num_ancestry=A, context=int(M * ctx_ratio))
    b.base_multithread = True
    b.set_window_weights(coefs, icpts)
    return b


def oracle_lr_fixed(X, coefs, icpts, C, M, ctx, A, limbs=7, want_f64=False):
    from oracle import c_oracle as co
    s = npo.lr_choose_scale(coefs, C, M, ctx, limbs)
    qf = npo.lr_quantize_fold(coefs, C, M, ctx, s)
    return co.lr_fixed(X, qf, np.stack(icpts), C, M, ctx, A, s, want_f64=want_f64), s


def smooth_B(rng, N, W, A):
    """Base-probability-like tensor: rows on the simplex, float32."""
    B = rng.dirichlet(np.full(A, 0.3), size=(N, W)).astype(np.float32)
    return B
