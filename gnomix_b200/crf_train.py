"""Host-side training of the linear-chain CRF smoother (not the accelerated path).

The reference trains through sklearn_crfsuite.CRF(algorithm="lbfgs", max_iterations=10000,
all_possible_transitions=True, all_possible_states=True) (src/Smooth/crf.py:7-15), i.e.
CRFsuite's L-BFGS on the L2-regularised (c2 = 1.0, c1 = 0: CRFsuite defaults) conditional
log-likelihood with one state feature per (attribute, label) pair -- attribute values are
the base probabilities -- and one transition feature per label pair.  sklearn_crfsuite is
not installable offline, so the same objective is minimised here with SciPy's L-BFGS.
"""
from __future__ import annotations

import numpy as np


def _logsumexp(a, axis):
    m = np.max(a, axis=axis, keepdims=True)
    return np.squeeze(m, axis) + np.log(np.sum(np.exp(a - m), axis=axis))


def crf_nll_grad(theta, B, y, A, L, c2):
    """B [N, W, A] float64, y [N, W] int.  Returns (objective, gradient)."""
    N, W, _ = B.shape
    sw = theta[:A * L].reshape(A, L)
    tw = theta[A * L:].reshape(L, L)
    state = B @ sw                                                     # [N, W, L]
    la = np.empty((N, W, L))
    lb = np.empty((N, W, L))
    la[:, 0] = state[:, 0]
    for t in range(1, W):
        la[:, t] = _logsumexp(la[:, t - 1][:, :, None] + tw[None], axis=1) + state[:, t]
    lb[:, W - 1] = 0.0
    for t in range(W - 2, -1, -1):
        lb[:, t] = _logsumexp(tw[None] + (state[:, t + 1] + lb[:, t + 1])[:, None, :], axis=2)
    logZ = _logsumexp(la[:, W - 1], axis=1)                            # [N]
    idx_n = np.arange(N)[:, None]
    idx_t = np.arange(W)[None, :]
    gold = state[idx_n, idx_t, y].sum() + tw[y[:, :-1], y[:, 1:]].sum()
    nll = logZ.sum() - gold
    marg = np.exp(la + lb - logZ[:, None, None])                       # [N, W, L]
    onehot = np.zeros_like(marg)
    onehot[idx_n, idx_t, y] = 1.0
    g_sw = np.einsum("nwa,nwl->al", B, marg - onehot)
    pair = np.exp(la[:, :-1, :, None] + tw[None, None] + (state[:, 1:] + lb[:, 1:])[:, :, None, :] - logZ[:, None, None, None])
    g_tw = pair.sum(axis=(0, 1))
    np.add.at(g_tw, (y[:, :-1].ravel(), y[:, 1:].ravel()), -1.0)
    g = np.concatenate([g_sw.ravel(), g_tw.ravel()])
    return nll + c2 * np.dot(theta, theta), g + 2.0 * c2 * theta


def fit_crf(B, y, A, c2=1.0, max_iterations=500, max_sequences=2000, seed=0):
    """Returns (state_w [A, L], trans_w [L, L])."""
    from scipy.optimize import minimize
    B = np.asarray(B, dtype=np.float64)
    y = np.asarray(y).astype(np.int64)
    if len(B) > max_sequences:
        idx = np.random.default_rng(seed).choice(len(B), max_sequences, replace=False)
        B, y = B[idx], y[idx]
    L = A
    theta0 = np.zeros(A * L + L * L)
    res = minimize(crf_nll_grad, theta0, args=(B, y, A, L, c2), jac=True, method="L-BFGS-B",
                   options={"maxiter": max_iterations})
    th = res.x
    return th[:A * L].reshape(A, L), th[A * L:].reshape(L, L)
