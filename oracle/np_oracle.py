"""NumPy restatement of the gnomix inference hot path (TEST INFRASTRUCTURE ONLY).

Each function cites the reference file:line (under /root/reference) it follows.
Readable and slow on purpose; `oracle/c_oracle.py` wraps a C version of the same
algorithms (oracle/gnx_oracle.c) for sizes where Python loops are too slow.
"""
from __future__ import annotations

import numpy as np

# ----------------------------------------------------------------------------
# Row A: windowing  (src/Base/base.py:41-44, 146-180)
# ----------------------------------------------------------------------------


def base_pad(X: np.ndarray, ctx: int) -> np.ndarray:
    """Base.pad, src/Base/base.py:41-44 (reflect including the edge element)."""
    if ctx == 0:
        return X
    pad_left = np.flip(X[:, 0:ctx], axis=1)
    pad_right = np.flip(X[:, -ctx:], axis=1)
    return np.concatenate([pad_left, X, pad_right], axis=1)


def base_window_ranges(C: int, M: int, ctx: int):
    """Padded-coordinate [lo, hi) of every base window.

    src/Base/base.py:157-164: windows 0..W-2 are `sliding_window_view(Xpad, M_)`
    at offsets arange(0,C,M)[:-2]; the last window is Xpad[:, -(M_+rem):].
    """
    W = C // M
    rem = C - M * W
    M_ = M + 2 * ctx
    out = [(w * M, w * M + M_) for w in range(W - 1)]
    out.append((C + 2 * ctx - (M_ + rem), C + 2 * ctx))
    return out


def padded_to_orig(p: np.ndarray, C: int, ctx: int) -> np.ndarray:
    """Index map of Base.pad: padded column p -> original SNP column."""
    p = np.asarray(p)
    return np.where(p < ctx, ctx - 1 - p, np.where(p >= ctx + C, 2 * C + ctx - 1 - p, p - ctx))


# ----------------------------------------------------------------------------
# Row B: logistic-regression base  (src/Base/models.py:12-21 -> sklearn
# LinearClassifierMixin._predict_proba_lr; OneVsRestClassifier.predict_proba)
# ----------------------------------------------------------------------------


def expit(x):
    return 1.0 / (1.0 + np.exp(-x))


def lr_window_proba(Xw: np.ndarray, coef: np.ndarray, intercept: np.ndarray) -> np.ndarray:
    """One window: decision = Xw(float64) @ coef.T + intercept; expit; row-normalise.
    coef is [A, M_w] (or [1, M_w] for the binary case), float64."""
    d = Xw.astype(np.float64) @ coef.T + intercept
    p = expit(d)
    if p.shape[1] == 1:
        return np.concatenate([1.0 - p, p], axis=1)
    p /= p.sum(axis=1).reshape((p.shape[0], -1))
    return p


def lr_base_predict_proba(X: np.ndarray, coefs, intercepts, C: int, M: int, ctx: int) -> np.ndarray:
    """Base.predict_proba_vectorized (src/Base/base.py:146-180) with LR windows.
    Returns float64 [N, W, A]."""
    Xp = base_pad(X, ctx)
    out = []
    for (lo, hi), cf, b in zip(base_window_ranges(C, M, ctx), coefs, intercepts):
        out.append(lr_window_proba(Xp[:, lo:hi], cf, b))
    return np.swapaxes(np.array(out), 0, 1)


# -- fixed-point form of row B (what the CUDA kernel evaluates exactly) ------


def lr_fold_ranges(C: int, M: int, ctx: int):
    """Original-coordinate SNP range [s_w, e_w) that window w reads after the
    reflect pads are folded onto the SNPs they mirror."""
    W = C // M
    rng = []
    for w in range(W):
        s = max(0, w * M - ctx)
        e = C if w == W - 1 else min(C, w * M + M + ctx)
        rng.append((s, e))
    return rng


def lr_choose_scale(coefs, C: int, M: int, ctx: int, limbs: int) -> int:
    """Fixed-point exponent s: q = rint(w * 2^s).  Largest s such that every folded
    |q| < 2^(8*limbs-2) and 2*sum|q| < 2^62 for every (window, class)."""
    amax = 0.0
    ssum = 0.0
    pr = base_window_ranges(C, M, ctx)
    fr = lr_fold_ranges(C, M, ctx)
    for (lo, hi), (s0, e0), cf in zip(pr, fr, coefs):
        orig = padded_to_orig(np.arange(lo, hi), C, ctx)
        for a in range(cf.shape[0]):
            f = np.zeros(e0 - s0)
            np.add.at(f, orig - s0, cf[a])
            amax = max(amax, float(np.abs(f).max()))
            ssum = max(ssum, float(np.abs(f).sum()))
    if amax == 0.0:
        return 8 * limbs - 2
    e1 = int(np.frexp(amax)[1])
    e2 = int(np.frexp(ssum)[1])
    return int(min(8 * limbs - 2 - e1, 60 - e2))


def lr_quantize_fold(coefs, C: int, M: int, ctx: int, s: int):
    """Per window: int64 [A_rows, e_w - s_w] = sum of rint(w*2^s) over the padded
    columns that map to each original SNP (quantise first, fold in integers)."""
    pr = base_window_ranges(C, M, ctx)
    fr = lr_fold_ranges(C, M, ctx)
    out = []
    for (lo, hi), (s0, e0), cf in zip(pr, fr, coefs):
        orig = padded_to_orig(np.arange(lo, hi), C, ctx)
        q = np.rint(np.asarray(cf, dtype=np.float64) * (2.0 ** s)).astype(np.int64)
        f = np.zeros((cf.shape[0], e0 - s0), dtype=np.int64)
        for a in range(cf.shape[0]):
            np.add.at(f[a], orig - s0, q[a])
        out.append(f)
    return out


def lr_fixed_logits(X: np.ndarray, qfold, intercepts, C: int, M: int, ctx: int, s: int) -> np.ndarray:
    """Exact integer dot products -> float64 logits [N, W, A_rows]."""
    fr = lr_fold_ranges(C, M, ctx)
    N = X.shape[0]
    out = np.zeros((N, len(fr), qfold[0].shape[0]))
    for w, ((s0, e0), q) in enumerate(zip(fr, qfold)):
        tot = X[:, s0:e0].astype(np.int64) @ q.T  # exact (int64)
        out[:, w, :] = tot.astype(np.float64) * (2.0 ** -s) + intercepts[w]
    return out


# ----------------------------------------------------------------------------
# Row D: slide_window  (src/Smooth/utils.py:4-29)
# ----------------------------------------------------------------------------


def smooth_pad(B: np.ndarray, S: int) -> np.ndarray:
    pad = (S + 1) // 2
    pad_left = np.flip(B[:, 0:pad, :], axis=1)
    pad_right = np.flip(B[:, -pad:, :], axis=1)
    return np.concatenate([pad_left, B, pad_right], axis=1)


def slide_window(B: np.ndarray, S: int) -> np.ndarray:
    """float32 [N*W, S*A]; row (n,w) = B_padded[n, w:w+S, :].ravel()."""
    N, W, A = B.shape
    Bp = smooth_pad(B, S)
    out = np.zeros((N, W, A * S), dtype="float32")
    for n in range(N):
        for w in range(W):
            out[n, w, :] = Bp[n, w:w + S].ravel()
    return out.reshape(N * W, A * S)


# ----------------------------------------------------------------------------
# Row E: gradient-boosted-tree smoother with xgboost multi:softprob semantics
# (src/Smooth/models.py:14-20, src/Smooth/smooth.py:40-65).  xgboost is absent
# from /root/reference (requirements.txt:11 pins xgboost==1.1.1); this restates
# its published CPU predictor: per class psum (float32, tree-index order, tree t
# belongs to class t % A) + base margin, then common/math.h::Softmax.
# ----------------------------------------------------------------------------


class GBTModel:
    """Flat xgboost-style forest.  Node arrays are concatenated over trees;
    tree t owns nodes tree_offsets[t]:tree_offsets[t+1]; child indices are
    relative to the tree's first node; feat < 0 marks a leaf."""

    def __init__(self, A, n_features, feat, thr, left, right, default_left, leaf, tree_offsets, base_margin):
        self.A = int(A)
        self.n_features = int(n_features)
        self.feat = np.ascontiguousarray(feat, dtype=np.int32)
        self.thr = np.ascontiguousarray(thr, dtype=np.float32)
        self.left = np.ascontiguousarray(left, dtype=np.int32)
        self.right = np.ascontiguousarray(right, dtype=np.int32)
        self.default_left = np.ascontiguousarray(default_left, dtype=np.uint8)
        self.leaf = np.ascontiguousarray(leaf, dtype=np.float32)
        self.tree_offsets = np.ascontiguousarray(tree_offsets, dtype=np.int32)
        self.base_margin = np.ascontiguousarray(base_margin, dtype=np.float32)

    @property
    def n_trees(self):
        return len(self.tree_offsets) - 1

    def max_depth(self):
        dmax = 0
        for t in range(self.n_trees):
            o = self.tree_offsets[t]
            stack = [(0, 0)]
            while stack:
                nid, d = stack.pop()
                if self.feat[o + nid] < 0:
                    dmax = max(dmax, d)
                else:
                    stack.append((self.left[o + nid], d + 1))
                    stack.append((self.right[o + nid], d + 1))
        return dmax


def expf_cr(x: np.ndarray) -> np.ndarray:
    """float32 exp rounded once from float64 (see include/gnx_math.h gnx_expf_cr;
    the C oracle and the kernels use the bit-reproducible gnx_exp)."""
    return np.exp(x.astype(np.float64)).astype(np.float32)


def gbt_margins(model: GBTModel, rows: np.ndarray) -> np.ndarray:
    """xgboost 1.1.1 src/predictor/cpu_predictor.cc, restated (PARITY UNPINNED until tests/golden/pin_xgboost.npz exists):
    `PredValue`: psum = 0 (bst_float), for the trees of output group gid in tree order: `tid = tree.GetLeafIndex(feats)`
    (include/xgboost/tree_model.h `GetNext`: missing -> cdefault(), else `fvalue < split_cond` -> left),
    psum += leaf value; `PredictBatchKernel` then does `preds[ridx * num_group + gid] += psum` where preds was
    initialised to base_margin (learner base_score for every class): float32 throughout, base added last."""
    rows = np.asarray(rows, dtype=np.float32)
    k = rows.shape[0]
    psum = np.zeros((k, model.A), dtype=np.float32)
    for t in range(model.n_trees):
        o = model.tree_offsets[t]
        nid = np.zeros(k, dtype=np.int64)
        active = model.feat[o + nid] >= 0
        while active.any():
            idx = o + nid[active]
            x = rows[np.nonzero(active)[0], model.feat[idx]]
            go_left = np.where(np.isnan(x), model.default_left[idx] != 0, x < model.thr[idx])
            nid[active] = np.where(go_left, model.left[idx], model.right[idx])
            active = model.feat[o + nid] >= 0
        psum[:, t % model.A] = psum[:, t % model.A] + model.leaf[o + nid]
    return (model.base_margin[None, :] + psum).astype(np.float32)


def softmax_xgb(m: np.ndarray) -> np.ndarray:
    """xgboost 1.1.1 src/common/math.h `Softmax(begin, end)` as called by src/objective/multiclass_obj.cu
    `SoftmaxMultiClassObj::Transform` (multi:softprob): wmax = max; `double wsum = 0; for each: *i = expf(*i - wmax);
    wsum += *i;` then `*i /= static_cast<float>(wsum)` -- float32 exp, float64 running sum, float32 division."""
    m = np.asarray(m, dtype=np.float32)
    wmax = m.max(axis=1, keepdims=True)
    e = expf_cr((m - wmax).astype(np.float32))
    wsum = np.zeros(len(m), dtype=np.float64)
    for c in range(m.shape[1]):
        wsum = wsum + e[:, c].astype(np.float64)
    return (e / wsum.astype(np.float32)[:, None]).astype(np.float32)


def gbt_predict_proba(model: GBTModel, rows: np.ndarray) -> np.ndarray:
    return softmax_xgb(gbt_margins(model, rows))


def xgb_smooth(model: GBTModel, B: np.ndarray, S: int):
    """Smoother.predict_proba + predict (src/Smooth/smooth.py:40-65)."""
    N, W, A = B.shape
    proba = gbt_predict_proba(model, slide_window(B, S)).reshape(-1, W, A)
    return proba, np.argmax(proba, axis=-1)


# ----------------------------------------------------------------------------
# Row F: linear-chain CRF marginals (src/Smooth/crf.py:62-67 ->
# sklearn_crfsuite.CRF.predict_marginals -> CRFsuite crf1d_context.c
# crf1dc_alpha_score / crf1dc_beta_score / crf1dc_marginal_point).
# sklearn-crfsuite==0.3.6 is absent from /root/reference; restated from the
# published algorithm.  state_w[a, y]: weight of attribute a for label y;
# trans_w[i, j]: weight of transition i -> j.
# ----------------------------------------------------------------------------


def crf_marginals(Bn: np.ndarray, state_w: np.ndarray, trans_w: np.ndarray) -> np.ndarray:
    """One sequence Bn [W, A] (float64) -> marginals [W, L] float64.
    CRFsuite 0.12 lib/crf/src/crf1d_context.c, restated (PARITY UNPINNED until tests/golden/pin_crfsuite.npz exists):
    `crf1dc_exp_state / crf1dc_exp_transition` (exp of the score tables), `crf1dc_alpha_score` (alpha[0] = exp_state[0];
    alpha[t] = (sum_i alpha[t-1][i] * exp_trans[i]) * exp_state[t]; each row scaled by 1 / its sum, scale kept),
    `crf1dc_beta_score` (beta[T-1] = scale[T-1]; beta[t][i] = sum_j exp_trans[i][j] * (beta[t+1][j] * exp_state[t+1][j]),
    times scale[t]), `crf1dc_marginal_point` (alpha[t][l] * beta[t][l] / scale[t]); state score of label y at t =
    sum over the item's attributes of value * weight (crf1d_tag.c `crf1dt_state_score`)."""
    T, A = Bn.shape
    L = state_w.shape[1]
    state = np.zeros((T, L))
    for a in range(A):  # items iterate attributes in insertion order "0","1",...
        state += Bn[:, a:a + 1] * state_w[a][None, :]
    es = np.exp(state)
    et = np.exp(trans_w)
    alpha = np.zeros((T, L))
    beta = np.zeros((T, L))
    scale = np.zeros(T)
    cur = es[0].copy()
    ssum = 0.0
    for v in cur:
        ssum += v
    scale[0] = 1.0 / ssum if ssum != 0.0 else 1.0
    alpha[0] = cur * scale[0]
    for t in range(1, T):
        cur = np.zeros(L)
        for i in range(L):
            cur += alpha[t - 1, i] * et[i]
        cur *= es[t]
        ssum = 0.0
        for v in cur:
            ssum += v
        scale[t] = 1.0 / ssum if ssum != 0.0 else 1.0
        alpha[t] = cur * scale[t]
    beta[T - 1] = scale[T - 1]
    for t in range(T - 2, -1, -1):
        row = beta[t + 1] * es[t + 1]
        cur = np.zeros(L)
        for i in range(L):
            acc = 0.0
            for j in range(L):
                acc += et[i, j] * row[j]
            cur[i] = acc
        beta[t] = cur * scale[t]
    return alpha * beta / scale[:, None]


def crf_smooth(B: np.ndarray, state_w, trans_w):
    proba = np.array([crf_marginals(np.asarray(b, dtype=np.float64), state_w, trans_w) for b in B])
    return proba, np.argmax(proba, axis=-1)


def crf_bruteforce_marginals(Bn, state_w, trans_w):
    """Enumerate all L^T label paths (tiny T only) -- independent pin for crf_marginals."""
    import itertools
    T, A = Bn.shape
    L = state_w.shape[1]
    state = Bn @ state_w
    marg = np.zeros((T, L))
    Z = 0.0
    for path in itertools.product(range(L), repeat=T):
        sc = sum(state[t, y] for t, y in enumerate(path))
        sc += sum(trans_w[path[t], path[t + 1]] for t in range(T - 1))
        p = np.exp(sc)
        Z += p
        for t, y in enumerate(path):
            marg[t, y] += p
    return marg / Z


# ----------------------------------------------------------------------------
# Row C: CovRSK string kernel + libsvm probability
# (src/Base/string_kernel.py:75-123, src/Base/models.py:195-215)
# ----------------------------------------------------------------------------


def cov_sample(M: int, alpha: float = 0.6, beta: float = 1.0, seed: int = 37):
    """CovSample, src/Base/string_kernel.py:80-89 (consumes the legacy global RNG
    stream the same way: one np.random.rand() per m)."""
    rs = np.random.RandomState(seed)
    u = rs.rand(max(M - 1, 0))
    Ms = [1]
    for i, m in enumerate(range(2, M + 1)):
        if (1 - (alpha ** (m - Ms[-1] + 1))) * (m ** (-beta)) >= u[i]:
            Ms.append(m)
    return Ms


def covrsk_closed_form(x: np.ndarray, Y: np.ndarray, Ms) -> np.ndarray:
    """K(x,y) = sum over maximal match runs of length Lr of G(Lr),
    G(L) = sum_{m in Ms, m<=L} (L-m+1)  (SURVEY.md 8(a) row C closed form of
    CovRSK_DP_triangular_numbers_vectorized, string_kernel.py:91-101)."""
    Ms = np.asarray(Ms)
    Mlen = x.shape[0]
    G = np.zeros(Mlen + 1, dtype=np.int64)
    for L in range(1, Mlen + 1):
        mm = Ms[Ms <= L]
        G[L] = int((L - mm + 1).sum())
    out = np.zeros(len(Y), dtype=np.int64)
    for r, y in enumerate(Y):
        z = np.concatenate([[0], (x == y).astype(np.int8), [0]])
        d = np.diff(z)
        starts = np.nonzero(d == 1)[0]
        ends = np.nonzero(d == -1)[0]
        out[r] = int(G[ends - starts].sum())
    return out


def covrsk_kernel(X: np.ndarray, Y: np.ndarray, Ms=None) -> np.ndarray:
    if Ms is None:
        Ms = cov_sample(X.shape[1])
    return np.array([covrsk_closed_form(x, Y, Ms) for x in X])


def svc_predict_proba(K: np.ndarray, n_support, dual_coef, intercept, probA, probB) -> np.ndarray:
    """libsvm svm_predict_probability for a precomputed kernel row block
    K [n, nSV] (columns already restricted to support vectors, grouped by class).
    dual_coef = SVC._dual_coef_ [k-1, nSV], intercept = SVC._intercept_.
    (libsvm svm.cpp sigmoid_predict + multiclass_probability)."""
    k = len(n_support)
    start = np.concatenate([[0], np.cumsum(n_support)])[:-1]
    n = K.shape[0]
    out = np.zeros((n, k))
    min_prob = 1e-7
    for r in range(n):
        kv = K[r].astype(np.float64)
        pair = np.zeros((k, k))
        p = 0
        for i in range(k):
            for j in range(i + 1, k):
                si, sj = start[i], start[j]
                ci, cj = n_support[i], n_support[j]
                ssum = 0.0
                for t in range(ci):
                    ssum += dual_coef[j - 1, si + t] * kv[si + t]
                for t in range(cj):
                    ssum += dual_coef[i, sj + t] * kv[sj + t]
                dec = ssum + intercept[p]
                # sklearn negates libsvm's rho into _intercept_ and flips sign conventions for
                # binary problems only on the public attributes; the underscore ones are libsvm's.
                fApB = dec * probA[p] + probB[p]
                if fApB >= 0:
                    v = np.exp(-fApB) / (1.0 + np.exp(-fApB))
                else:
                    v = 1.0 / (1 + np.exp(fApB))
                v = min(max(v, min_prob), 1 - min_prob)
                pair[i, j] = v
                pair[j, i] = 1 - v
                p += 1
        if k == 2:
            out[r] = [pair[0, 1], pair[1, 0]]
            continue
        out[r] = multiclass_probability(k, pair)
    return out


def multiclass_probability(k: int, r: np.ndarray) -> np.ndarray:
    """libsvm multiclass_probability (Wu, Lin, Weng 2004, method 2)."""
    max_iter = max(100, k)
    Q = np.zeros((k, k))
    Qp = np.zeros(k)
    p = np.full(k, 1.0 / k)
    eps = 0.005 / k
    for t in range(k):
        Q[t, t] = 0.0
        for j in range(t):
            Q[t, t] += r[j, t] * r[j, t]
            Q[t, j] = Q[j, t]
        for j in range(t + 1, k):
            Q[t, t] += r[j, t] * r[j, t]
            Q[t, j] = -r[j, t] * r[t, j]
    for _ in range(max_iter):
        pQp = 0.0
        for t in range(k):
            Qp[t] = 0.0
            for j in range(k):
                Qp[t] += Q[t, j] * p[j]
            pQp += p[t] * Qp[t]
        max_error = 0.0
        for t in range(k):
            error = abs(Qp[t] - pQp)
            if error > max_error:
                max_error = error
        if max_error < eps:
            break
        for t in range(k):
            diff = (-Qp[t] + pQp) / Q[t, t]
            p[t] += diff
            pQp = (pQp + diff * (diff * Q[t, t] + 2 * Qp[t])) / (1 + diff) / (1 + diff)
            for j in range(k):
                Qp[j] = (Qp[j] + diff * Q[t, j]) / (1 + diff)
                p[j] /= (1 + diff)
    return p


# ----------------------------------------------------------------------------
# Row G: gnofix with the reference's defaults (src/Gnofix/gnofix.py:58-208,
# src/Gnofix/phasing.py:182-198, called from src/model.py:188-214)
# ----------------------------------------------------------------------------- calibrator
# Calibrator.transform (reference src/Smooth/Calibration.py:57-69) with the normalisation of
# lines 24-39.  Each class model is sklearn IsotonicRegression(out_of_bounds="clip"); its
# transform (sklearn/isotonic.py::_transform) casts the input to the dtype of the fitted
# thresholds, clips to [X_min_, X_max_], interpolates with scipy interp1d(kind="linear") and
# casts back to that dtype.  interp1d delegates to numpy.interp when the thresholds are float64
# (CRF smoother probabilities) and otherwise runs its own two-weight formula in the
# thresholds' dtype (float32 thresholds = XGB smoother probabilities).

def np_interp_restated(x, xp, fp):
    """numpy.interp for x inside [xp[0], xp[-1]] (numpy/_core/src/multiarray/compiled_base.c,
    arr_interp): j with xp[j] <= x < xp[j+1]; exact hit -> fp[j]; else
    slope*(x - xp[j]) + fp[j] with slope = (fp[j+1]-fp[j])/(xp[j+1]-xp[j]), all float64."""
    x = np.asarray(x, dtype=np.float64)
    xp = np.asarray(xp, dtype=np.float64)
    fp = np.asarray(fp, dtype=np.float64)
    n = len(xp)
    out = np.full(x.shape, np.nan)
    ok = ~np.isnan(x)
    xv = x[ok]
    j = np.searchsorted(xp, xv, side="right") - 1
    last = j >= n - 1
    jj = np.clip(j, 0, max(n - 2, 0))
    if n == 1:
        res = np.full(xv.shape, fp[0])
    else:
        with np.errstate(invalid="ignore", divide="ignore"):
            slope = (fp[jj + 1] - fp[jj]) / (xp[jj + 1] - xp[jj])
            res = slope * (xv - xp[jj]) + fp[jj]
            alt = slope * (xv - xp[jj + 1]) + fp[jj + 1]
        bad = np.isnan(res)
        res = np.where(bad, alt, res)
        res = np.where(np.isnan(res) & (fp[jj] == fp[jj + 1]), fp[jj], res)
        res = np.where(xp[jj] == xv, fp[jj], res)
        res = np.where(last, fp[n - 1], res)
    out[ok] = res
    return out


def interp1d_linear_restated(x, xp, fp):
    """scipy.interpolate.interp1d._call_linear in the arrays' own dtype (float32 here):
    hi = clip(searchsorted(xp, x, 'left'), 1, n-1), lo = hi-1,
    y = ((x - x_lo)/(x_hi - x_lo))*y_hi + ((x_hi - x)/(x_hi - x_lo))*y_lo."""
    hi = np.clip(np.searchsorted(xp, x, side="left"), 1, len(xp) - 1)
    lo = hi - 1
    x_lo, x_hi, y_lo, y_hi = xp[lo], xp[hi], fp[lo], fp[hi]
    with np.errstate(invalid="ignore", divide="ignore"):
        return ((x - x_lo) / (x_hi - x_lo)) * y_hi + ((x_hi - x) / (x_hi - x_lo)) * y_lo


def isotonic_transform(t, x_thr, y_thr):
    """IsotonicRegression(out_of_bounds="clip").transform for fitted thresholds."""
    dt = x_thr.dtype
    t = np.asarray(t).astype(dt)
    t = np.clip(t, x_thr[0], x_thr[-1])
    if len(y_thr) == 1:
        return np.repeat(y_thr, t.shape[0]).astype(dt)
    if dt == np.float64 and y_thr.dtype == np.float64:
        return np_interp_restated(t, x_thr, y_thr).astype(dt)
    return interp1d_linear_restated(t, x_thr, y_thr.astype(dt)).astype(dt)


def calibrator_transform(proba, thresholds):
    """proba [..., A]; thresholds = [(X_thresholds_, y_thresholds_)] per class.  Returns float64."""
    A = len(thresholds)
    shape = proba.shape
    flat = proba.reshape(-1, A)
    iso = np.zeros((flat.shape[0], A))
    for i, (xt, yt) in enumerate(thresholds):
        iso[:, i] = isotonic_transform(flat[:, i], xt, yt)
    with np.errstate(invalid="ignore", divide="ignore"):
        if A == 2:
            iso[:, 0] = 1. - iso[:, 1]
        else:
            iso /= np.sum(iso, axis=1)[:, np.newaxis]
    iso[np.isnan(iso)] = 1. / A
    iso[(1.0 < iso) & (iso <= 1.0 + 1e-5)] = 1.0
    return iso.reshape(*shape)


# ----------------------------------------------------------------------------


def mode_filter(pred: np.ndarray, size) -> np.ndarray:
    """mode_filter + mode of src/Smooth/utils.py:31-46 for one label row: positions ends <= i < len - ends
    (ends = size // 2) become the most frequent value of pred[i-ends : i+ends+1] when the smallest and the largest
    most-frequent value coincide (scipy.stats.mode returns the smallest; the reference calls it on arr and -arr),
    else the window's centre value; `range(len)[ends:-ends]` is empty for ends == 0, so size 1 / True changes nothing."""
    pred = np.asarray(pred)
    if not size:
        return pred
    out = np.copy(pred)
    ends = int(size) // 2
    for i in range(len(pred))[ends:-ends]:
        arr = pred[i - ends:i + ends + 1]
        vals, cnt = np.unique(arr, return_counts=True)
        top = vals[cnt == cnt.max()]
        out[i] = top[0] if len(top) == 1 else arr[len(arr) // 2]
    return out


def gnofix_default(X_m, X_p, B, S, predict_rows, smooth_predict, max_it=50):
    """gnofix(M,P,B,smoother) with every keyword at its default.
    predict_rows(rows[k,S*A]) -> proba[k,A]  (smoother.model.predict_proba)
    smooth_predict(B[2,W,A]) -> labels[2,W]  (smoother.predict)
    Returns X_m, X_p, Y_m, Y_p, tracker(2,W)."""
    _, W, A = B.shape
    window_size = len(X_m) // W
    X_m = np.array(X_m).astype(int).copy()
    X_p = np.array(X_p).astype(int).copy()
    B = np.array(B).copy()
    Y_m, Y_p = smooth_predict(B).reshape(2, W)
    half = (S - 1) // 2
    c_lo, c_hi = half, W - S + half  # centers[0], centers[-1]
    trk_m, trk_p = np.zeros(W, dtype=int), np.ones(W, dtype=int)
    X_m_its = []
    for _ in range(max_it):
        if any(np.all(X_m == prev) for prev in X_m_its):
            break
        X_m_its.append(X_m.copy())
        for w in range(1, W):
            if Y_m[w] != Y_m[w - 1] or Y_p[w] != Y_p[w - 1]:
                center = min(max(w, c_lo), c_hi)
                lo = center - half
                m_orig = B[0, lo:lo + S].copy()
                p_orig = B[1, lo:lo + S].copy()
                m_sw = np.concatenate([B[0, lo:w], B[1, w:lo + S]])
                p_sw = np.concatenate([B[1, lo:w], B[0, w:lo + S]])
                mps = np.array([m_orig, p_orig, m_sw, p_sw])
                outs = predict_rows(mps.reshape(4, -1)).reshape(-1, 2, A)
                probs = np.max(np.max(outs, axis=2), axis=1)
                if probs[1] * 0.5 > probs[0] * (1 - 0.5):
                    m = np.concatenate([B[0, :w], B[1, w:]])
                    p = np.concatenate([B[1, :w], B[0, w:]])
                    B = np.array([m, p])
                    trk_m, trk_p = (np.concatenate([trk_m[:w], trk_p[w:]]),
                                    np.concatenate([trk_p[:w], trk_m[w:]]))
                    i = w * window_size
                    tmp = X_m.copy()
                    X_m[i:] = X_p[i:]
                    X_p[i:] = tmp[i:]
                    Y_m, Y_p = smooth_predict(B).reshape(2, W)
    return X_m, X_p, Y_m, Y_p, np.array([trk_m, trk_p])


def gnofix_crf_extension(X_m, X_p, B, S, state_w, trans_w, crf_smooth_fn=None, max_it=50):
    """Gnofix with the CRF smoother -- AN EXTENSION, NO REFERENCE ORACLE: the reference refuses the combination
    (src/model.py:194; src/Gnofix/gnofix.py:157 needs `smoother.model.predict_proba` on flattened S-window rows).
    Defined as SURVEY.md section 8a row G words it: the reference's gnofix control flow (`gnofix_default` above, which
    is pinned to the reference's own gnofix) with
        smooth_predict(B[2, W, A])  = argmax of the CRF marginals of each whole chain (CRF_Smoother.predict),
        predict_rows(rows[k, S*A])  = the CRF marginal at the centre item (S - 1) // 2 of each scope run as a chain
                                      of S items of its own.
    crf_smooth_fn(B[n, T, A], state_w, trans_w) -> (marginals, labels); default: `crf_smooth` of this module
    (tests pass the C twin, same bits).  B is float64, as the CRF smoother reads it."""
    fn = crf_smooth_fn or crf_smooth
    A = np.asarray(B).shape[-1]
    c = (S - 1) // 2

    def predict_rows(rows):
        marg, _ = fn(np.asarray(rows, dtype=np.float64).reshape(-1, S, A), state_w, trans_w)
        return np.asarray(marg)[:, c, :]

    def smooth_predict(b):
        return np.asarray(fn(np.asarray(b, dtype=np.float64), state_w, trans_w)[1])

    return gnofix_default(X_m, X_p, np.asarray(B, dtype=np.float64), S, predict_rows, smooth_predict, max_it=max_it)
