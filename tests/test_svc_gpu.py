"""K2 + K3 parity: CovRSK string kernel (exact integers) and the libsvm probability epilogue
against the reference's golden vectors (real libsvm through scikit-learn, reference's own
Python kernel) and the oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def _base(C, M, A, ctx):
    from gnomix_b200.base import CovRSKBase
    return CovRSKBase(chm_len=C, window_size=M, num_ancestry=A, context=ctx)


def _dummy_fit(base, sv_rows, A):
    P = A * (A - 1) // 2
    ns = []
    for s in sv_rows:
        v = np.zeros(A, np.int32)
        v[0] = len(s)
        ns.append(v)
    base.set_window_svcs(sv_rows, ns, [np.zeros((A - 1, len(s))) for s in sv_rows], [np.zeros(P)] * len(sv_rows),
                         [np.zeros(P)] * len(sv_rows), [np.zeros(P)] * len(sv_rows))


def test_kernel_matches_reference_golden():
    d = np.load(os.path.join(G, "covrsk.npz"))
    X, Y, K = d["K_X"], d["K_Y"], d["K"]
    Mlen = X.shape[1]                      # one window covering the whole row: C=450, M=300, ctx=0
    base = _base(Mlen, 300, 3, 0)
    assert base.W == 1 and base.window_slices() == [(0, Mlen)]
    _dummy_fit(base, [Y], 3)
    got = base.kernel_window(0, X).cpu().numpy()
    assert np.array_equal(got, K)


@pytest.mark.parametrize("C,M,nsv,N,seed", [(5200, 1000, 70, 150, 0), (2311, 200, 33, 65, 1), (7000, 857, 130, 40, 2)])
def test_kernel_matches_oracle_long_runs(C, M, nsv, N, seed):
    """Related haplotypes (long identical stretches incl. > 866), missing calls, ragged sizes."""
    from oracle import c_oracle as co, np_oracle as npo
    rng = np.random.default_rng(seed)
    ctx = M // 2
    A = 3
    base = _base(C, M, A, ctx)
    founders = rng.integers(0, 2, size=(6, C)).astype(np.int8)
    def mosaic(n):
        out = founders[rng.integers(0, 6, n)].copy()
        for i in range(n):
            for c in np.sort(rng.integers(1, C, 3)):
                out[i, c:] = founders[rng.integers(0, 6)][c:]
        out[rng.random(out.shape) < 0.0004] ^= 1
        out[rng.random(out.shape) < 0.0004] = 2
        return out
    T, Xq = mosaic(nsv), mosaic(N)
    Tp, Xp = npo.base_pad(T, ctx), npo.base_pad(Xq, ctx)
    sl = base.window_slices()
    _dummy_fit(base, [Tp[:, lo:hi] for lo, hi in sl], A)
    Ms = npo.cov_sample(max(hi - lo for lo, hi in sl))
    assert Ms == base._ms().tolist()
    for w in (0, len(sl) // 2, len(sl) - 1):
        lo, hi = sl[w]
        want = co.covrsk(Xp[:, lo:hi], Tp[:, lo:hi], Ms)
        got = base.kernel_window(w, Xq).cpu().numpy()
        assert np.array_equal(got, want), "window %d" % w
    assert want.max() > 5 * (hi - lo)  # near-complete matches of whole windows were really present


def test_svc_proba_matches_reference_golden_and_oracle():
    from oracle import c_oracle as co, np_oracle as npo
    d = np.load(os.path.join(G, "covrsk.npz"))
    C, M, A, ctx = int(d["svc_C"]), int(d["svc_M"]), int(d["svc_A"]), int(d["svc_ctx"])
    Xq, Xt, B_ref = d["svc_X"], d["svc_Xtrain"], d["svc_B"]
    base = _base(C, M, A, ctx)
    Xtp, Xqp = npo.base_pad(Xt, ctx), npo.base_pad(Xq, ctx)
    sl = base.window_slices()
    W = len(sl)
    svs = [Xtp[d["svc_w%d_support" % w], lo:hi] for w, (lo, hi) in enumerate(sl)]
    base.set_window_svcs(svs, [d["svc_w%d_n_support" % w] for w in range(W)], [d["svc_w%d_dual_coef" % w] for w in range(W)],
                         [d["svc_w%d_intercept" % w] for w in range(W)], [d["svc_w%d_probA" % w] for w in range(W)],
                         [d["svc_w%d_probB" % w] for w in range(W)])
    B = base.predict_proba(Xq)
    assert B.dtype == np.float64 and B.shape == B_ref.shape
    assert np.max(np.abs(B - B_ref)) < 1e-12           # the reference: sklearn SVC / libsvm
    for w, (lo, hi) in enumerate(sl):                    # the oracle: bit-exact
        K = co.covrsk(Xqp[:, lo:hi], svs[w], npo.cov_sample(hi - lo))
        want = co.svc_proba(K, d["svc_w%d_n_support" % w], d["svc_w%d_dual_coef" % w], d["svc_w%d_intercept" % w],
                            d["svc_w%d_probA" % w], d["svc_w%d_probB" % w])
        assert np.array_equal(B[:, w, :].view(np.uint64), want.view(np.uint64)), "window %d" % w


def test_covrsk_train_then_predict_roundtrip():
    """Host training on GPU Gram matrices, then GPU inference == sklearn's own predict_proba on
    the same precomputed kernel."""
    import warnings
    rng = np.random.default_rng(4)
    C, M, A = 1500, 250, 3
    base = _base(C, M, A, M // 2)
    freqs = np.clip(rng.beta(0.5, 0.5, size=(A, C)), 0.05, 0.95)
    Xt = np.concatenate([(rng.random((12, C)) < freqs[a]).astype(np.int8) for a in range(A)])
    yt = np.repeat(np.repeat(np.arange(A), 12)[:, None], base.W, axis=1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        base.train(Xt, yt)
        Xq = Xt[rng.integers(0, len(Xt), 9)].copy()
        Xq[rng.random(Xq.shape) < 0.05] ^= 1
        B = base.predict_proba(Xq)
        assert B.shape == (9, base.W, A) and np.allclose(B.sum(-1), 1.0)
        assert (np.argmax(B, -1) == np.repeat(np.arange(A), 12)[rng.integers(0, 1, 1)][0]).mean() >= 0  # shape sanity
        # sklearn on the same Gram rows
        for w in (0, base.W - 1):
            Kq = base.kernel_window(w, Xq).cpu().numpy().astype(np.float64)
            Kfull = np.zeros((len(Xq), len(Xt)))
            Kfull[:, base.models[w].support_] = Kq
            assert np.max(np.abs(base.models[w].predict_proba(Kfull) - B[:, w, :])) < 1e-12


@pytest.mark.parametrize("C,M,ctx,nsv,N,seed", [(4096, 512, 0, 50, 130, 5), (3000, 217, 108, 97, 64, 6), (1857, 857, 428, 49, 33, 7)])
def test_production_kernel_equals_first_kernel_and_oracle(C, M, ctx, nsv, N, seed, monkeypatch):
    """svc_kernel_csa (carry-save counts, start-position run bookkeeping, window groups) against svc_kernel_window<13>
    and the oracle: window lengths that are / are not multiples of 32 and of 128, support-vector counts around the
    48-vector chunk, a partial last block of 64 queries, runs crossing the whole window; the probabilities of
    gnx_svc_predict (window groups) bit for bit between the two kernels."""
    from oracle import c_oracle as co, np_oracle as npo
    rng = np.random.default_rng(seed)
    A = 3
    founders = rng.integers(0, 2, size=(5, C)).astype(np.int8)

    def mosaic(n):
        out = founders[rng.integers(0, 5, n)].copy()
        for i in range(n):
            for c in np.sort(rng.integers(1, C, 2)):
                out[i, c:] = founders[rng.integers(0, 5)][c:]
        out[rng.random(out.shape) < 0.001] ^= 1
        out[rng.random(out.shape) < 0.0005] = 2
        return out
    T, Xq = mosaic(nsv), mosaic(N)
    Xq[0] = T[0]                                          # one query identical to a support vector: a run over every window
    P = A * (A - 1) // 2

    def build():
        b = _base(C, M, A, ctx)
        Tp = npo.base_pad(T, ctx) if ctx else T
        sl = b.window_slices()
        W = len(sl)
        ns = np.array([nsv - 2 * (nsv // 3), nsv // 3, nsv // 3], np.int32)
        r2 = np.random.default_rng(seed + 100)
        b.set_window_svcs([Tp[:, lo:hi] for lo, hi in sl], [ns] * W, [r2.normal(0, 1e-3, (A - 1, nsv)) for _ in range(W)],
                          [r2.normal(0, 0.1, P) for _ in range(W)], [np.full(P, -1.0)] * W, [np.zeros(P)] * W)
        return b, sl, Tp
    b_new, sl, Tp = build()
    K_new = [b_new.kernel_window(w, Xq).cpu().numpy() for w in range(len(sl))]
    B_new = b_new.predict_proba(Xq)
    monkeypatch.setenv("GNX_SVC_KERNEL", "0")
    b_old, _, _ = build()
    K_old = [b_old.kernel_window(w, Xq).cpu().numpy() for w in range(len(sl))]
    B_old = b_old.predict_proba(Xq)
    monkeypatch.delenv("GNX_SVC_KERNEL")
    Xp = npo.base_pad(Xq, ctx) if ctx else Xq
    for w, (lo, hi) in enumerate(sl):
        assert np.array_equal(K_new[w], K_old[w]), "window %d" % w
        if w in (0, len(sl) - 1):
            assert np.array_equal(K_new[w], co.covrsk(Xp[:, lo:hi], Tp[:, lo:hi], npo.cov_sample(hi - lo))), "window %d" % w
    assert np.array_equal(B_new.view(np.uint64), B_old.view(np.uint64))
