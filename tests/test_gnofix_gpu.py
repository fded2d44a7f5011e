"""K6 parity: gnx_gnofix against (i) the golden outputs of the reference's own gnofix and
(ii) the oracle restatement on larger seeded cases -- bit-exact X, Y and tracker."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def _model(C, W, A, S, forest):
    from gnomix_b200 import Gnomix
    M = C // W
    if C % M == 0:
        M -= 0
    m = Gnomix.__new__(Gnomix)
    from gnomix_b200.smooth import XGB_Smoother
    m.C, m.M, m.A, m.S, m.W = C, M, A, S, W
    m.smooth = XGB_Smoother(n_windows=W, num_ancestry=A, smooth_window_size=S)
    m.smooth.model = forest
    m.base = None
    return m


def _forest_from(d):
    from gnomix_b200 import GBTForest
    return GBTForest(int(d["A"]), int(d["S"]) * int(d["A"]), d["feat"], d["thr"], d["left"], d["right"], d["default_left"],
                     d["leaf"], d["tree_offsets"], d["base_margin"])


# "" = default (rejected checks are remembered until their scope changes), "0" = every check re-run: same results
@pytest.mark.parametrize("memo", ["", "0"])
def test_gnofix_matches_reference_golden(memo, monkeypatch):
    if memo:
        monkeypatch.setenv("GNX_GNOFIX_MEMO", memo)
    else:
        monkeypatch.delenv("GNX_GNOFIX_MEMO", raising=False)
    d = np.load(os.path.join(G, "gnofix.npz"))
    W, A, S, C = int(d["W"]), int(d["A"]), int(d["S"]), int(d["C"])
    model = _model(C, W, A, S, _forest_from(d))
    n = len(d["X"])
    X = d["X"].reshape(2 * n, C)
    B = d["B"].reshape(2 * n, W, A)
    Xp, Yp, trk = model.phase(X, B=B, want_tracker=True)
    assert np.array_equal(Xp, d["X_out"].reshape(2 * n, C))
    assert np.array_equal(Yp, d["Y_out"].reshape(2 * n, W))
    assert np.array_equal(trk, d["tracker"].reshape(2 * n, W))


def _planted(rng, n_ind, W, A, C, n_switch):
    """Pairs with clean ancestry tracks whose tails were exchanged at random windows."""
    B = np.empty((2 * n_ind, W, A), dtype=np.float32)
    for i in range(n_ind):
        anc = np.zeros((2, W), dtype=int)
        for h in range(2):
            cuts = np.sort(rng.integers(1, W, 3))
            vals = rng.integers(0, A, 4)
            anc[h] = vals[np.searchsorted(cuts, np.arange(W), side="right")]
        b = 0.05 + rng.random((2, W, A)) * 0.25
        for h in range(2):
            b[h, np.arange(W), anc[h]] += 0.6
        for sw in rng.integers(2, W - 2, n_switch):
            b[:, sw:] = b[::-1, sw:].copy()
        B[2 * i:2 * i + 2] = b / b.sum(-1, keepdims=True)
    X = rng.integers(0, 2, size=(2 * n_ind, C)).astype(np.int8)
    return X, B


# memo: "" = default (rejected checks remembered), "0" = every check re-run.  split: "" = default (re-smoothing as (row, class)
# chain tasks), "0" = one thread per row.  Same results in every combination.
@pytest.mark.parametrize("split", ["", "0"])
@pytest.mark.parametrize("memo", ["", "0"])
@pytest.mark.parametrize("W,A,S,n_ind,seed", [(160, 7, 75, 5, 0), (90, 3, 11, 12, 1), (200, 5, 25, 5, 2), (64, 2, 31, 6, 3)])
def test_gnofix_matches_oracle(W, A, S, n_ind, seed, memo, split, monkeypatch):
    if memo:
        monkeypatch.setenv("GNX_GNOFIX_MEMO", memo)
    else:
        monkeypatch.delenv("GNX_GNOFIX_MEMO", raising=False)
    if split:
        monkeypatch.setenv("GNX_GNOFIX_SPLIT", split)
    else:
        monkeypatch.delenv("GNX_GNOFIX_SPLIT", raising=False)
    from gnomix_b200 import GBTForest
    from oracle import c_oracle as co, np_oracle as npo
    rng = np.random.default_rng(seed)
    C = W * 23 + 7
    forest = GBTForest.random(rng, A, S, n_rounds=40, depth=4)
    forest.thr[:] = (forest.thr * 0.7).astype(np.float32)
    X, B = _planted(rng, n_ind, W, A, C, n_switch=4)
    # identical haplotypes over some blocks exercise the "X_m repeats" rule
    X[1::2, : C // 3] = X[0::2, : C // 3]
    model = _model(C, W, A, S, forest)
    Xp, Yp, trk = model.phase(X, B=B, want_tracker=True)
    rows_fn = lambda rows: co.gbt_rows(forest, rows)
    smooth_fn = lambda b: co.gbt_smooth(forest, b, S, want_proba=False)[1]
    n_sw = 0
    for i in range(n_ind):
        X_m, X_p, Y_m, Y_p, t = npo.gnofix_default(X[2 * i], X[2 * i + 1], B[2 * i:2 * i + 2], S, rows_fn, smooth_fn)
        assert np.array_equal(Yp[2 * i:2 * i + 2], np.array([Y_m, Y_p])), "labels, individual %d" % i
        assert np.array_equal(trk[2 * i:2 * i + 2], t), "tracker, individual %d" % i
        assert np.array_equal(Xp[2 * i:2 * i + 2], np.array([X_m, X_p])), "X, individual %d" % i
        n_sw += int((t[0][:-1] != t[0][1:]).sum())
    assert n_sw >= n_ind  # switches really happen


def test_gnofix_device_resident_and_B_permuted():
    import torch
    from gnomix_b200 import GBTForest
    from gnomix_b200.gnofix import phase_device
    rng = np.random.default_rng(9)
    W, A, S, n_ind = 150, 7, 75, 8
    C = W * 11 + 3
    forest = GBTForest.random(rng, A, S, n_rounds=30, depth=4)
    X, B = _planted(rng, n_ind, W, A, C, n_switch=3)
    model = _model(C, W, A, S, forest)
    ld = (C + 127) // 128 * 128
    Xd = torch.zeros((2 * n_ind, ld), dtype=torch.int8, device="cuda")
    Xd[:, :C] = torch.from_numpy(X).cuda()
    Bd = torch.from_numpy(B).cuda()
    Y, trk = phase_device(model.smooth, Xd, ld, C, Bd, want_tracker=True)
    t = trk.cpu().numpy()[0::2].astype(bool)           # [n, W]: m row comes from the original p
    B2 = B.reshape(n_ind, 2, W, A)
    want_m = np.where(t[:, :, None], B2[:, 1], B2[:, 0])
    want_p = np.where(t[:, :, None], B2[:, 0], B2[:, 1])
    got = Bd.cpu().numpy().reshape(n_ind, 2, W, A)
    assert np.array_equal(got[:, 0], want_m) and np.array_equal(got[:, 1], want_p)
    # final labels are the smoother's labels of the final B
    assert np.array_equal(model.smooth.predict(Bd).cpu().numpy(), Y.cpu().numpy())


def test_gnofix_refuses_nan_base_probabilities():
    """A pair whose base probabilities hold a NaN is refused (labels -1 from the C ABI, ValueError from the plugin) rather
    than phased with NaN semantics that differ from the smoother's default-child rule; the other pairs are unaffected."""
    import ctypes as C
    import torch
    from gnomix_b200 import GBTForest, _lib
    from oracle import c_oracle as co, np_oracle as npo
    rng = np.random.default_rng(9)
    W, A, S, n_ind = 90, 3, 11, 4
    Cc = W * 23 + 7
    forest = GBTForest.random(rng, A, S, n_rounds=30, depth=4)
    X, B = _planted(rng, n_ind, W, A, Cc, n_switch=3)
    B[2, 40, 1] = np.nan                       # individual 1
    model = _model(Cc, W, A, S, forest)
    with pytest.raises(ValueError, match="hold NaN"):
        model.phase(X, B=B)
    Xd, Bd = torch.from_numpy(X).cuda(), torch.from_numpy(B).cuda()
    Y = torch.empty((2 * n_ind, W), dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(_lib.lib().gnx_gnofix(forest.handle(S), Xd.data_ptr(), Xd.stride(0), Cc, Bd.data_ptr(), n_ind, W, 50, Y.data_ptr(), None, st))
    Yh = Y.cpu().numpy()
    assert (Yh[2:4] == -1).all()
    assert np.array_equal(Xd[2:4].cpu().numpy(), X[2:4])
    rows_fn = lambda rows: co.gbt_rows(forest, rows)
    smooth_fn = lambda b: co.gbt_smooth(forest, b, S, want_proba=False)[1]
    for i in (0, 2, 3):
        X_m, X_p, Y_m, Y_p, t = npo.gnofix_default(X[2 * i], X[2 * i + 1], B[2 * i:2 * i + 2], S, rows_fn, smooth_fn)
        assert np.array_equal(Yh[2 * i:2 * i + 2], np.array([Y_m, Y_p]))
        assert np.array_equal(Xd[2 * i:2 * i + 2].cpu().numpy(), np.array([X_m, X_p]))
