"""K6c parity: gnx_gnofix_crf (Gnofix with the CRF smoother -- an extension, the reference refuses the combination,
src/model.py:194) against oracle/np_oracle.py::gnofix_crf_extension, i.e. the oracle's restatement of the reference's
gnofix control flow (pinned to the reference's own gnofix by tests/golden/gnofix.npz) with the oracle's CRF plugged in:
bit-exact X, Y and tracker.  NO REFERENCE ORACLE exists for this path; the test says what the extension is defined as."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _model(C, W, A, S, crf):
    from gnomix_b200 import Gnomix
    from gnomix_b200.smooth import CRF_Smoother
    m = Gnomix.__new__(Gnomix)
    m.C, m.M, m.A, m.S, m.W = C, C // W, A, S, W
    m.smooth = CRF_Smoother(n_windows=W, num_ancestry=A, smooth_window_size=S)
    m.smooth.model = crf
    m.base = None
    return m


def _planted(rng, n_ind, W, A, C, n_switch):
    B = np.empty((2 * n_ind, W, A), dtype=np.float64)
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


def _crf(rng, A):
    from gnomix_b200.smooth import CRFModel
    return CRFModel(np.eye(A) * 4.0 + rng.normal(0, 0.3, (A, A)), np.eye(A) * 2.0 + rng.normal(0, 0.3, (A, A)))


@pytest.mark.parametrize("memo", ["", "0"])
@pytest.mark.parametrize("W,A,S,n_ind,seed", [(160, 7, 75, 5, 0), (90, 3, 11, 12, 1), (200, 5, 25, 6, 2), (64, 2, 31, 6, 3), (75, 7, 75, 3, 4)])
def test_gnofix_crf_matches_oracle(W, A, S, n_ind, seed, memo, monkeypatch):
    if memo:
        monkeypatch.setenv("GNX_GNOFIX_MEMO", memo)
    else:
        monkeypatch.delenv("GNX_GNOFIX_MEMO", raising=False)
    from oracle import c_oracle as co, np_oracle as npo
    rng = np.random.default_rng(seed)
    C = W * 23 + 7
    crf = _crf(rng, A)
    X, B = _planted(rng, n_ind, W, A, C, n_switch=4)
    model = _model(C, W, A, S, crf)
    with pytest.raises(AssertionError):   # without the explicit opt-in the reference's refusal stands
        model.phase(X, B=B)
    Xp, Yp, trk = model.phase(X, B=B, want_tracker=True, crf_extension=True)
    switched = 0
    for i in range(n_ind):
        X_m, X_p, Y_m, Y_p, t = npo.gnofix_crf_extension(X[2 * i], X[2 * i + 1], B[2 * i:2 * i + 2], S, crf.state_w, crf.trans_w,
                                                         crf_smooth_fn=co.crf_smooth)
        assert np.array_equal(Yp[2 * i:2 * i + 2], np.array([Y_m, Y_p])), "labels of individual %d" % i
        assert np.array_equal(trk[2 * i:2 * i + 2], t), "tracker of individual %d" % i
        assert np.array_equal(Xp[2 * i:2 * i + 2], np.array([X_m, X_p])), "X of individual %d" % i
        switched += int((t[0, 1:] != t[0, :-1]).sum())
    assert switched > 0, "the case exercises no switch"


def test_gnofix_crf_device_tensors_and_numpy_oracle():
    """device tensors in / out, B updated in place; the NumPy CRF of the oracle (not its C twin) as the checker"""
    import torch
    from gnomix_b200.gnofix import phase_device_crf
    from oracle import np_oracle as npo
    rng = np.random.default_rng(11)
    W, A, S, n_ind = 48, 3, 9, 3
    C = W * 5 + 3
    crf = _crf(rng, A)
    X, B = _planted(rng, n_ind, W, A, C, n_switch=3)
    model = _model(C, W, A, S, crf)
    Xd = torch.from_numpy(X).cuda()
    Bd = torch.from_numpy(B).cuda()
    Y, trk = phase_device_crf(model.smooth, Xd, Xd.stride(0), C, Bd, want_tracker=True)
    torch.cuda.synchronize()
    for i in range(n_ind):
        X_m, X_p, Y_m, Y_p, t = npo.gnofix_crf_extension(X[2 * i], X[2 * i + 1], B[2 * i:2 * i + 2], S, crf.state_w, crf.trans_w)
        assert np.array_equal(Y[2 * i:2 * i + 2].cpu().numpy(), np.array([Y_m, Y_p]))
        assert np.array_equal(trk[2 * i:2 * i + 2].cpu().numpy(), t)
        assert np.array_equal(Xd[2 * i:2 * i + 2].cpu().numpy(), np.array([X_m, X_p]))
        # B in place: window j of row h is the original row h ^ tracker
        for h in range(2):
            exp = np.stack([B[2 * i + t[h][j], j] for j in range(W)])
            assert np.array_equal(Bd[2 * i + h].cpu().numpy(), exp)


def test_crf_phase_pipeline_equals_phase_then_predict():
    """gnx_infer_host_ex with phase and a CRF smoother (crf_phase_S): the extension inside the one-call host pipeline ==
    Gnomix.phase(crf_extension=True) followed by predict_proba on the phased haplotypes; refused without the opt-in."""
    from gnomix_b200 import Gnomix, synth
    from gnomix_b200.smooth import CRF_Smoother
    from tests import util
    rng = np.random.default_rng(31)
    C, M, A, S = 24_000, 400, 4, 9
    model = Gnomix(C, M, A, S)
    coefs, icpts, _ = util.random_lr(rng, C, M, A)
    model.base.set_window_weights(coefs, icpts)
    model.smooth = CRF_Smoother(n_windows=C // M, num_ancestry=A, smooth_window_size=S)
    model.smooth.model = _crf(rng, A)
    freqs = synth.population_frequencies(rng, C, A, fst=0.3)
    fx, fpop = synth.founders(rng, freqs, per_pop=6)
    X, _ = synth.admix_host(rng, fx, fpop, 41, morgans=1.0)     # odd: the trailing haplotype is dropped
    with pytest.raises(AssertionError):
        model.predict_host(X, phase=True)
    Xp_ref, y_ref = model.phase(X, crf_extension=True)
    p_ref = model.predict_proba(Xp_ref.astype(np.int8))
    y, p, Xp = model.predict_host(X, want_proba=True, phase=True, want_phased=True, chunk_haps=16, crf_extension=True)
    assert y.shape == (40, model.W) and np.array_equal(y, y_ref)
    assert np.array_equal(Xp, Xp_ref)
    assert np.array_equal(np.ascontiguousarray(p).view(np.uint64), np.ascontiguousarray(p_ref).view(np.uint64))
