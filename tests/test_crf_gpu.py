"""K5 parity: gnx_crf_smooth against the oracle -- bit-exact float64 marginals and labels."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _smoother(W, A, sw, tw):
    from gnomix_b200.smooth import CRF_Smoother, CRFModel
    sm = CRF_Smoother(n_windows=W, num_ancestry=A, smooth_window_size=75)
    sm.model = CRFModel(sw, tw)
    return sm


# "" = default (lane-parallel kernel for A <= 8), "0" = thread-per-haplotype kernel: same bits
@pytest.mark.parametrize("crf_kernel", ["", "0"])
@pytest.mark.parametrize("W,A,N", [(1430, 7, 70), (317, 7, 129), (50, 3, 10), (12, 2, 3), (200, 12, 5), (1, 7, 4), (33, 8, 37)])
def test_crf_matches_oracle(W, A, N, crf_kernel, monkeypatch):
    from oracle import c_oracle as co
    if crf_kernel:
        monkeypatch.setenv("GNX_CRF_KERNEL", crf_kernel)
    else:
        monkeypatch.delenv("GNX_CRF_KERNEL", raising=False)
    rng = np.random.default_rng(W + A)
    sw, tw = rng.normal(0, 1.5, size=(A, A)), rng.normal(0, 1.0, size=(A, A))
    B = rng.dirichlet(np.full(A, 0.4), size=(N, W))
    sm = _smoother(W, A, sw, tw)
    proba = sm.predict_proba(B)
    label = sm.predict(B)
    p_o, l_o = co.crf_smooth(B, sw, tw)
    assert proba.dtype == np.float64 and proba.shape == (N, W, A)
    assert np.array_equal(proba.view(np.uint64), p_o.view(np.uint64))
    assert np.array_equal(label, l_o)
    assert np.allclose(proba.sum(-1), 1.0, atol=1e-9)


def test_crf_after_lr_base_f64_on_device():
    """LR base (float64 out) -> CRF, device-resident end to end (BASELINE config 5 front half)."""
    import torch
    from oracle import c_oracle as co
    from tests import util
    rng = np.random.default_rng(8)
    C, M, A, N = 9000, 300, 7, 40
    coefs, icpts, ctx = util.random_lr(rng, C, M, A)
    X = util.random_haplotypes(rng, N, C)
    base = util.make_lr_base(C, M, A, coefs, icpts)
    W = C // M
    sw, tw = rng.normal(0, 2, size=(A, A)), rng.normal(0, 1, size=(A, A))
    sm = _smoother(W, A, sw, tw)
    Bd = base.predict_proba_f64(torch.from_numpy(X).cuda())
    proba = sm.predict_proba(Bd)
    assert proba.is_cuda
    (_, B_o), _ = util.oracle_lr_fixed(X, coefs, icpts, C, M, ctx, A, want_f64=True)
    p_o, l_o = co.crf_smooth(B_o, sw, tw)
    assert np.array_equal(proba.cpu().numpy().view(np.uint64), p_o.view(np.uint64))
    assert np.array_equal(sm.predict(Bd).cpu().numpy(), l_o)
