"""Edge geometries through the C ABI: no context, one or two windows, a single haplotype, wide
windows, ancestry counts at the limb-group boundaries, context ratios other than 0.5."""
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("C,M,A,N,ctx_ratio", [
    (1000, 300, 7, 5, 0.0),       # no context: windows are disjoint, last one takes the remainder
    (700, 500, 3, 9, 0.5),        # W = 1: the only window is also the last (remainder + both pads)
    (1100, 500, 4, 130, 0.5),     # W = 2
    (9000, 4200, 7, 1, 0.5),      # single haplotype, windows of 8400+ SNPs
    (6000, 500, 8, 70, 0.5),      # A = 8: a full 8-column limb group
    (6000, 500, 9, 70, 0.5),      # A = 9: split model (8 + 1 classes), 7 limbs
    (5000, 400, 16, 33, 0.5),     # A = 16
    (8000, 640, 7, 200, 0.25),    # context ratio 0.25
    (8000, 640, 7, 200, 1.0),     # context as wide as the window (5 windows live per chunk -> dp4a path)
    (4097, 128, 7, 64, 0.5),      # M = one chunk
])
def test_lr_edge_geometries(C, M, A, N, ctx_ratio):
    import torch
    from oracle import c_oracle as co
    rng = np.random.default_rng(C + M + A)
    coefs, icpts, ctx = util.random_lr(rng, C, M, A, ctx_ratio=ctx_ratio)
    X = util.random_haplotypes(rng, N, C)
    base = util.make_lr_base(C, M, A, coefs, icpts, ctx_ratio=ctx_ratio)
    limbs = 7
    (Bf_o, Bd_o), s = util.oracle_lr_fixed(X, coefs, icpts, C, M, ctx, A, limbs=limbs, want_f64=True)
    Xd = torch.from_numpy(X).cuda()
    for kernel in (0, 1):
        base.kernel = kernel
        Bf = base.predict_proba(Xd).cpu().numpy()
        assert np.array_equal(Bf.view(np.uint32), Bf_o.view(np.uint32)), "kernel %d" % kernel
    B64 = co.lr_f64(X, coefs, np.stack(icpts), C, M, ctx, A)
    assert np.max(np.abs(Bd_o - B64)) < 1e-12
    Bh = base.predict_proba(X)                       # numpy in -> float64 out
    assert Bh.dtype == np.float64 and np.array_equal(Bh.view(np.uint64), Bd_o.view(np.uint64))


@pytest.mark.parametrize("W,A,S,N", [(6, 3, 3, 4), (150, 7, 75, 1), (10, 2, 5, 2050), (33, 16, 15, 7)])
def test_smoother_edge_shapes(W, A, S, N):
    from gnomix_b200 import GBTForest
    from gnomix_b200.smooth import XGB_Smoother, CRF_Smoother, CRFModel
    from oracle import c_oracle as co
    rng = np.random.default_rng(W * 7 + A)
    forest = GBTForest.random(rng, A, S, n_rounds=15, depth=4)
    B = util.smooth_B(rng, N, W, A)
    sm = XGB_Smoother(n_windows=W, num_ancestry=A, smooth_window_size=S)
    sm.model = forest
    p_o, l_o = co.gbt_smooth(forest, B, S)
    assert np.array_equal(sm.predict_proba(B).view(np.uint32), p_o.view(np.uint32)) and np.array_equal(sm.predict(B), l_o)
    crf = CRF_Smoother(n_windows=W, num_ancestry=A, smooth_window_size=S)
    crf.model = CRFModel(rng.normal(0, 1, (A, A)), rng.normal(0, 1, (A, A)))
    pc_o, lc_o = co.crf_smooth(B.astype(np.float64), crf.model.state_w, crf.model.trans_w)
    assert np.array_equal(crf.predict_proba(B.astype(np.float64)).view(np.uint64), pc_o.view(np.uint64))
    assert np.array_equal(crf.predict(B.astype(np.float64)), lc_o)


def test_error_paths_report_instead_of_crashing():
    import ctypes as C
    import torch
    from gnomix_b200 import _lib, GBTForest
    lib = _lib.lib()
    rng = np.random.default_rng(0)
    forest = GBTForest.random(rng, 3, 5, n_rounds=2, depth=4)
    h = forest.handle(5)
    B = torch.zeros((2, 4, 3), device="cuda")            # W=4 < reflect pad 3? pad = 3 -> allowed; W=2 is not
    assert lib.gnx_gbt_smooth(h, B.data_ptr(), 2, 2, B.data_ptr(), None, None) != 0 and b"reflect pad" in lib.gnx_last_error()
    assert lib.gnx_gbt_smooth(None, B.data_ptr(), 2, 4, B.data_ptr(), None, None) != 0
    out = C.c_void_p()
    bad = np.array([0, -1, -1], dtype=np.int32)
    z = np.zeros(3, dtype=np.float32)
    offs = np.array([0, 3], dtype=np.int32)
    # malformed tree: child index out of range
    lft = np.array([7, 0, 0], dtype=np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert lib.gnx_gbt_model_create(C.byref(out), 3, 5, 3, p(bad), p(z), p(lft), p(lft), None, p(z), p(np.array([0, 3, 3, 3], dtype=np.int32)), p(z)) != 0
    assert lib.gnx_crf_model_create(C.byref(out), 3, 3, None, None) != 0
    with pytest.raises(AssertionError):
        from gnomix_b200.smooth import XGB_Smoother
        XGB_Smoother(n_windows=10, num_ancestry=3, smooth_window_size=9)   # W < 2S, as the reference asserts
