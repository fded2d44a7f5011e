"""K1 parity: gnx_lr_predict (tcgen05 and dp4a kernels) against the oracle.
Bit-exact against the fixed-point oracle (same integers, same float64 epilogue);
within 1e-12 of the float64 restatement of sklearn's predict_proba."""
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu

CASES = [
    # C, M, A, N
    (5000, 300, 7, 300),      # ragged N, rem=200
    (4099, 256, 3, 64),       # ctx=128 == one chunk
    (20011, 857, 7, 513),     # demo-like window size, N = 2 tiles + 1
    (9000, 1000, 2, 100),     # binary layout [1-p, p]
    (3001, 200, 12, 130),     # A > 8 -> split model: two class groups of <= 8, 7 limbs each, one exponent
    (2900, 180, 16, 70),      # A = 16: two full groups (numpy's 8-lane pairwise sum over 16 terms)
    (2900, 180, 9, 70),       # A = 9: a group of one class
    (2500, 100, 5, 257),      # windows narrower than a chunk (dp4a fallback inside tc path)
]


@pytest.mark.parametrize("kernel", [0, 1])
@pytest.mark.parametrize("C,M,A,N", CASES)
def test_lr_matches_oracle(C, M, A, N, kernel):
    import torch
    from oracle import c_oracle as co
    rng = np.random.default_rng(C * 7 + A)
    coefs, icpts, ctx = util.random_lr(rng, C, M, A)
    X = util.random_haplotypes(rng, N, C)
    base = util.make_lr_base(C, M, A, coefs, icpts)
    base.kernel = kernel
    limbs = 7
    (Bf_o, Bd_o), s = util.oracle_lr_fixed(X, coefs, icpts, C, M, ctx, A, limbs=limbs, want_f64=True)
    assert base.fixed_point_scale() == s
    Xd = torch.from_numpy(X).cuda()
    Bf = base.predict_proba(Xd).cpu().numpy()
    assert Bf.dtype == np.float32 and Bf.shape == (N, C // M, A)
    assert np.array_equal(Bf.view(np.uint32), Bf_o.view(np.uint32)), "float32 B not bit-exact vs fixed-point oracle"
    Bd = base.predict_proba_f64(Xd).cpu().numpy()
    assert np.array_equal(Bd.view(np.uint64), Bd_o.view(np.uint64)), "float64 B not bit-exact vs fixed-point oracle"
    # and the fixed-point path agrees with the float64 restatement of sklearn
    B64 = co.lr_f64(X, coefs, np.stack(icpts), C, M, ctx, A)
    assert np.max(np.abs(Bd - B64)) < 1e-12


def test_lr_numpy_in_numpy_out():
    rng = np.random.default_rng(3)
    C, M, A, N = 3000, 250, 4, 33
    coefs, icpts, ctx = util.random_lr(rng, C, M, A)
    X = util.random_haplotypes(rng, N, C)
    base = util.make_lr_base(C, M, A, coefs, icpts)
    B = base.predict_proba(X)
    assert isinstance(B, np.ndarray) and B.dtype == np.float64 and B.shape == (N, C // M, A)
    (Bf_o, _), _ = util.oracle_lr_fixed(X, coefs, icpts, C, M, ctx, A, want_f64=True)
    assert np.array_equal(B.astype(np.float32), Bf_o)
    assert np.array_equal(base.predict(X), np.argmax(Bf_o, axis=-1))


def test_lr_empty_and_bad_shape():
    import torch
    from gnomix_b200 import _lib
    rng = np.random.default_rng(4)
    C, M, A = 2000, 200, 3
    coefs, icpts, ctx = util.random_lr(rng, C, M, A)
    base = util.make_lr_base(C, M, A, coefs, icpts)
    out = base.predict_proba(torch.empty((0, C), dtype=torch.int8, device="cuda"))
    assert tuple(out.shape) == (0, C // M, A)
    with pytest.raises(AssertionError):
        base.predict_proba(torch.zeros((4, C - 1), dtype=torch.int8, device="cuda"))


@pytest.mark.parametrize("kernel", [0, 1])
@pytest.mark.parametrize("name", ["base_lr_a3.npz", "base_lr_a7.npz", "base_lr_a2.npz"])
def test_lr_against_reference_golden(name, kernel):
    """K1 straight against the reference: real scikit-learn liblinear models trained through the reference's
    Base.train and predicted through its Base.predict_proba (tests/golden/base_lr_a*.npz, oracle/make_golden.py):
    float64 output within 1e-12, float32 output == the reference's float64 rounded to float32 except where the
    reference value sits within 1e-12 of a float32 rounding boundary, labels (argmax) identical."""
    import os
    import torch
    from tests.test_oracle_golden import _load, _split_coefs
    d = _load(name)
    C, M, A, ctx, coefs, icpts = _split_coefs(d)
    X, B_ref = d["X"], d["B"]
    base = util.make_lr_base(C, M, A, coefs, icpts, ctx_ratio=ctx / M)
    assert base.context == ctx
    base.kernel = kernel
    Xd = torch.from_numpy(X).cuda()
    Bd = base.predict_proba_f64(Xd).cpu().numpy()
    assert Bd.shape == B_ref.shape and np.max(np.abs(Bd - B_ref)) < 1e-12
    Bf = base.predict_proba(Xd).cpu().numpy()
    ref32 = B_ref.astype(np.float32)
    differ = Bf.view(np.uint32) != ref32.view(np.uint32)
    # a float32 result may differ from round(reference float64) only by one ulp and only where the two float64
    # values straddle a rounding boundary, i.e. where they differ at all (<= 1e-12)
    assert differ.mean() < 1e-4
    assert np.all(np.abs(Bf[differ].astype(np.float64) - B_ref[differ]) <= np.spacing(ref32[differ]).astype(np.float64))
    assert np.array_equal(np.argmax(Bf, -1), np.argmax(B_ref, -1)) and np.array_equal(np.argmax(Bd, -1), np.argmax(B_ref, -1))
    # the numpy-in / numpy-out plugin call returns the reference's dtype and values
    B_np = base.predict_proba(X)
    assert B_np.dtype == np.float64 and np.max(np.abs(B_np - B_ref)) < 1e-12


@pytest.mark.parametrize("kernel", [0, 1])
def test_lr_wide_model_with_four_limbs_stays_selectable(kernel):
    """A > 8 with limbs=4 asked for explicitly: the single-pass 16-column model (one read of X, |logit error| ~1e-6) --
    bit-exact against the fixed-point oracle at ITS exponent, 1e-5 from the float64 path."""
    import torch
    from oracle import c_oracle as co
    rng = np.random.default_rng(5)
    C, M, A, N = 3001, 200, 12, 40
    coefs, icpts, ctx = util.random_lr(rng, C, M, A)
    X = util.random_haplotypes(rng, N, C)
    base = util.make_lr_base(C, M, A, coefs, icpts)
    base.kernel = kernel
    base.limbs = 4
    (Bf_o, Bd_o), s = util.oracle_lr_fixed(X, coefs, icpts, C, M, ctx, A, limbs=4, want_f64=True)
    assert base.fixed_point_scale() == s
    Xd = torch.from_numpy(X).cuda()
    assert np.array_equal(base.predict_proba(Xd).cpu().numpy().view(np.uint32), Bf_o.view(np.uint32))
    Bd = base.predict_proba_f64(Xd).cpu().numpy()
    assert np.array_equal(Bd.view(np.uint64), Bd_o.view(np.uint64))
    assert np.max(np.abs(Bd - co.lr_f64(X, coefs, np.stack(icpts), C, M, ctx, A))) < 1e-5
