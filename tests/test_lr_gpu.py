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
    (3001, 200, 12, 130),     # A > 8 -> 16-wide limb groups
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
    limbs = 7 if A <= 8 else 4
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
    tol = 1e-12 if limbs == 7 else 1e-5
    assert np.max(np.abs(Bd - B64)) < tol


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
