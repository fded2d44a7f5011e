"""K4 parity: gnx_gbt_smooth / gnx_gbt_rows against the oracle -- bit-exact float32
probabilities and labels (the kernels and the oracle share include/gnx_math.h)."""
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu


def _smoother(W, A, S, forest):
    from gnomix_b200.smooth import XGB_Smoother
    sm = XGB_Smoother(n_windows=W, num_ancestry=A, smooth_window_size=S)
    sm.model = forest
    return sm


# 0 default (tile kernel for batches, row kernel otherwise), 1 generic float traversal, 10 / 14 the row kernel's two
# walks, 16 the tile kernel forced (gnx.h)
@pytest.mark.parametrize("kernel", [0, 1, 10, 14, 16])
@pytest.mark.parametrize("W,A,S,N,depth", [
    (160, 7, 75, 37, 4),
    (317, 7, 75, 5, 4),
    (60, 3, 25, 64, 3),
    (1430, 7, 75, 3, 4),
    (40, 12, 9, 10, 5),     # A > 8: generic-A kernel
    (64, 2, 31, 9, 2),
])
def test_gbt_smooth_matches_oracle(W, A, S, N, depth, kernel):
    from gnomix_b200 import GBTForest
    from oracle import c_oracle as co
    rng = np.random.default_rng(W * 31 + A)
    forest = GBTForest.random(rng, A, S, n_rounds=20 if W > 1000 else 100, depth=depth)
    forest.kernel = kernel
    B = util.smooth_B(rng, N, W, A)
    if N > 8:
        B[N // 2, W // 3, 0] = np.nan      # NaN follows the default child (slow path inside the fast kernel)
    sm = _smoother(W, A, S, forest)
    proba = sm.predict_proba(B)
    label = sm.predict(B)
    p_o, l_o = co.gbt_smooth(forest, B, S)
    assert proba.dtype == np.float32
    assert np.array_equal(proba.view(np.uint32), p_o.view(np.uint32))
    assert np.array_equal(label, l_o)
    assert np.array_equal(label, np.argmax(np.nan_to_num(proba, nan=-1.0), axis=-1))


def test_gbt_many_windows_are_segmented():
    """W far beyond what one shared-memory row holds (fine window sizes, e.g. 0.05 cM on chr1):
    the rank-form kernel walks the haplotype in segments with an S-1 halo; same bits."""
    from gnomix_b200 import GBTForest
    from oracle import c_oracle as co
    rng = np.random.default_rng(77)
    W, A, S, N = 9001, 7, 75, 3
    forest = GBTForest.random(rng, A, S, n_rounds=10, depth=4)
    B = util.smooth_B(rng, N, W, A)
    B[1, 4000, 2] = np.nan
    sm = _smoother(W, A, S, forest)
    proba, label = sm.predict_proba(B), sm.predict(B)
    p_o, l_o = co.gbt_smooth(forest, B, S)
    assert np.array_equal(proba.view(np.uint32), p_o.view(np.uint32)) and np.array_equal(label, l_o)


def test_gbt_thresholds_hit_exactly():
    """Inputs equal to split thresholds (and one ulp either side) take the same branch as the
    float compare: the rank transform is exact."""
    from gnomix_b200 import GBTForest
    from oracle import c_oracle as co
    rng = np.random.default_rng(2)
    W, A, S, N = 150, 7, 75, 6
    forest = GBTForest.random(rng, A, S, n_rounds=100, depth=4)
    thr = forest.thr[forest.feat >= 0]
    pick = rng.choice(thr, size=(N, W, A)).astype(np.float32)
    jitter = rng.integers(-1, 2, size=pick.shape)
    B = np.where(jitter < 0, np.nextafter(pick, np.float32(-1)), np.where(jitter > 0, np.nextafter(pick, np.float32(2)), pick)).astype(np.float32)
    sm = _smoother(W, A, S, forest)
    proba = sm.predict_proba(B)
    p_o, l_o = co.gbt_smooth(forest, B, S)
    assert np.array_equal(proba.view(np.uint32), p_o.view(np.uint32))
    assert np.array_equal(sm.predict(B), l_o)


def test_gbt_rows_matches_oracle_and_slide_window():
    from gnomix_b200 import GBTForest
    from oracle import c_oracle as co, np_oracle as npo
    rng = np.random.default_rng(11)
    W, A, S, N = 90, 5, 21, 4
    forest = GBTForest.random(rng, A, S, n_rounds=30, depth=4)
    B = util.smooth_B(rng, N, W, A)
    rows = npo.slide_window(B, S)
    got = forest.predict_proba(rows)
    want = co.gbt_rows(forest, rows)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    sm = _smoother(W, A, S, forest)
    assert np.array_equal(sm.predict_proba(B).reshape(-1, A), got)


def test_gbt_ragged_trees_and_nan_default():
    """Unbalanced trees (leaves above the bottom level) and NaN inputs -> default child."""
    from gnomix_b200 import GBTForest
    from oracle import c_oracle as co
    A, S = 3, 5
    F = S * A
    # tree: root split f0<0.5 -> leaf | split f7<0.25 -> (leaf, split f3<0.9 -> leaf, leaf)
    feat = [0, -1, 7, -1, 3, -1, -1]
    thr = [0.5, 0, 0.25, 0, 0.9, 0, 0]
    left = [1, 0, 3, 0, 5, 0, 0]
    right = [2, 0, 4, 0, 6, 0, 0]
    dl = [1, 0, 0, 0, 1, 0, 0]
    leaf = [0, 0.3, 0, -0.2, 0, 0.7, -0.4]
    feats, thrs, lefts, rights, dls, leafs, offs = [], [], [], [], [], [], [0]
    for t in range(6):
        feats += feat; thrs += thr; lefts += left; rights += right; dls += dl
        leafs += [v * (1 + 0.1 * t) for v in leaf]
        offs.append(len(feats))
    forest = GBTForest(A, F, feats, thrs, lefts, rights, dls, leafs, offs, np.full(A, 0.5, np.float32))
    rng = np.random.default_rng(5)
    rows = rng.random((64, F)).astype(np.float32)
    rows[::3, 0] = np.nan
    rows[::5, 3] = np.nan
    got = forest.predict_proba(rows)
    want = co.gbt_rows(forest, rows)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("W,A,S,N", [(203, 7, 25, 70), (200, 3, 31, 33), (151, 7, 75, 40), (65, 5, 9, 64)])
def test_gbt_tile_pair_words_at_segment_seams(W, A, S, N):
    """The tile kernel's pair words (rank(slot) << 17 | rank(slot + 1)): segments of odd length (the last row pair of a
    tile is clamped), the slot BEHIND a tile (it supplies the low half of the tile's last words) holding NaN or an exact
    threshold hit, ranks equal to node thresholds next to large neighbours, a partial last haplotype block."""
    from gnomix_b200 import GBTForest
    from oracle import c_oracle as co
    rng = np.random.default_rng(W * 7 + S)
    forest = GBTForest.random(rng, A, S, n_rounds=60, depth=4)
    forest.kernel = 16
    thr = forest.thr[forest.feat >= 0]
    B = util.smooth_B(rng, N, W, A)
    # half of the values sit exactly on split thresholds (rank == k: the low halves of pair and node word decide)
    hit = rng.random(B.shape) < 0.5
    B = np.where(hit, rng.choice(thr, size=B.shape), B).astype(np.float32)
    # NaN around every place a segment can end, on different haplotypes
    pad = (S + 1) // 2
    for n in range(0, N, 3):
        for seam in range(32, W + S, 16):
            w = seam + (n % 7) - 3 - pad
            if 0 <= w < W:
                B[n, w, n % A] = np.nan
    sm = _smoother(W, A, S, forest)
    proba, label = sm.predict_proba(B), sm.predict(B)
    p_o, l_o = co.gbt_smooth(forest, B, S)
    assert np.array_equal(proba.view(np.uint32), p_o.view(np.uint32))
    assert np.array_equal(label, l_o)
