"""Packed H2D path on the device: unpack(pack(X)) == X, and gnx_infer_host gives the same labels and
probabilities whatever share of each chunk crosses the bus packed (including chunks that cannot be
packed because they hold values outside 0..3)."""
import ctypes as C

import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu


def test_unpack_inverts_pack(libgnx):
    import torch
    from gnomix_b200 import _lib
    _lib.require_gpu()
    rng = np.random.default_rng(8)
    for n, cols in [(1, 1), (3, 64), (7, 129), (33, 60037), (300, 317_408)]:
        ld = (cols + 127) // 128 * 128
        X = rng.integers(0, 4, size=(n, ld), dtype=np.int8)
        X[:, cols:] = 0
        pw = ld // 32
        packed = np.empty((n, pw), dtype=np.uint64)
        bad = C.c_int(0)
        assert libgnx.gnx_pack_rows_host(X.ctypes.data, n, ld, cols, packed.ctypes.data, pw, 0, C.byref(bad)) == 0 and bad.value == 0
        pd = torch.from_numpy(packed.view(np.int64)).cuda()
        Xd = torch.full((n, ld), 9, dtype=torch.int8, device="cuda")
        _lib.check(libgnx.gnx_unpack_dev(pd.data_ptr(), n, pw, cols, Xd.data_ptr(), ld, None), "gnx_unpack_dev")
        torch.cuda.synchronize()
        assert np.array_equal(Xd.cpu().numpy(), X), (n, cols)


def _model(C_, M, A, S, seed):
    from gnomix_b200 import Gnomix, GBTForest
    rng = np.random.default_rng(seed)
    model = Gnomix(C_, M, A, S)
    coefs, icpts, _ = util.random_lr(rng, C_, M, A)
    model.base.set_window_weights(coefs, icpts)
    model.smooth.model = GBTForest.random(rng, A, model.smooth.S, n_rounds=20, depth=4)
    return model, rng


@pytest.mark.parametrize("pinned", [False, True])
def test_infer_host_same_results_for_every_packed_fraction(pinned, monkeypatch):
    import torch
    C_, M, A, S, N = 40_013, 500, 7, 15, 1100
    model, rng = _model(C_, M, A, S, 21)
    X = util.random_haplotypes(rng, N, C_)
    Xh = torch.from_numpy(X).pin_memory() if pinned else X
    monkeypatch.setenv("GNX_HOST_PACK", "0")
    ref_l, ref_p = model.predict_host(Xh, want_proba=True, chunk_haps=256)
    # resident path
    B = model.base.predict_proba(torch.from_numpy(X).cuda())
    P, L = model.smooth._device_smooth(B)
    assert np.array_equal(ref_l, L.cpu().numpy()) and np.array_equal(ref_p, P.cpu().numpy())
    monkeypatch.delenv("GNX_HOST_PACK")
    for frac in [None, "0", "0.37", "1"]:
        if frac is None:
            monkeypatch.delenv("GNX_HOST_PACK_FRAC", raising=False)
        else:
            monkeypatch.setenv("GNX_HOST_PACK_FRAC", frac)
        for chunk in (256, 0):
            l, p = model.predict_host(Xh, want_proba=True, chunk_haps=chunk)
            assert np.array_equal(l, ref_l) and np.array_equal(p, ref_p), (frac, chunk)
    monkeypatch.delenv("GNX_HOST_PACK_FRAC", raising=False)
    # a chunk with an unpackable value goes raw; results still equal the unpacked path on the same input
    X2 = X.copy()
    X2[300, 17] = 5
    X2[900, C_ - 1] = -3
    X2h = torch.from_numpy(X2).pin_memory() if pinned else X2
    monkeypatch.setenv("GNX_HOST_PACK", "0")
    r2 = model.predict_host(X2h, chunk_haps=256)
    monkeypatch.delenv("GNX_HOST_PACK")
    assert np.array_equal(model.predict_host(X2h, chunk_haps=256), r2)
    if pinned:
        pk, h2d = C.c_double(0), C.c_double(0)
        from gnomix_b200 import _lib
        _lib.lib().gnx_infer_host_rates(C.byref(pk), C.byref(h2d))
        assert pk.value > 0 and h2d.value > 0


def test_upload_haplotypes_equals_plain_copy(libgnx, monkeypatch):
    """gnx_upload_haplotypes (what Base.predict_proba uses for a numpy matrix): the device matrix equals the host
    matrix for pageable / pinned input, ragged shapes, strided rows, unpackable values, packing on or off."""
    import torch
    from gnomix_b200 import _lib
    from gnomix_b200.base import to_device_haplotypes
    _lib.require_gpu()
    rng = np.random.default_rng(31)
    for N, Cc in [(1, 1), (5, 127), (300, 129), (1000, 40_013), (2600, 9_001)]:
        X = rng.integers(0, 3, size=(N, Cc + 5), dtype=np.int8)[:, :Cc]      # row stride != C
        X = np.ascontiguousarray(X) if N == 5 else X
        for pinned in (False, True):
            src = torch.from_numpy(np.ascontiguousarray(X)).pin_memory() if pinned else X
            for frac in (None, "0.4", "0"):
                if frac is None:
                    monkeypatch.delenv("GNX_HOST_PACK_FRAC", raising=False)
                else:
                    monkeypatch.setenv("GNX_HOST_PACK_FRAC", frac)
                Xd, ld = to_device_haplotypes(src)
                assert ld % 128 == 0 and np.array_equal(Xd.cpu().numpy(), X), (N, Cc, pinned, frac)
        monkeypatch.delenv("GNX_HOST_PACK_FRAC", raising=False)
        Y = X.copy()
        Y[N // 2, Cc // 2] = -7
        Yd, _ = to_device_haplotypes(Y)
        assert np.array_equal(Yd.cpu().numpy(), Y)
        monkeypatch.setenv("GNX_HOST_PACK", "0")
        Yd, _ = to_device_haplotypes(Y)
        assert np.array_equal(Yd.cpu().numpy(), Y)
        monkeypatch.delenv("GNX_HOST_PACK")
        assert libgnx.gnx_release_workspace() == 0          # buffers come back on demand


def test_infer_host_ships_raw_when_the_cores_are_scarcer_than_the_bus(monkeypatch):
    """One host thread packs slower than the bus copies (the situation of many ranks sharing one host): the pipeline
    must then ship every row raw instead of packing a fraction (a packing rank takes memory bandwidth from the other
    ranks' copies) -- same labels, packed fraction 0."""
    import torch
    from gnomix_b200 import _lib
    C_, M, A, S, N = 400_013, 2000, 7, 15, 768
    model, rng = _model(C_, M, A, S, 5)
    X = util.random_haplotypes(rng, N, C_)
    Xh = torch.from_numpy(X).pin_memory()
    monkeypatch.setenv("GNX_HOST_PACK", "0")
    ref = model.predict_host(Xh, chunk_haps=256)
    monkeypatch.delenv("GNX_HOST_PACK")
    monkeypatch.setenv("GNX_HOST_THREADS", "1")
    _lib.check(_lib.lib().gnx_release_workspace())          # forget the calibration of earlier tests
    got = model.predict_host(Xh, chunk_haps=256)
    pk, h2d = C.c_double(0), C.c_double(0)
    _lib.lib().gnx_infer_host_rates(C.byref(pk), C.byref(h2d))
    frac, a, b = C.c_double(-1), C.c_int64(0), C.c_int64(0)
    _lib.lib().gnx_infer_host_last_transfer(C.byref(frac), C.byref(a), C.byref(b))
    monkeypatch.delenv("GNX_HOST_THREADS")
    _lib.check(_lib.lib().gnx_release_workspace())
    assert np.array_equal(got, ref)
    assert 0 < pk.value < h2d.value, (pk.value, h2d.value)   # one thread: ~8 GB/s against ~50 GB/s
    assert frac.value == 0.0
