"""Host half of the packed H2D path (gnx_pack_rows_host, csrc/host_pack.cpp): pure host code,
checked against a numpy bit-plane restatement for every ISA variant and ragged sizes."""
import ctypes as C
import os

import numpy as np
import pytest


def _pack(lib, X, cols, threads=0):
    n, ld = X.shape
    pw = 2 * ((cols + 127) // 128 * 128 // 64)
    out = np.full((n, pw), 0xDEADBEEF, dtype=np.uint64)
    bad = C.c_int(-1)
    assert lib.gnx_pack_rows_host(X.ctypes.data, n, ld, cols, out.ctypes.data, pw, threads, C.byref(bad)) == 0
    return out, bad.value


def _planes(X, cols, pw):
    n, G = X.shape[0], pw // 2
    Z = np.zeros((n, G * 64), dtype=np.uint8)
    Z[:, :cols] = X[:, :cols].view(np.uint8)
    p = [np.packbits(((Z >> k) & 1).reshape(n, G, 64), axis=2, bitorder="little").view(np.uint64).reshape(n, G) for k in (0, 1)]
    return np.stack(p, axis=2).reshape(n, pw)


@pytest.mark.parametrize("isa", ["scalar", "avx2", "avx512"])
def test_pack_rows_matches_numpy_bit_planes(libgnx, isa, monkeypatch):
    monkeypatch.setenv("GNX_HOST_PACK_ISA", isa)   # variants the CPU lacks fall back to the best available one
    rng = np.random.default_rng(3)
    for cols in [1, 63, 64, 65, 127, 128, 129, 1000, 60037]:
        X = rng.integers(0, 3, size=(5, cols + 7), dtype=np.int8)
        out, bad = _pack(libgnx, X, cols)
        assert bad == 0 and np.array_equal(out, _planes(X, cols, out.shape[1])), (isa, cols)
        # the value 3 still packs losslessly; anything else is reported
        X[1, cols // 2] = 3
        out, bad = _pack(libgnx, X, cols)
        assert bad == 0 and np.array_equal(out, _planes(X, cols, out.shape[1]))
        for v in (4, 5, -1, -128):
            Y = X.copy()
            Y[3, cols - 1] = v
            assert _pack(libgnx, Y, cols)[1] == 1, (isa, cols, v)
        # columns beyond C never leak into the planes
        Y = X.copy()
        Y[:, cols:] = 77
        out2, bad = _pack(libgnx, Y, cols)
        assert bad == 0 and np.array_equal(out2, out)


def test_pack_rows_threads_agree_and_empty(libgnx):
    rng = np.random.default_rng(4)
    cols = 300_001
    X = rng.integers(0, 3, size=(37, cols), dtype=np.int8)
    ref, _ = _pack(libgnx, X, cols, threads=1)
    for th in (2, 3, 8, 0):
        out, bad = _pack(libgnx, X, cols, threads=th)
        assert bad == 0 and np.array_equal(out, ref)
    assert libgnx.gnx_host_threads() >= 1
    assert libgnx.gnx_pack_rows_host(None, 0, cols, cols, None, 2 * ((cols + 63) // 64), 1, None) == 0
    # bad arguments: pitch too small
    o = np.zeros(4, dtype=np.uint64)
    assert libgnx.gnx_pack_rows_host(X.ctypes.data, 1, cols, cols, o.ctypes.data, 4, 1, None) != 0


def test_host_threads_env(libgnx, monkeypatch):
    """GNX_HOST_THREADS pins the count; under torchrun the ranks of a node split the cores."""
    monkeypatch.delenv("GNX_HOST_THREADS", raising=False)
    monkeypatch.delenv("LOCAL_WORLD_SIZE", raising=False)
    n = libgnx.gnx_host_threads()
    assert 1 <= n <= 128
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "4")
    assert libgnx.gnx_host_threads() == max(1, n // 4)
    monkeypatch.setenv("GNX_HOST_THREADS", "3")
    assert libgnx.gnx_host_threads() == 3


def test_pack_after_fork(libgnx):
    """The worker pool is rebuilt in a forked child instead of waiting for threads that only exist in the parent."""
    rng = np.random.default_rng(6)
    X = rng.integers(0, 3, size=(64, 70_001), dtype=np.int8)
    ref, _ = _pack(libgnx, X, 70_001, threads=4)          # creates the pool in this process
    pid = os.fork()
    if pid == 0:
        ok = 1
        try:
            out, bad = _pack(libgnx, X, 70_001, threads=4)
            ok = 0 if (bad == 0 and np.array_equal(out, ref)) else 2
        finally:
            os._exit(ok)
    import time
    for _ in range(200):
        done, status = os.waitpid(pid, os.WNOHANG)
        if done:
            break
        time.sleep(0.05)
    else:
        os.kill(pid, 9)
        os.waitpid(pid, 0)
        raise AssertionError("forked child hung in gnx_pack_rows_host")
    assert os.WIFEXITED(status) and os.WEXITSTATUS(status) == 0
