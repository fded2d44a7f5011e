"""The exact device math helpers (include/gnx_math.h) equal the builtin IEEE operations."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_device_math_helpers_match_ieee(libgnx):
    from gnomix_b200 import _lib
    bad = np.zeros(5, dtype=np.int64)
    for seed in (1, 2, 3):
        _lib.check(libgnx.gnx_selftest_math(200_000_000, seed, bad.ctypes.data_as(C.c_void_p)), "gnx_selftest_math")
        assert bad.tolist() == [0, 0, 0, 0, 0], "mismatches [div, ll2d, d2f, f2d, exp] = %s" % bad.tolist()
