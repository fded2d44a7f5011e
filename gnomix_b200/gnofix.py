"""Gnofix phase-swap search (reference src/Gnofix/gnofix.py:58-208 with default
arguments, src/Gnofix/phasing.py:182-198) for all individuals on the GPU."""
from __future__ import annotations

import numpy as np

from . import _lib


def phase_all(model, X, B=None, max_it=50, verbose=False):
    """Gnomix.phase (src/model.py:188-214): returns X_phased [N, C] int, Y_phased [N, W] int."""
    import torch
    from .base import to_device_haplotypes
    _lib.require_gpu()
    X = np.asarray(X)
    N, Cc = X.shape
    n = N // 2
    W, A = model.W, model.A
    Xd_view, ld = to_device_haplotypes(X[:2 * n])
    Xd_view = Xd_view.clone() if Xd_view.data_ptr() % 16 else Xd_view
    if B is None:
        Bd = model.base._device_predict(Xd_view, ld)
    else:
        Bd = torch.from_numpy(np.ascontiguousarray(np.asarray(B)[:2 * n], dtype=np.float32)).cuda()
    Y = torch.empty((2 * n, W), dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(_lib.lib().gnx_gnofix(model.smooth.model.handle(model.smooth.S), Xd_view.data_ptr(), ld, Cc, Bd.data_ptr(),
                                     n, W, int(max_it), Y.data_ptr(), None, st), "gnx_gnofix")
    torch.cuda.current_stream().synchronize()
    return Xd_view.cpu().numpy().astype(int), Y.cpu().numpy().astype(int)
