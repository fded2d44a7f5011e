"""Gnofix phase-swap search (reference src/Gnofix/gnofix.py:58-208 with default
arguments, src/Gnofix/phasing.py:182-198) for all individuals on the GPU."""
from __future__ import annotations

import numpy as np

from . import _lib


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def phase_device(smoother, Xd, ld, C, Bd, max_it=50, want_tracker=False):
    """In-place on device tensors: Xd int8 [2n, ld] (or None), Bd float32 [2n, W, A].
    Returns (Y int32 [2n, W], tracker int32 [2n, W] or None)."""
    import torch
    n2, W, A = Bd.shape
    assert n2 % 2 == 0, "gnofix works on haplotype pairs (rows 2i, 2i+1)"
    Y = torch.empty((n2, W), dtype=torch.int32, device=Bd.device)
    trk = torch.empty((n2, W), dtype=torch.int32, device=Bd.device) if want_tracker else None
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(_lib.lib().gnx_gnofix(smoother.model.handle(smoother.S), Xd.data_ptr() if Xd is not None else None, int(ld), int(C),
                                     Bd.data_ptr(), n2 // 2, W, int(max_it), Y.data_ptr(),
                                     trk.data_ptr() if want_tracker else None, st), "gnx_gnofix")
    refused = int((Y[:, 0] < 0).sum().item()) // 2
    if refused:
        raise ValueError("gnofix: the base probabilities of %d individual(s) hold NaN; the rank-form search cannot follow the "
                         "tree nodes' default children (include/gnx.h, gnx_gnofix) -- clean B first" % refused)
    return Y, trk


def phase_device_crf(smoother, Xd, ld, C, Bd, max_it=50, want_tracker=False):
    """Gnofix with the CRF smoother (include/gnx.h gnx_gnofix_crf): AN EXTENSION WITHOUT A REFERENCE ORACLE -- the
    reference refuses the combination (src/model.py:194).  In place on device tensors: Xd int8 [2n, ld] (or None),
    Bd float64 [2n, W, A].  Returns (Y int32 [2n, W], tracker int32 [2n, W] or None)."""
    import torch
    n2, W, A = Bd.shape
    assert n2 % 2 == 0, "gnofix works on haplotype pairs (rows 2i, 2i+1)"
    assert Bd.dtype == torch.float64 and Bd.is_contiguous()
    Y = torch.empty((n2, W), dtype=torch.int32, device=Bd.device)
    trk = torch.empty((n2, W), dtype=torch.int32, device=Bd.device) if want_tracker else None
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(_lib.lib().gnx_gnofix_crf(smoother.model.handle(), int(smoother.S), Xd.data_ptr() if Xd is not None else None, int(ld),
                                         int(C), Bd.data_ptr(), n2 // 2, W, int(max_it), Y.data_ptr(),
                                         trk.data_ptr() if want_tracker else None, st), "gnx_gnofix_crf")
    return Y, trk


def phase_all(model, X, B=None, max_it=50, verbose=False, want_tracker=False):
    """Gnomix.phase (src/model.py:188-214): returns X_phased [N, C] int, Y_phased [N, W] int
    (an odd trailing haplotype is dropped, as the reference's N//2 reshape does).  With a CRF smoother (only reached
    through Gnomix.phase(..., crf_extension=True)) the extension above runs instead, on float64 B."""
    import torch
    from .base import to_device_haplotypes
    from .smooth import CRFModel
    _lib.require_gpu()
    if isinstance(getattr(model.smooth, "model", None), CRFModel):
        on_device = _is_torch(X) and X.is_cuda
        N, Cc = X.shape
        n = N // 2
        Xv, ld = to_device_haplotypes(X[:2 * n])
        if on_device and Xv.data_ptr() == X.data_ptr():
            Xv = Xv.clone()
            ld = Xv.stride(0)
        if B is None:
            Bd = model.base._device_predict(Xv, ld, dtype=torch.float64).contiguous()   # what the CRF smoother reads
        elif _is_torch(B):
            Bd = B[:2 * n].to(device="cuda", dtype=torch.float64).clone().contiguous()
        else:
            Bd = torch.from_numpy(np.ascontiguousarray(np.asarray(B)[:2 * n], dtype=np.float64)).cuda()
        Y, trk = phase_device_crf(model.smooth, Xv, ld, Cc, Bd, max_it=max_it, want_tracker=want_tracker)
        if on_device:
            out = (Xv, Y)
        else:
            torch.cuda.current_stream().synchronize()
            out = (Xv.cpu().numpy().astype(int), Y.cpu().numpy().astype(int))
        if want_tracker:
            return out + ((trk if on_device else trk.cpu().numpy()),)
        return out
    on_device = _is_torch(X) and X.is_cuda
    N, Cc = X.shape
    n = N // 2
    Xv, ld = to_device_haplotypes(X[:2 * n])
    if on_device and Xv.data_ptr() == X.data_ptr():
        Xv = Xv.clone()  # the reference returns fresh arrays and leaves its input alone
        ld = Xv.stride(0)
    if B is None:
        Bd = model.base._device_predict(Xv, ld)
        if Bd.dtype != torch.float32:      # the string-kernel base returns float64; the tree smoother reads float32
            Bd = Bd.to(torch.float32)
    elif _is_torch(B):
        Bd = B[:2 * n].to(device="cuda", dtype=torch.float32).clone()
    else:
        Bd = torch.from_numpy(np.ascontiguousarray(np.asarray(B)[:2 * n], dtype=np.float32)).cuda()
    Y, trk = phase_device(model.smooth, Xv, ld, Cc, Bd, max_it=max_it, want_tracker=want_tracker)
    if on_device:
        out = (Xv, Y)
    else:
        torch.cuda.current_stream().synchronize()
        out = (Xv.cpu().numpy().astype(int), Y.cpu().numpy().astype(int))
    if want_tracker:
        return out + ((trk if on_device else trk.cpu().numpy()),)
    return out
