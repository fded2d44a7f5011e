"""reference: src/Gnofix/gnofix.py:58-208 -- the per-individual entry point, default arguments only
(the ones Gnomix.phase uses, src/model.py:208), evaluated on the GPU."""
import numpy as np


def gnofix(M, P, B, smoother, max_it=50, non_lin_s=0, check_criterion="disc_smooth", max_center_offset=0, prob_comp="max", d=None,
           prior_switch_prob=0.5, naive_switch=None, end_naive_switch=None, padding=True, verbose=False):
    assert (non_lin_s == 0 and check_criterion == "disc_smooth" and max_center_offset == 0 and prob_comp == "max" and d is None
            and prior_switch_prob == 0.5 and not naive_switch and not end_naive_switch and padding), \
        "gnomix_b200 implements gnofix with the reference's default arguments only"
    import torch
    from gnomix_b200.base import to_device_haplotypes
    from gnomix_b200.gnofix import phase_device
    X = np.stack([np.asarray(M), np.asarray(P)]).astype(np.int8)
    Xd, ld = to_device_haplotypes(X)
    Bd = torch.from_numpy(np.ascontiguousarray(np.asarray(B), dtype=np.float32)).cuda()
    Y, trk = phase_device(smoother, Xd, ld, X.shape[1], Bd, max_it=max_it, want_tracker=True)
    torch.cuda.current_stream().synchronize()
    Xo, Yo, T = Xd.cpu().numpy().astype(int), Y.cpu().numpy().astype(int), trk.cpu().numpy().astype(int)
    history = np.array([Yo[0], Yo[1]])
    return Xo[0], Xo[1], Yo[0], Yo[1], history, (T[0], T[1])
