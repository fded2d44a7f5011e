"""Parity at BASELINE.json's geometries through size-independent properties: the two
independent implementations of each stage agree bit for bit (tcgen05 vs dp4a for K1,
rank-form vs float traversal for K4), results do not depend on how haplotypes are batched or
sharded, the host-buffer pipeline equals the resident path, and a CPU-oracle spot check."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup(workload, N, seed):
    import torch
    import bench
    from gnomix_b200 import synth
    bench.WORKLOAD = workload
    geom = synth.GEOMETRY[workload]
    base, smooth, (fx, fpop), (coefs, icpts, ctx), _ = bench.build_models(geom)
    X = synth.admix_device(torch.from_numpy(fx).cuda(), N, geom[4], seed=seed)
    return geom, base, smooth, X, (coefs, icpts, ctx)


@pytest.mark.parametrize("workload,N", [("chr22_m1000", 3000), ("chr1", 700)])
def test_two_implementations_agree_and_batching_invariant(workload, N):
    import torch
    geom, base, smooth, X, _ = _setup(workload, N, seed=11)
    C, M, A, S, _ = geom
    Xv = X[:, :C]
    base.kernel = 0
    B_tc = base.predict_proba(Xv)
    base.kernel = 1
    B_dp = base.predict_proba(Xv)
    base.kernel = 0
    assert torch.equal(B_tc, B_dp), "tcgen05 and dp4a logistic kernels differ"
    smooth.model.kernel = 0
    P0, L0 = smooth._device_smooth(B_tc)
    smooth.model.kernel = 1
    P1, L1 = smooth._device_smooth(B_tc)
    smooth.model.kernel = 0
    assert torch.equal(P0, P1) and torch.equal(L0, L1), "rank-form and float-traversal smoothers differ"
    assert torch.equal(L0.long(), P0.argmax(-1))
    # sharding / batching invariance: any sub-range gives the same rows
    lo, hi = N // 3 + 1, N // 3 + 1 + 257
    Bs = base.predict_proba(X[lo:hi, :C])
    assert torch.equal(Bs, B_tc[lo:hi])
    Ps, Ls = smooth._device_smooth(Bs)
    assert torch.equal(Ps, P0[lo:hi]) and torch.equal(Ls, L0[lo:hi])
    # rows are proper distributions
    assert float((B_tc.sum(-1) - 1).abs().max()) < 1e-6 and float((P0.sum(-1) - 1).abs().max()) < 1e-5


def test_host_pipeline_equals_resident_path_and_oracle_spot_check():
    import torch
    from gnomix_b200 import Gnomix
    from oracle import np_oracle as npo, c_oracle as co
    geom, base, smooth, X, (coefs, icpts, ctx) = _setup("chr22_m1000", 1500, seed=5)
    C, M, A, S, _ = geom
    model = Gnomix.__new__(Gnomix)
    model.C, model.M, model.A, model.S, model.W = C, M, A, S, C // M
    model.base, model.smooth = base, smooth
    Xh = X[:, :C].cpu().numpy()
    labels, proba = model.predict_host(Xh, want_proba=True, chunk_haps=512)   # 3 chunks, ragged tail
    B = base.predict_proba(X[:, :C])
    P, L = smooth._device_smooth(B)
    assert np.array_equal(labels, L.cpu().numpy()) and np.array_equal(proba, P.cpu().numpy())
    # CPU oracle on a sample of the same haplotypes: bit-exact labels and probabilities
    n = 24
    s = npo.lr_choose_scale(coefs, C, M, ctx, 7)
    B_o = co.lr_fixed(Xh[:n], npo.lr_quantize_fold(coefs, C, M, ctx, s), np.stack(icpts), C, M, ctx, A, s)
    assert np.array_equal(B_o, B[:n].cpu().numpy())
    p_o, l_o = co.gbt_smooth(smooth.model, B_o, S)
    assert np.array_equal(l_o, labels[:n]) and np.array_equal(p_o, proba[:n])
    # and within 1e-5 of the float64 restatement of the reference's CPU path (north-star tolerance)
    B64 = npo.lr_base_predict_proba(Xh[:n], coefs, icpts, C, M, ctx)
    assert np.max(np.abs(B64 - B_o)) < 1e-6
    p64, l64 = co.gbt_smooth(smooth.model, B64.astype(np.float32), S)
    assert np.array_equal(l64, l_o) and np.max(np.abs(p64 - p_o)) < 1e-5
