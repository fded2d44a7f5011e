"""Seeded synthetic workloads of the shapes BASELINE.json names.

Founders follow a Balding-Nichols model (A populations, F_st = 0.1); admixed query
haplotypes are mosaics of founder haplotypes with crossover breakpoints, the way the
reference's simulator builds them (src/laidataset.py:119-176: #crossovers ~
Poisson(generations * morgans), breakpoints uniform in genetic distance, every
segment copies one founder haplotype); 1 % of the query SNPs are set to 2 (missing,
src/utils.py:150-153).  The per-window logistic weights are a linear discriminant on
the population allele frequencies -- random-init-grade weights of the reference's
architecture with a meaningful signal, not a trained model.

Geometry of the BASELINE configs.
"""
from __future__ import annotations

import numpy as np

GEOMETRY = {
    # name: (C, M, A, S, morgans)
    "chr22_demo": (317_408, 857, 7, 75, 0.741096),
    "chr22_m1000": (317_408, 1000, 7, 75, 0.741096),
    "chr1": (1_226_139, 857, 7, 75, 2.86279),
}
GENERATIONS = (2, 4, 6, 8, 12, 16, 24)  # config.yaml:12


def population_frequencies(rng, C, A, fst=0.1):
    """[A, C] float32 allele frequencies (Balding-Nichols around Beta(0.5, 0.5) ancestral)."""
    p = np.clip(rng.beta(0.5, 0.5, size=C), 0.01, 0.99)
    a = p * (1 - fst) / fst
    b = (1 - p) * (1 - fst) / fst
    return rng.beta(a[None, :].repeat(A, 0), b[None, :].repeat(A, 0)).astype(np.float32)


def founders(rng, freqs, per_pop):
    """([A*per_pop, C] int8 founder haplotypes, [A*per_pop] population of each)."""
    A, C = freqs.shape
    out = np.empty((A * per_pop, C), dtype=np.int8)
    for a in range(A):
        out[a * per_pop:(a + 1) * per_pop] = rng.random((per_pop, C), dtype=np.float32) < freqs[a][None, :]
    return out, np.repeat(np.arange(A), per_pop)


def discriminant_lr_weights(freqs, C, M, ctx, gain=0.08, dtype=np.float64):
    """Per-window (coef [A, M_w], intercept [A]) in the reference's padded-window
    feature order (src/Base/base.py:41-44,157-164)."""
    A = freqs.shape[0]
    W = C // M
    rem = C - M * W
    M_ = M + 2 * ctx
    d = (freqs - freqs.mean(axis=0, keepdims=True)).astype(dtype)          # [A, C]
    mean_x = freqs.mean(axis=0).astype(dtype)
    dp = np.concatenate([d[:, :ctx][:, ::-1], d, d[:, C - ctx:][:, ::-1]], axis=1) if ctx else d
    mp = np.concatenate([mean_x[:ctx][::-1], mean_x, mean_x[C - ctx:][::-1]]) if ctx else mean_x
    coefs, icpts = [], []
    for w in range(W):
        lo, hi = (w * M, w * M + M_) if w < W - 1 else (C + 2 * ctx - (M_ + rem), C + 2 * ctx)
        cf = np.ascontiguousarray(dp[:, lo:hi]) * gain
        coefs.append(cf)
        icpts.append(-(cf @ mp[lo:hi]) - 1.0)
    return coefs, icpts


def admix_host(rng, founders_x, founders_pop, n, morgans, missing=0.01):
    """CPU twin of `admix_device` for small n: (X [n, C] int8, y [n, C] int8 ancestry)."""
    F, C = founders_x.shape
    X = np.empty((n, C), dtype=np.int8)
    y = np.empty((n, C), dtype=np.int8)
    for i in range(n):
        gen = GENERATIONS[rng.integers(len(GENERATIONS))]
        k = rng.poisson(gen * morgans)
        brk = np.sort(rng.integers(1, C, size=k)) if k else np.empty(0, dtype=np.int64)
        edges = np.concatenate([[0], brk, [C]])
        for s in range(len(edges) - 1):
            f = rng.integers(F)
            X[i, edges[s]:edges[s + 1]] = founders_x[f, edges[s]:edges[s + 1]]
            y[i, edges[s]:edges[s + 1]] = founders_pop[f]
        if missing > 0:
            X[i, rng.random(C, dtype=np.float32) < missing] = 2
    return X, y


def admix_device(founders_dev, n, morgans, seed, ld=None, missing=0.01, chunk=64, out=None):
    """n admixed haplotypes generated on the GPU: int8 [n, ld] (ld = C rounded up to 128
    by default; columns >= C are zero).  Deterministic for a given (seed, chunk)."""
    import torch
    dev = founders_dev.device
    F, C = founders_dev.shape
    ld = ld or (C + 127) // 128 * 128
    X = out if out is not None else torch.zeros((n, ld), dtype=torch.int8, device=dev)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    gens = torch.tensor(GENERATIONS, device=dev, dtype=torch.float32)
    kmax = 160
    cols = torch.arange(C, device=dev)
    for n0 in range(0, n, chunk):
        m = min(chunk, n - n0)
        gen = gens[torch.randint(len(GENERATIONS), (m,), device=dev, generator=g)]
        k = torch.poisson(gen * morgans, generator=g).clamp_(max=kmax).long()          # [m]
        brk = torch.randint(1, C, (m, kmax), device=dev, generator=g)
        brk = torch.where(torch.arange(kmax, device=dev)[None, :] < k[:, None], brk, torch.full_like(brk, C))
        ind = torch.zeros((m, C + 1), dtype=torch.int16, device=dev)
        ind.scatter_add_(1, brk, torch.ones_like(brk, dtype=torch.int16))
        seg = torch.cumsum(ind[:, :C], dim=1, dtype=torch.int32).long()                 # [m, C] segment id
        seg_founder = torch.randint(F, (m, kmax + 1), device=dev, generator=g)
        fidx = torch.gather(seg_founder, 1, seg)                                          # [m, C]
        x = founders_dev[fidx, cols[None, :]]
        if missing > 0:
            x = torch.where(torch.rand((m, C), device=dev, generator=g) < missing, torch.full_like(x, 2), x)
        X[n0:n0 + m, :C] = x
        del ind, seg, fidx, x
    return X
