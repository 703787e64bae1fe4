"""Seeded synthetic inputs of the BASELINE.json shapes.

There are no datasets and no network here, so base/query vectors follow a low-intrinsic-dimension
law (``base = Z·A + noise·E``) and the projection net is the reference architecture
(Linear-BN-ReLU-Linear-BN-ReLU-Linear-Normalize, dim_red/triplet.py:203-212) at random init,
exported in the reference's ``[out][in+1]`` matrix format (dim_red/support_func.py:517-555).
With PyTorch's default BatchNorm state (gamma=1, beta=0, mean=0, var=1) the reference's BN fold
(prepare_net_layer, support_func.py:528-548: W*gamma/sqrt(var), (b-mean)*gamma/sqrt(var)+beta) is
the identity, so the exported matrices are just the Linear layers' default init
U(-1/sqrt(in), 1/sqrt(in)) for weights and biases.
"""
from __future__ import annotations

import numpy as np

SHAPES = {
    # name: n, d, n_q, d_low, d_hidden            (BASELINE.json configs / SURVEY §8)
    "c1": dict(n=100_000, d=128, n_q=1_000, d_low=32, d_hidden=256),
    "sift1m": dict(n=1_000_000, d=128, n_q=10_000, d_low=32, d_hidden=256),
    "deep1m": dict(n=1_000_000, d=96, n_q=10_000, d_low=16, d_hidden=128),
    "gist1m": dict(n=1_000_000, d=960, n_q=1_000, d_low=32, d_hidden=1024),
}


def make_vectors(n, d, n_q, latent=8, noise=0.1, seed=1234):
    """base [n x d], queries [n_q x d] float32 from the same law."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((latent, d), dtype=np.float32)

    def draw(m):
        out = np.empty((m, d), dtype=np.float32)
        step = 1 << 18
        for i in range(0, m, step):
            j = min(m, i + step)
            z = rng.standard_normal((j - i, latent), dtype=np.float32)
            e = rng.standard_normal((j - i, d), dtype=np.float32)
            out[i:j] = z @ A + np.float32(noise) * e
        return out

    base = draw(n)
    queries = draw(n_q)
    return base, queries


def make_part(n, d, part, latent=8, noise=0.1, seed=1234):
    """Rows of one part (a database shard, or the query set) of a sharded synthetic dataset: every part is
    drawn from the same law (latent map from `seed`) with its own stream (`seed`, `part`), so ranks can
    generate their shards independently."""
    A = np.random.default_rng(seed).standard_normal((latent, d), dtype=np.float32)
    rng = np.random.default_rng([seed, 7919 + int(part)])
    out = np.empty((n, d), dtype=np.float32)
    step = 1 << 18
    for i in range(0, n, step):
        j = min(n, i + step)
        z = rng.standard_normal((j - i, latent), dtype=np.float32)
        e = rng.standard_normal((j - i, d), dtype=np.float32)
        out[i:j] = z @ A + np.float32(noise) * e
    return out


def make_net(d, d_hidden, d_low, seed=1234, d_hidden2=None):
    """Three matrices in the reference layout: l1 [dh x (d+1)], l2 [dh2 x (dh+1)], l3 [d_low x (dh2+1)]."""
    d_hidden2 = d_hidden2 or d_hidden
    rng = np.random.default_rng(seed + 7)

    def layer(o, i):
        b = 1.0 / np.sqrt(i)
        w = rng.uniform(-b, b, size=(o, i)).astype(np.float32)
        bias = rng.uniform(-b, b, size=(o, 1)).astype(np.float32)
        return np.ascontiguousarray(np.concatenate([w, bias], axis=1))

    return layer(d_hidden, d), layer(d_hidden2, d_hidden), layer(d_low, d_hidden2)


def project_numpy(l1, l2, l3, x):
    """float64 evaluation of the exported net (tolerance reference for the tensor-core path)."""
    x = np.asarray(x, dtype=np.float64)

    def lin(m, v):
        return v @ m[:, :-1].astype(np.float64).T + m[:, -1].astype(np.float64)

    h = np.maximum(lin(l1, x), 0)
    h = np.maximum(lin(l2, h), 0)
    y = lin(l3, h)
    d_low = y.shape[1]
    nrm = np.sqrt((y[:, : (d_low // 4) * 4] ** 2).sum(1, keepdims=True))  # normalizeVector ignores d%4 tail
    return y / nrm


def make_entry_points(n, n_q, seed=1234):
    """One uniform random entry vertex per query (performRealTests, search_function.h:297-307;
    injected identically into both sides)."""
    rng = np.random.default_rng(seed + 13)
    return rng.integers(0, n, size=n_q, dtype=np.uint32)
