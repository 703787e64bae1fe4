"""Small seeded datasets for the parity tests, built with the CPU oracle (test infrastructure)."""
from __future__ import annotations

import functools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from gbnns_dim_red_b200 import synth, xvecs  # noqa: E402

from . import _oracle as O  # noqa: E402


@functools.lru_cache(maxsize=8)
def small_case(n=3000, d=64, n_q=256, d_low=16, dh=64, seed=1, M=12, knn_k=100, latent=8):
    base, queries = synth.make_vectors(n, d, n_q, latent=latent, seed=seed)
    l1, l2, l3 = synth.make_net(d, dh, d_low, seed=seed)
    db_low = O.orc_project(l1, l2, l3, base)
    q_low = O.orc_project(l1, l2, l3, queries)
    knn_ids, _ = O.orc_knn(db_low, db_low, knn_k)
    koff, kedges = xvecs.adjacency_from_matrix(knn_ids)
    goff, gedges = O.orc_gd_prune(koff, kedges, db_low, M=M, reverse=True)
    truth, _ = O.orc_knn(queries, base, 10)
    entry = synth.make_entry_points(n, n_q, seed=seed)
    return dict(base=base, queries=queries, net=(l1, l2, l3), db_low=db_low, q_low=q_low, knn_ids=knn_ids,
                knn=(koff, kedges), graph=(goff, gedges), truth=truth, entry=entry, n=n, d=d, d_low=d_low, n_q=n_q,
                dh=dh, M=M)


def long_link_graph(n, degree=4, seed=5):
    """A sparse random second graph (stand-in for the KL "long link" graph of naive_test.cpp:98-105): `degree`
    distinct random targets per vertex, a few vertices left without any."""
    rng = np.random.default_rng(seed)
    nbrs = rng.integers(0, n, size=(n, degree), dtype=np.uint32)
    deg = np.full(n, degree, np.uint64)
    deg[rng.integers(0, n, size=max(1, n // 50))] = 0
    off = np.zeros(n + 1, np.uint64)
    off[1:] = np.cumsum(deg)
    edges = np.concatenate([nbrs[i, : int(deg[i])] for i in range(n)]).astype(np.uint32)
    return off, edges


def hub_points(n=2400, d=16, seed=3):
    """Clusters with a few very tight cores: the core points are in most neighbour lists of their cluster, so
    their rows fill to 2M during the reverse pass (the order-dependent part of addReverseEdgesForGD)."""
    rng = np.random.default_rng(seed)
    centers = rng.standard_normal((12, d)).astype(np.float32) * 4
    lab = rng.integers(0, 12, size=n)
    r = np.where(rng.random(n) < 0.03, 0.02, 1.0).astype(np.float32)
    return (centers[lab] + r[:, None] * rng.standard_normal((n, d)).astype(np.float32)).astype(np.float32)


def mid_case():
    """A larger seeded case (10 000 x 64 base, 2000 queries, 32-dim projection): enough queries for the north star's
    aggregate bars (ids on >= 99.9 % of queries, recall within 0.1 pt) to be meaningful against the as-shipped reference
    outputs of tests/golden/fast.npz."""
    return small_case(n=10000, d=64, n_q=2000, d_low=32, dh=96, seed=5, M=16, knn_k=100, latent=8)


def exact_rerank_topk(ids_low, queries, base, k):
    """The re-rank of the reference (getRealNearest, search_function.h:105-125) extended to k results: exact squared
    L2 in the original dimension (float64) over the low-dimensional survivors, ascending (dist, id).  PAD ids ignored."""
    out = np.full((ids_low.shape[0], k), 0xFFFFFFFF, np.uint32)
    for i in range(ids_low.shape[0]):
        cand = ids_low[i][ids_low[i] != 0xFFFFFFFF].astype(np.int64)
        diff = base[cand].astype(np.float64) - queries[i].astype(np.float64)
        d2 = (diff * diff).sum(axis=1)
        order = np.lexsort((cand, d2))[:k]
        out[i, : order.size] = cand[order]
    return out
