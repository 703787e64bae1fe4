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
