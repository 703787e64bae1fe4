"""GPU drop-ins for the kNN helpers of the reference's training code (dim_red/support_func.py).

    from gbnns_dim_red_b200.wrap.support_func import get_nearestneighbors, get_nearestneighbors_partly

Same positional signatures and return conventions as dim_red/support_func.py:20-74 (`get_nearestneighbors` =
faiss IndexFlatL2 when faiss imports, else the torch cdist2 + topk fallback) and :374-384
(`get_nearestneighbors_partly`, which also writes the ivecs file the C++ side reads).  The reference spends most
of every hard-negative-mining epoch here (dim_red/triplet.py:59,148,199,227,268; dim_red/angular.py:77,...).

Differences, all in favour of exactness: distances are the direct-difference fp32 form of the C++ side
(search/support_func.h:107-128) ordered by (dist, id), never the HNSW32 approximation of `needs_exact=False`;
every row of `xq` gets an answer (the torch fallback drops the last len(xq) % 500 rows, :62-64).  Everything
runs on the GPU through gbdr_knn; there is no CPU fallback (`device` is accepted for signature compatibility).
"""
from __future__ import annotations

import os
import time

import numpy as np

from .. import capi, xvecs


def sanitize(x):
    """dim_red/support_func.py:77-78"""
    return np.ascontiguousarray(x, dtype="float32")


def get_nearestneighbors(xq, xb, k, device="cuda", needs_exact=True, verbose=False):
    """ids [len(xq), k] (int64, like faiss' `I`) of the k nearest rows of xb, squared L2, a row of xb equal to the
    query included (rank 0 when xq is xb)."""
    if verbose:
        print("Computing nearest neighbors (gbdr_knn, B200)")
    start = time.time()
    xb = sanitize(xb)
    xq = xb if xq is xb else sanitize(xq)
    ids, _ = capi.knn(xq, xb, int(k), device=int(os.environ.get("GBDR_DEVICE", "0")))
    if verbose:
        print("  NN search (%s) done in %.2f s" % ("cuda", time.time() - start))
    return ids.astype(np.int64)


def get_nearestneighbors_partly(xq, xb, k, device="cuda", bs=10**5, needs_exact=True, path=""):
    """dim_red/support_func.py:374-384: query blocks of `bs` rows, optional ivecs dump of the whole matrix."""
    xb = sanitize(xb)
    same = xq is xb
    xq = xb if same else sanitize(xq)
    knn = [get_nearestneighbors(xq[i0:i0 + bs], xb, k, device, needs_exact) for i0 in range(0, xq.shape[0], bs)]
    out = np.vstack(knn)
    if path != "":
        xvecs.write_ivecs(path, out)
    return out
