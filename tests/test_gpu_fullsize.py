"""Parity at BASELINE.json's full size (configs[1]: SIFT-1M shape, 1M x 128 base, 10k queries, d_low 32).

The CPU oracle cannot redo a whole 1M-vertex build in test time, so the checks are (a) bit-exact comparison
with the oracle on SAMPLES that are cheap for it (a few hundred queries over the full 1M-vertex graph, a few
rows of the kNN self-join, the forward prune of a few vertices) and (b) size-independent properties over the
whole output (sortedness, self at rank 0, degree bounds, no duplicate edges, run-to-run identity, recall).
The workload is the bench's own (gbnns_dim_red_b200.workload, built on the GPU through the C ABI, cached)."""
import os

import numpy as np
import pytest

from gbnns_dim_red_b200 import capi, workload, xvecs

from . import _oracle as O

pytestmark = pytest.mark.gpu

CACHE = os.environ.get("GBDR_BENCH_CACHE", "/tmp/gbdr_bench_cache")
EF = 53  # the bench's operating point on this workload (recall@1 = 0.95)


@pytest.fixture(scope="module")
def w():
    return workload.build_workload("sift1m", device=0, cache_dir=CACHE)


@pytest.fixture(scope="module")
def index(w):
    ix = capi.Index(0)
    ix.set_base(w["base"])
    ix.set_low(w["db_low"])
    ix.set_graph(*w["graph"])
    ix.set_net(*w["net"])
    yield ix
    ix.close()


def test_search_sample_is_bit_exact_at_1m(w, index):
    rng = np.random.default_rng(11)
    pick = np.sort(rng.choice(w["shape"]["n_q"], size=256, replace=False))
    q = np.ascontiguousarray(w["queries"][pick])
    entry = np.ascontiguousarray(w["entry"][pick])
    q_low = O.orc_project(*w["net"], q)
    goff, ged = w["graph"]
    for mode, flags, ef, k in ((0, capi.SEARCH_RERANK, EF, 1), (0, capi.SEARCH_RERANK, 120, 10), (1, 0, 100, 10),
                               (2, capi.SEARCH_PLAIN, 20, 5)):
        o = O.orc_search(q, q_low, w["base"], w["db_low"], goff, ged, ef, k, mode, entry)
        g = index.search(q, q_low, ef, k, entry, flags=flags)
        for key in ("ids", "dists", "hops", "dist_calc"):
            assert np.array_equal(g[key], o[key]), (mode, ef, key)


def test_whole_batch_properties_at_1m(w, index):
    n, n_q = w["shape"]["n"], w["shape"]["n_q"]
    a = index.search(w["queries"], None, EF, 1, w["entry"], flags=capi.SEARCH_RERANK)
    assert a["ids"].max() < n
    assert workload.recall_at_1(a["ids"], w["truth"], w["base"]) >= 0.945
    assert (a["dist_calc"] >= a["hops"] + EF).all() and (a["hops"] >= 1).all()
    # identical on a second run and on a view with other batches in flight (no run-to-run state)
    b = index.search(w["queries"], None, EF, 1, w["entry"], flags=capi.SEARCH_RERANK)
    v = index.view()
    qp = capi.pinned_empty(w["queries"].shape, np.float32)
    qp[:] = w["queries"]
    ep = capi.pinned_empty((n_q,), np.uint32)
    ep[:] = w["entry"]
    index.search_submit(qp, None, EF, 1, ep, flags=capi.SEARCH_RERANK)
    v.search_submit(qp, None, EF, 1, ep, flags=capi.SEARCH_RERANK)
    c, d = index.search_wait(), v.search_wait()
    v.close()
    for other in (b, c, d):
        for key in ("ids", "dists", "hops", "dist_calc"):
            assert np.array_equal(a[key], other[key]), key
    # top-k lists come out ascending, and the top-1 of a top-10 call is the top-1 call's answer
    t = index.search(w["queries"], None, EF, 10, w["entry"], flags=capi.SEARCH_RERANK)
    assert (np.diff(t["dists"], axis=1) >= 0).all()
    assert np.array_equal(t["ids"][:, 0], a["ids"][:, 0])
    # exact distances: recomputed on the host in float64 for every answer
    diff = w["base"][a["ids"][:, 0]].astype(np.float64) - w["queries"].astype(np.float64)
    assert np.allclose((diff * diff).sum(axis=1), a["dists"][:, 0], rtol=1e-5)


def test_knn_build_and_gd_prune_at_1m(w):
    n, k, M = w["shape"]["n"], 1000, 30
    db_low = w["db_low"]
    pinned = capi.PinnedArray((n, k), np.uint32)
    try:
        ids, _ = capi.knn(db_low, db_low, k, out_ids=pinned.array)
        # self at rank 0 (distance 0 is the minimum; a duplicate vector with a smaller id would come first)
        self_first = ids[:, 0] == np.arange(n, dtype=np.uint32)
        assert self_first.mean() > 0.9999
        rng = np.random.default_rng(5)
        rows = np.sort(rng.choice(n, size=24, replace=False))
        oi, _ = O.orc_knn(db_low[rows], db_low, k)
        assert np.array_equal(ids[rows], oi)
        # rows are duplicate-free and in range (checked on a slice: np.sort of 1e9 ids is too slow for a test)
        sl = ids[:: n // 2000]
        assert sl.max() < n
        assert (np.diff(np.sort(sl, axis=1), axis=1) > 0).all()

        # hnswlikeGD: forward prune of sampled vertices against the oracle (candidate lists emptied elsewhere,
        # reverse off), then the full graph: the sampled rows start with exactly those forward lists
        deg = np.zeros(n, np.uint64)
        deg[rows] = k
        soff = np.zeros(n + 1, np.uint64)
        soff[1:] = np.cumsum(deg)
        sed = np.ascontiguousarray(ids[rows]).reshape(-1)
        ooff, oed = O.orc_gd_prune(soff, sed, db_low, M=M, reverse=False)
        koff, ked = xvecs.adjacency_from_matrix(ids)
        goff, ged, _ = capi.gd_prune(koff, ked, db_low, M=M, reverse=True)
    finally:
        pinned.close()
    assert np.array_equal(goff, w["graph"][0]) and np.array_equal(ged, w["graph"][1])  # same as the cached build
    gdeg = np.diff(goff.astype(np.int64))
    assert gdeg.max() <= 2 * M and gdeg.min() >= 1
    for r in rows:
        fwd = oed[int(ooff[r]):int(ooff[r + 1])]
        row = ged[int(goff[r]):int(goff[r + 1])]
        assert len(fwd) >= 1 and np.array_equal(row[: len(fwd)], fwd), r
        assert len(set(row.tolist())) == len(row) and r not in row
    # every edge target in range, no vertex lists itself (checked on all edges)
    assert ged.max() < n
    src = np.repeat(np.arange(n, dtype=np.uint32), gdeg)
    assert not (src == ged).any()
