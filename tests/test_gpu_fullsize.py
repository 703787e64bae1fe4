"""Parity at BASELINE.json's full sizes: configs[1] SIFT-1M shape (1M x 128, 10k queries, d_low 32, GD graph), configs[2]
Deep-1M shape (1M x 96, d_low 16, fixed kNN-32 graph), configs[3] GIST-1M shape (1M x 960, 1000 queries, d_low 32), and
an index above 4 M vertices (the 32-bit visited-slot plan that Deep-100M-sized shards run on).

The CPU oracle cannot redo a whole 1M-vertex build in test time, so the checks are (a) bit-exact comparison with the
oracle on SAMPLES that are cheap for it (a few hundred queries over the full 1M-vertex graph, a few rows of the kNN
self-join, the forward prune of a few vertices), (b) the north star's tolerance bars against the reference AS SHIPPED
(oracle/_ref/libgbdr_ref_fast.so, the README's -Ofast build, run here on the box's host cores over the whole query
set), (c) size-independent properties over the whole output (sortedness, self at rank 0, degree bounds, no duplicate
edges, run-to-run identity, float64 recomputation of every answer and of the ground truth, recall@1 / recall@10).
The workloads are the bench's own (gbnns_dim_red_b200.workload, built on the GPU through the C ABI, cached)."""
import os

import numpy as np
import pytest

from gbnns_dim_red_b200 import capi, workload, xvecs

from . import _oracle as O
from ._data import exact_rerank_topk

pytestmark = pytest.mark.gpu

CACHE = os.environ.get("GBDR_BENCH_CACHE", "/tmp/gbdr_bench_cache")
# the bench's operating points (smallest ef with recall@1 >= 0.95) and a wide beam per workload
EF = {"sift1m": 53, "deep1m": 294, "gist1m": 47}
WIDE = {"sift1m": 120, "deep1m": 500, "gist1m": 160}


@pytest.fixture(scope="module", params=["sift1m", "deep1m", "gist1m"])
def w(request):
    out = workload.build_workload(request.param, device=0, cache_dir=CACHE)
    out["name"] = request.param
    return out


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
    ef, wide = EF[w["name"]], WIDE[w["name"]]
    for mode, flags, e, k in ((0, capi.SEARCH_RERANK, ef, 1), (0, capi.SEARCH_RERANK, wide, 10), (1, 0, 100, 10),
                              (2, capi.SEARCH_PLAIN, 20, 5)):
        o = O.orc_search(q, q_low, w["base"], w["db_low"], goff, ged, e, k, mode, entry)
        g = index.search(q, q_low, e, k, entry, flags=flags)
        for key in ("ids", "dists", "hops", "dist_calc"):
            assert np.array_equal(g[key], o[key]), (w["name"], mode, e, key)


def test_projection_at_full_width(w, index):
    """K1 on the workload's own net (128-256-256-32, 96-128-128-16, 960-1024-1024-32 with sliced weight tiles) against
    the float64 evaluation: 3xTF32 keeps the unit-norm outputs within 1e-5."""
    from gbnns_dim_red_b200 import synth

    q = w["queries"][:1000]
    got = index.project(q)
    want = synth.project_numpy(*w["net"], q)
    assert np.abs(got - want).max() < 1e-5


def test_as_shipped_reference_bars_at_1m(w, index):
    """North star: per-query ids equal on >= 99.9 % of queries, distances within 1e-5 relative, recall@1 / @10 within
    0.1 pt — against the unmodified reference built with the README flags, end to end (its own GetLowQueryFromNet, its
    own search and re-rank), on every query of the workload."""
    if O.ref("fast") is None:
        pytest.skip("oracle/_ref/libgbdr_ref_fast.so not built (needs /root/reference at build time)")
    name, ef = w["name"], EF[w["name"]]
    goff, ged = w["graph"]
    threads = max(1, len(os.sched_getaffinity(0)))
    q_low_ref = O.ref_project(*w["net"], w["queries"], kind="fast")
    ref = O.ref_search(w["queries"], q_low_ref, w["base"], w["db_low"], goff, ged, ef, 1, 0, w["entry"], kind="fast",
                       threads=threads)
    g = index.search(w["queries"], None, ef, 1, w["entry"], flags=capi.SEARCH_RERANK)
    same = g["ids"] == ref["ids"]
    assert same.mean() >= 0.999, (name, same.mean())
    assert np.allclose(g["dists"][same], ref["dists"][same], rtol=1e-5)
    r1, r1_ref = workload.recall_at_1(g["ids"], w["truth"], w["base"]), workload.recall_at_1(ref["ids"], w["truth"], w["base"])
    assert abs(r1 - r1_ref) <= 0.001, (name, r1, r1_ref)
    # recall@10: the GPU's top-10 re-rank vs the exact re-rank of the reference's own ef survivors
    g10 = index.search(w["queries"], None, ef, 10, w["entry"], flags=capi.SEARCH_RERANK)
    r10 = workload.recall_at_k(g10["ids"], w["truth"], 10)
    r10_ref = workload.recall_at_k(exact_rerank_topk(ref["low_ids"], w["queries"], w["base"], 10), w["truth"], 10)
    assert abs(r10 - r10_ref) <= 0.001, (name, r10, r10_ref)


def test_ground_truth_is_the_float64_brute_force(w):
    """The recall figures are scored against truth built by the library's own kNN kernel: check it against an independent
    float64 brute force (torch, expansion form, all queries x all base vectors)."""
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    q = torch.from_numpy(w["queries"]).to(dev).double()
    qn = (q * q).sum(1)
    n = w["shape"]["n"]
    best_d = torch.full((q.shape[0],), float("inf"), dtype=torch.float64, device=dev)
    best_i = torch.zeros(q.shape[0], dtype=torch.int64, device=dev)
    step = 50_000 if w["shape"]["d"] <= 128 else 25_000
    for b0 in range(0, n, step):
        b = torch.from_numpy(w["base"][b0:b0 + step]).to(dev).double()
        d2 = qn[:, None] + (b * b).sum(1)[None, :] - 2.0 * (q @ b.T)
        m, i = d2.min(dim=1)
        upd = m < best_d
        best_d = torch.where(upd, m, best_d)
        best_i = torch.where(upd, i + b0, best_i)
    best_i = best_i.cpu().numpy()
    t0 = w["truth"][:, 0].astype(np.int64)
    agree = t0 == best_i
    # where the ids differ the two candidates must be (near-)equidistant: fp32 direct-difference vs fp64 ranking
    if not agree.all():
        bad = np.nonzero(~agree)[0]
        qa = w["queries"][bad].astype(np.float64)
        da = ((w["base"][t0[bad]].astype(np.float64) - qa) ** 2).sum(1)
        db = ((w["base"][best_i[bad]].astype(np.float64) - qa) ** 2).sum(1)
        assert np.allclose(da, db, rtol=1e-6), (w["name"], bad[:5])
    assert agree.mean() >= 0.999
    # and the truth rows are ascending in float64 distance (top-10 used by recall@10)
    qa = w["queries"][:200].astype(np.float64)
    d10 = ((w["base"][w["truth"][:200, :10].astype(np.int64)].astype(np.float64) - qa[:, None, :]) ** 2).sum(2)
    assert (np.diff(d10, axis=1) >= -1e-6 * d10[:, 1:]).all()


def test_whole_batch_properties_at_1m(w, index):
    n, n_q, ef = w["shape"]["n"], w["shape"]["n_q"], EF[w["name"]]
    a = index.search(w["queries"], None, ef, 1, w["entry"], flags=capi.SEARCH_RERANK)
    assert a["ids"].max() < n
    assert workload.recall_at_1(a["ids"], w["truth"], w["base"]) >= 0.945
    assert (a["dist_calc"] >= a["hops"] + ef).all() and (a["hops"] >= 1).all()
    # identical on a second run and on a view with other batches in flight (no run-to-run state)
    b = index.search(w["queries"], None, ef, 1, w["entry"], flags=capi.SEARCH_RERANK)
    v = index.view()
    qp = capi.pinned_empty(w["queries"].shape, np.float32)
    qp[:] = w["queries"]
    ep = capi.pinned_empty((n_q,), np.uint32)
    ep[:] = w["entry"]
    index.search_submit(qp, None, ef, 1, ep, flags=capi.SEARCH_RERANK)
    v.search_submit(qp, None, ef, 1, ep, flags=capi.SEARCH_RERANK)
    c, d = index.search_wait(), v.search_wait()
    v.close()
    for other in (b, c, d):
        for key in ("ids", "dists", "hops", "dist_calc"):
            assert np.array_equal(a[key], other[key]), key
    # top-k lists come out ascending, and the top-1 of a top-10 call is the top-1 call's answer
    t = index.search(w["queries"], None, ef, 10, w["entry"], flags=capi.SEARCH_RERANK)
    assert (np.diff(t["dists"], axis=1) >= 0).all()
    assert np.array_equal(t["ids"][:, 0], a["ids"][:, 0])
    # recall@10 (re-rank all ef survivors, top-10 vs the exact top-10) is bounded by what the beam of the
    # operating point holds, so the bar is not a constant: it must equal, within the north star's 0.1 pt, the recall@10
    # of a float64 re-rank of the SAME low-dimensional survivors
    m = min(n_q, 2000)
    s = index.search(w["queries"][:m], None, ef, ef, w["entry"][:m], flags=0)
    want10 = exact_rerank_topk(s["ids"], w["queries"][:m], w["base"], 10)
    r_gpu = workload.recall_at_k(t["ids"][:m], w["truth"][:m], 10)
    r_ref = workload.recall_at_k(want10, w["truth"][:m], 10)
    assert abs(r_gpu - r_ref) <= 1e-3, (r_gpu, r_ref)
    assert (t["ids"][:m] == want10).all(axis=1).mean() >= 0.999
    assert workload.recall_at_k(t["ids"], w["truth"], 10) >= 0.5
    # exact distances: recomputed on the host in float64 for every answer
    diff = w["base"][a["ids"][:, 0]].astype(np.float64) - w["queries"].astype(np.float64)
    assert np.allclose((diff * diff).sum(axis=1), a["dists"][:, 0], rtol=1e-5)


def test_knn_build_and_gd_prune_at_1m(w):
    if w["name"] != "sift1m":
        pytest.skip("the build chain is size-checked on the SIFT-1M shape (same d_low = 32 as GIST-1M; Deep-1M's kNN-32 graph below)")
    n, k, M = w["shape"]["n"], 1000, 30
    db_low = w["db_low"]
    pinned = capi.PinnedArray((n, k), np.uint32)
    try:
        # the HBM-resident chain: kNN lists stream to the host behind the computation, the graph comes back pruned
        goff, ged, t = capi.build_graph(db_low, knn_k=k, M=M, reverse=True, knn_out=pinned.array)
        ids = pinned.array
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
        # the host-buffer entry point (the prepare_graph.cpp drop-in) on the same lists gives the same graph
        koff, ked = xvecs.adjacency_from_matrix(ids)
        goff2, ged2, _ = capi.gd_prune(koff, ked, db_low, M=M, reverse=True)
    finally:
        pinned.close()
    assert np.array_equal(goff, goff2) and np.array_equal(ged, ged2)
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


def test_fixed_degree_graph_at_1m(w):
    """Deep-1M searches the README's "fixed" constant-degree graph: cutKNNbyK(k = 32) of the kNN lists
    (support_func.h:309-340), i.e. the 32 nearest of every list INCLUDING the vertex itself (rank 0, distance 0)."""
    if w["name"] != "deep1m":
        pytest.skip("Deep-1M shape only")
    goff, ged = w["graph"]
    n = w["shape"]["n"]
    assert (np.diff(goff.astype(np.int64)) == 32).all()
    rng = np.random.default_rng(5)
    rows = np.sort(rng.choice(n, size=32, replace=False))
    oi, _ = O.orc_knn(w["db_low"][rows], w["db_low"], 33)
    assert np.array_equal(ged.reshape(n, 32)[rows], oi[:, :32])
    assert np.array_equal(ged.reshape(n, 32)[rows, 0], rows.astype(np.uint32))


def test_index_above_4m_vertices_uses_32bit_visited_slots_and_stays_exact():
    """4.5 M vertices: ids no longer split into (bucket, <= 14-bit tag), so the plan switches to 32-bit visited slots
    (what Deep-100M-sized shards of 12.5 M rows run on).  Low-dimensional search only (16-dim vectors, kNN-32 graph
    built by the library's kNN kernel), bit-exact against the oracle on sampled queries at a small and a wide beam."""
    n, d_low, n_q = 4_500_000, 16, 2000
    assert capi.beam_plan_info(53, d_low, n)["tag_bits"] == 0
    rng = np.random.default_rng(77)
    A = rng.standard_normal((6, d_low), dtype=np.float32)
    low = np.empty((n, d_low), np.float32)
    for i in range(0, n, 1 << 20):
        j = min(n, i + (1 << 20))
        low[i:j] = rng.standard_normal((j - i, 6), dtype=np.float32) @ A + 0.05 * rng.standard_normal((j - i, d_low), dtype=np.float32)
    low /= np.linalg.norm(low, axis=1, keepdims=True)
    q = rng.standard_normal((n_q, 6), dtype=np.float32) @ A + 0.05 * rng.standard_normal((n_q, d_low), dtype=np.float32)
    q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    ids, _ = capi.knn(low, low, 33)
    rows = np.sort(rng.choice(n, size=8, replace=False))
    oi, _ = O.orc_knn(low[rows], low, 33)
    assert np.array_equal(ids[rows], oi)
    goff, ged = xvecs.adjacency_from_matrix(np.ascontiguousarray(ids[:, 1:]))
    del ids
    entry = rng.integers(0, n, size=n_q, dtype=np.uint32)
    ix = capi.Index(0)
    try:
        ix.set_low(low)
        ix.set_graph(goff, ged)
        pick = np.arange(0, n_q, 8)
        for ef, k in ((53, 10), (294, 10)):
            g = ix.search(None, q, ef, k, entry, flags=0)
            o = O.orc_search(None, q[pick], None, low, goff, ged, ef, k, 1, entry[pick])
            for key in ("ids", "dists", "hops", "dist_calc"):
                assert np.array_equal(g[key][pick], o[key]), (ef, key)
            assert (np.diff(g["dists"], axis=1) >= 0).all() and g["ids"].max() < n
    finally:
        ix.close()
