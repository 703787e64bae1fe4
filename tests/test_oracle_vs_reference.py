"""The oracle against the reference's own C++ (oracle/_ref/libgbdr_ref_strict.so = the headers under
/root/reference/search compiled strict-IEEE by oracle/Makefile) on fresh seeded inputs, wider than
the committed golden vectors.  Skipped where oracle/_ref was not built (it needs /root/reference
at build time; the built .so travels with the repo snapshot)."""
import numpy as np
import pytest

from gbnns_dim_red_b200 import synth, xvecs

from . import _oracle as O

pytestmark = pytest.mark.skipif(O.ref("strict") is None, reason="oracle/_ref not built (needs /root/reference)")


@pytest.fixture(scope="module")
def case():
    n, d, n_q, d_low, dh, M = 4000, 48, 200, 20, 40, 10   # d_low % 8 != 0: exercises Angular's 4-wide step
    base, queries = synth.make_vectors(n, d, n_q, latent=6, seed=3)
    net = synth.make_net(d, dh, d_low, seed=3)
    db_low = O.ref_project(*net, base)
    q_low = O.ref_project(*net, queries)
    knn_ids, _ = O.orc_knn(db_low, db_low, 80)
    koff, ked = xvecs.adjacency_from_matrix(knn_ids)
    goff, ged = O.ref_gd_prune(koff, ked, db_low, M=M, reverse=True)
    entry = synth.make_entry_points(n, n_q, seed=3)
    return dict(base=base, queries=queries, net=net, db_low=db_low, q_low=q_low, knn=(koff, ked), graph=(goff, ged),
                entry=entry, M=M)


def test_metrics_random():
    rng = np.random.default_rng(0)
    Lo, Lr = O.oracle(), O.ref("strict")
    for d in (4, 7, 8, 15, 16, 32, 96, 128, 301, 960):
        for _ in range(20):
            a = rng.standard_normal(d, dtype=np.float32)
            b = rng.standard_normal(d, dtype=np.float32)
            assert np.float32(Lo.orc_l2(O._p(a), O._p(b), d)) == np.float32(Lr.ref_l2(O._p(a), O._p(b), d))
            assert np.float32(Lo.orc_angular(O._p(a), O._p(b), d)) == np.float32(Lr.ref_angular(O._p(a), O._p(b), d))


def test_projection(case):
    assert np.array_equal(O.orc_project(*case["net"], case["queries"]), case["q_low"])
    assert np.array_equal(O.orc_project(*case["net"], case["base"][:500]), case["db_low"][:500])


@pytest.mark.parametrize("reverse,const_degree", [(True, False), (False, False), (True, True)])
def test_gd_prune(case, reverse, const_degree):
    koff, ked = case["knn"]
    a = O.orc_gd_prune(koff, ked, case["db_low"], M=case["M"], reverse=reverse, const_degree=const_degree)
    b = O.ref_gd_prune(koff, ked, case["db_low"], M=case["M"], reverse=reverse, const_degree=const_degree)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("M", [3, 4, 10])
def test_gd_prune_with_rows_that_fill_up(M):
    """Hub-heavy data: dozens of rows reach 2M during addReverseEdgesForGD, so the order-dependent "row full"
    test (support_func.h:429) decides edges."""
    from ._data import hub_points

    x = hub_points()
    ids, _ = O.orc_knn(x, x, 60)
    koff, ked = xvecs.adjacency_from_matrix(ids)
    for const_degree in (False, True):
        a = O.orc_gd_prune(koff, ked, x, M=M, reverse=True, const_degree=const_degree)
        b = O.ref_gd_prune(koff, ked, x, M=M, reverse=True, const_degree=const_degree)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert (np.diff(a[0].astype(np.int64)) == 2 * M).sum() >= 50


@pytest.mark.parametrize("ef", [1, 2, 7, 32, 33, 100, 300])
def test_search_rerank(case, ef):
    goff, ged = case["graph"]
    a = O.orc_search(case["queries"], case["q_low"], case["base"], case["db_low"], goff, ged, ef, 1, 0, case["entry"])
    b = O.ref_search(case["queries"], case["q_low"], case["base"], case["db_low"], goff, ged, ef, 1, 0, case["entry"])
    for key in ("ids", "dists", "hops", "dist_calc"):
        assert np.array_equal(a[key], b[key]), key


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("ef,k", [(5, 5), (50, 10), (120, 120)])
def test_search_modes(case, mode, ef, k):
    goff, ged = case["graph"]
    args = (case["queries"], case["q_low"], case["base"], case["db_low"], goff, ged, ef, k, mode, case["entry"])
    a, b = O.orc_search(*args), O.ref_search(*args)
    for key in ("ids", "dists", "hops", "dist_calc"):
        assert np.array_equal(a[key], b[key]), key


@pytest.mark.parametrize("llf", [False, True])
@pytest.mark.parametrize("hops_bound", [0, 3, 50])
@pytest.mark.parametrize("mode,ef,k", [(0, 20, 1), (1, 40, 10), (2, 7, 7)])
def test_search_second_graph(case, llf, hops_bound, mode, ef, k):
    """use_second_graph == true (search_function.h:73-89): auxiliary row first while hops < hops_bound, main row
    skipped when llf and the auxiliary row produced a candidate."""
    from ._data import long_link_graph

    goff, ged = case["graph"]
    aux = long_link_graph(case["base"].shape[0])
    args = (case["queries"], case["q_low"], case["base"], case["db_low"], goff, ged, ef, k, mode, case["entry"])
    a = O.orc_search(*args, aux=aux, llf=llf, hops_bound=hops_bound)
    b = O.ref_search(*args, aux=aux, llf=llf, hops_bound=hops_bound)
    for key in ("ids", "dists", "hops", "dist_calc"):
        assert np.array_equal(a[key], b[key]), key
    if hops_bound == 0:  # the second graph is never consulted: identical to the single-graph search
        c = O.orc_search(*args)
        for key in ("ids", "dists", "hops", "dist_calc"):
            assert np.array_equal(a[key], c[key]), key
    else:
        assert not np.array_equal(a["dist_calc"], O.orc_search(*args)["dist_calc"])


def test_search_with_exact_distance_ties():
    """Duplicate base vectors make exact float ties everywhere: the (dist,id) tie rules of the two
    priority queues (search_function.h:50,55) and the strict comparisons (:31,:67,:117) must match."""
    n, d, n_q = 1200, 16, 100
    rng = np.random.default_rng(9)
    uniq = rng.integers(-2, 3, size=(n // 4, d)).astype(np.float32)  # small integer grid: many equal distances
    base = np.repeat(uniq, 4, axis=0)[rng.permutation(n)]
    queries = rng.integers(-2, 3, size=(n_q, d)).astype(np.float32)
    knn_ids, _ = O.orc_knn(base, base, 24)
    off, ed = xvecs.adjacency_from_matrix(knn_ids[:, 1:])
    entry = synth.make_entry_points(n, n_q, seed=2)
    for ef, k in ((1, 1), (6, 6), (20, 7), (64, 64)):
        a = O.orc_search(queries, None, base, None, off, ed, ef, k, 2, entry)
        b = O.ref_search(queries, None, base, None, off, ed, ef, k, 2, entry)
        for key in ("ids", "dists", "hops", "dist_calc"):
            assert np.array_equal(a[key], b[key]), (ef, k, key)
        a = O.orc_search(queries, queries, base, base, off, ed, ef, 1, 0, entry)
        b = O.ref_search(queries, queries, base, base, off, ed, ef, 1, 0, entry)
        assert np.array_equal(a["ids"], b["ids"])


def test_as_shipped_build_agrees_within_tolerance(case):
    """The README-flag (-Ofast) build may reassociate; ids must agree on >= 99.9 % of queries and
    distances within 1e-5 relative (BASELINE.json north_star)."""
    if O.ref("fast") is None:
        pytest.skip("fast reference build absent")
    goff, ged = case["graph"]
    args = (case["queries"], case["q_low"], case["base"], case["db_low"], goff, ged, 40, 1, 0, case["entry"])
    a = O.orc_search(*args)
    b = O.ref_search(*args, kind="fast")
    assert (a["ids"] == b["ids"]).mean() >= 0.999
    same = a["ids"] == b["ids"]
    assert np.allclose(a["dists"][same], b["dists"][same], rtol=1e-5)
    pf = O.ref_project(*case["net"], case["queries"], kind="fast")
    assert np.abs(pf - case["q_low"]).max() < 2e-6


def test_random_small_graphs_property():
    """Randomised differential check (seeded): arbitrary small digraphs — self loops, repeated neighbours, vertices
    without out-edges, integer-grid vectors (many exact ties) — random ef / k / mode, with and without a second
    graph: the oracle must reproduce the reference's ids, distances, hops and dist_calc."""
    rng = np.random.default_rng(2024)
    for trial in range(60):
        n = int(rng.integers(2, 60))
        d = int(rng.choice([4, 8, 12]))
        grid = trial % 2 == 0
        base = (rng.integers(-2, 3, size=(n, d)) if grid else rng.standard_normal((n, d))).astype(np.float32)
        n_q = 12
        queries = (rng.integers(-2, 3, size=(n_q, d)) if grid else rng.standard_normal((n_q, d))).astype(np.float32)

        def rand_graph(max_deg):
            lists = []
            for i in range(n):
                deg = int(rng.integers(0, max_deg + 1))
                lists.append(rng.integers(0, n, size=deg).astype(np.uint32).tolist())   # repeats and self loops allowed
            return xvecs.adjacency_from_lists(lists)

        off, ed = rand_graph(6)
        aux = rand_graph(3) if trial % 3 == 0 else None
        entry = rng.integers(0, n, size=n_q, dtype=np.uint32)
        ef = int(rng.integers(1, 12))
        k = int(rng.integers(1, ef + 1))
        mode = int(rng.integers(0, 3))
        kw = dict(aux=aux, llf=bool(trial % 2), hops_bound=int(rng.integers(0, 5))) if aux is not None else {}
        if mode == 0:
            k = 1   # the reference's re-rank returns one id
        a = O.orc_search(queries, queries, base, base, off, ed, ef, k, mode, entry, **kw)
        b = O.ref_search(queries, queries, base, base, off, ed, ef, k, mode, entry, **kw)
        for key in ("ids", "dists", "hops", "dist_calc"):
            assert np.array_equal(a[key], b[key]), (trial, n, d, ef, k, mode, key)


@pytest.mark.parametrize("knn_size", [1, 7, 32, 200])
def test_knn_cut(case, knn_size):
    """cutKNNbyK (support_func.h:309-340): the oracle's restatement against the reference, on the kNN-80 lists in a
    shuffled order (so the sort matters) and with a few short rows (the reference's "Size knn less than you want")."""
    if not hasattr(O.ref("strict"), "ref_knn_cut"):
        pytest.skip("oracle/_ref predates ref_knn_cut")
    koff, ked = case["knn"]
    rng = np.random.default_rng(knn_size)
    rows = [rng.permutation(ked[int(koff[i]):int(koff[i + 1])]) for i in range(koff.size - 1)]
    for i in rng.integers(0, len(rows), 25):
        rows[i] = rows[i][: rng.integers(0, 6)]
    off = np.zeros(len(rows) + 1, np.uint64)
    off[1:] = np.cumsum([r.size for r in rows])
    edges = np.concatenate(rows).astype(np.uint32)
    a = O.orc_knn_cut(off, edges, case["db_low"], knn_size)
    b = O.ref_knn_cut(off, edges, case["db_low"], knn_size)
    assert np.array_equal(a[0], b[0])
    # identical except inside groups of EXACTLY equal distances, which the reference's std::sort (dist only, :63-66)
    # leaves in an unspecified order and the oracle orders by id
    db = case["db_low"]
    L = O.oracle()
    for j in np.nonzero(a[1] != b[1])[0]:
        i = int(np.searchsorted(a[0], j, side="right") - 1)
        da = L.orc_l2(O._p(db[i]), O._p(db[a[1][j]]), db.shape[1])
        dr = L.orc_l2(O._p(db[i]), O._p(db[b[1][j]]), db.shape[1])
        assert da == dr, (i, j)
        lo, hi = int(a[0][i]), int(a[0][i + 1])
        if hi - lo == int(off[i + 1] - off[i]):  # nothing was cut off: same multiset
            assert sorted(a[1][lo:hi]) == sorted(b[1][lo:hi])
    assert (a[1] != b[1]).mean() < 1e-3
    # sorted kNN lists are their own prefix
    c = O.orc_knn_cut(koff, ked, case["db_low"], min(knn_size, 80))
    assert np.array_equal(c[1].reshape(-1, min(knn_size, 80)), ked.reshape(-1, 80)[:, : min(knn_size, 80)])
