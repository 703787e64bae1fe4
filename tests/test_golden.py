"""The CPU oracle (oracle/gbdr_oracle.c) against the committed golden vectors in tests/golden/,
which were produced by the REFERENCE itself (tests/golden/make_golden.py: the reference's C++
headers compiled strict-IEEE, and its Python kNN).  This is what pins the oracle; it runs anywhere
(no GPU, no /root/reference)."""
import os

import numpy as np
import pytest

from gbnns_dim_red_b200 import xvecs

from . import _oracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def g():
    return dict(np.load(os.path.join(G, "search.npz")))


def test_l2_and_angular_values(g):
    L = O.oracle()
    for i in range(g["mv_a"].shape[0]):
        a, b = np.ascontiguousarray(g["mv_a"][i]), np.ascontiguousarray(g["mv_b"][i])
        for j, d in enumerate(g["dims"]):
            assert np.float32(L.orc_l2(O._p(a), O._p(b), int(d))) == g["l2_vals"][i, j]
            assert np.float32(L.orc_angular(O._p(a), O._p(b), int(d))) == g["ang_vals"][i, j]


def test_projection_bit_exact(g):
    assert np.array_equal(O.orc_project(g["l1"], g["l2"], g["l3"], g["base"]), g["db_low"])
    assert np.array_equal(O.orc_project(g["l1"], g["l2"], g["l3"], g["queries"]), g["q_low"])


@pytest.mark.parametrize("variant", ["", "_nr", "_cd"])
def test_gd_prune_graph_identical(g, variant):
    koff, ked = xvecs.adjacency_from_matrix(g["knn_ids"])
    reverse = variant != "_nr"
    off, ed = O.orc_gd_prune(koff, ked, g["db_low"], M=int(g["M"]), reverse=reverse, const_degree=variant == "_cd")
    assert np.array_equal(off, g["goff" + variant])
    assert np.array_equal(ed, g["gedges" + variant])


@pytest.mark.parametrize("ef", [1, 4, 16, 40])
def test_search_rerank(g, ef):
    r = O.orc_search(g["queries"], g["q_low"], g["base"], g["db_low"], g["goff"], g["gedges"], ef, 1, 0, g["entry"])
    for key in ("ids", "dists", "hops", "dist_calc"):
        assert np.array_equal(r[key], g[f"rerank_ef{ef}_{key}"]), key
    # the low-dim heap handed to getRealNearest
    low = O.orc_search(None, g["q_low"], None, g["db_low"], g["goff"], g["gedges"], ef, ef, 1, g["entry"])
    assert np.array_equal(low["ids"], g[f"rerank_ef{ef}_low_ids"])
    assert np.array_equal(low["dists"], g[f"rerank_ef{ef}_low_dists"])


@pytest.mark.parametrize("ef,k", [(8, 8), (24, 5)])
def test_search_low_and_plain(g, ef, k):
    r = O.orc_search(None, g["q_low"], None, g["db_low"], g["goff"], g["gedges"], ef, k, 1, g["entry"])
    for key in ("ids", "dists", "hops", "dist_calc"):
        assert np.array_equal(r[key], g[f"low_ef{ef}_k{k}_{key}"]), key
    r = O.orc_search(g["queries"], None, g["base"], None, g["goff"], g["gedges"], ef, k, 2, g["entry"])
    for key in ("ids", "dists", "hops", "dist_calc"):
        assert np.array_equal(r[key], g[f"plain_ef{ef}_k{k}_{key}"]), key


def test_knn_against_reference_python():
    """get_nearestneighbors_torch (dim_red/support_func.py:54-68) computes ||a||^2-2ab+||b||^2 in
    fp32, which reorders near-ties; the oracle ranks exact direct-difference distances by
    (dist,id).  Bar: same neighbour SET on >= 99 % of rows, identical sequence wherever the oracle's
    consecutive distances are separated by more than the expansion form's rounding error."""
    z = np.load(os.path.join(G, "knn.npz"))
    db_low, want = z["db_low"], z["torch_knn"]
    k = want.shape[1]
    ids, dists = O.orc_knn(db_low, db_low, k + 1)
    m = want.shape[0]  # the reference drops the last n % 500 rows (:63-64)
    assert m == (db_low.shape[0] // 500) * 500
    same_set = sum(set(ids[i, :k].tolist()) == set(want[i].tolist()) for i in range(m))
    assert same_set >= 0.99 * m
    tol = 2e-6  # |x|=1 after normalisation -> expansion-form error ~ a few ulp of 2.0
    checked = 0
    for i in range(m):
        gaps = np.diff(dists[i])
        ok = gaps > tol
        for j in range(k):
            if (j == 0 or ok[j - 1]) and ok[j]:
                assert ids[i, j] == want[i, j]
                checked += 1
    assert checked > 0.9 * m * k
    assert (ids[:m, 0] == np.arange(m)).all()  # self at rank 0


@pytest.fixture(scope="module")
def g2():
    return dict(np.load(os.path.join(G, "second.npz")))


@pytest.mark.parametrize("llf", [0, 1])
@pytest.mark.parametrize("hb", [3, 50])
def test_second_graph_search(g, g2, llf, hb):
    """use_second_graph == true (search_function.h:73-89), recorded from the reference's own C++."""
    for mode, ef, k in ((0, 16, 1), (1, 24, 5), (2, 8, 8)):
        r = O.orc_search(g["queries"], g["q_low"], g["base"], g["db_low"], g["goff"], g["gedges"], ef, k, mode, g["entry"],
                         aux=(g2["aoff"], g2["aedges"]), llf=bool(llf), hops_bound=hb)
        for key in ("ids", "dists", "hops", "dist_calc"):
            assert np.array_equal(r[key], g2[f"aux_llf{llf}_hb{hb}_m{mode}_{key}"]), (mode, key)


@pytest.mark.parametrize("M", [3, 6])
@pytest.mark.parametrize("cd", [0, 1])
def test_gd_prune_hub_graph_identical(g2, M, cd):
    """hnswlikeGD where rows fill to 2M during addReverseEdgesForGD (support_func.h:423-442)."""
    koff, ked = xvecs.adjacency_from_matrix(g2["hub_knn"])
    off, ed = O.orc_gd_prune(koff, ked, g2["hub_x"], M=M, reverse=True, const_degree=bool(cd))
    assert np.array_equal(off, g2[f"hub_M{M}_cd{cd}_off"]) and np.array_equal(ed, g2[f"hub_M{M}_cd{cd}_edges"])


# ------------------------------------------------------------------------------------------------------------------
# the reference AS SHIPPED (README flags, -Ofast): tests/golden/fast.npz, the north star's tolerance bars
# ------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def fast():
    return dict(np.load(os.path.join(G, "fast.npz")))


def _cases():
    from ._data import mid_case, small_case

    return {"small": small_case, "mid": mid_case}


@pytest.mark.parametrize("name", ["small", "mid"])
def test_oracle_within_the_north_star_bars_of_the_as_shipped_reference(fast, name):
    """End to end as final_test.cpp runs it (projection of every query, low-dim search, original-dim re-rank): result
    ids equal on >= 99.9 % of queries, distances within 1e-5 relative, recall@1 / @10 within 0.1 pt."""
    from gbnns_dim_red_b200 import workload

    from ._data import exact_rerank_topk

    c = _cases()[name]()
    goff, ged = c["graph"]
    assert np.isclose(c["base"].astype(np.float64).sum(), fast[f"{name}_base_sum"]), "seeded inputs drifted"
    assert int(ged.astype(np.uint64).sum()) == int(fast[f"{name}_edges_sum"]), "seeded graph drifted"
    assert np.abs(c["q_low"] - fast[f"{name}_q_low"]).max() < 2e-6  # strict vs -Ofast GetLowQueryFromNet
    truth, _ = O.orc_knn(c["queries"], c["base"], 10)
    for ef in (10, 40, 100):
        o = O.orc_search(c["queries"], c["q_low"], c["base"], c["db_low"], goff, ged, ef, 1, 0, c["entry"])
        ref_ids, ref_d = fast[f"{name}_ef{ef}_ids"], fast[f"{name}_ef{ef}_dists"]
        same = o["ids"] == ref_ids
        assert same.mean() >= 0.999, (name, ef, same.mean())
        assert np.allclose(o["dists"][same], ref_d[same], rtol=1e-5)
        assert abs(workload.recall_at_1(o["ids"], truth) - workload.recall_at_1(ref_ids, truth)) <= 0.001
        low = O.orc_search(None, c["q_low"], None, c["db_low"], goff, ged, ef, ef, 1, c["entry"])
        r10 = workload.recall_at_k(exact_rerank_topk(low["ids"], c["queries"], c["base"], 10), truth, 10)
        r10_ref = workload.recall_at_k(exact_rerank_topk(fast[f"{name}_ef{ef}_low_ids"], c["queries"], c["base"], 10), truth, 10)
        assert abs(r10 - r10_ref) <= 0.001, (name, ef, r10, r10_ref)
