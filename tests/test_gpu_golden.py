"""GPU parity against the committed golden vectors (outputs of the reference's own C++, see
tests/golden/make_golden.py) through the C ABI."""
import os

import numpy as np
import pytest

from gbnns_dim_red_b200 import capi, xvecs

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def g():
    return dict(np.load(os.path.join(G, "search.npz")))


@pytest.fixture()
def ix(gpu_index_factory, g):
    ix = gpu_index_factory()
    ix.set_base(g["base"])
    ix.set_low(g["db_low"])
    ix.set_graph(g["goff"], g["gedges"])
    return ix


@pytest.mark.parametrize("ef", [1, 4, 16, 40])
def test_search_rerank_golden(ix, g, ef):
    r = ix.search(g["queries"], g["q_low"], ef, 1, g["entry"], flags=capi.SEARCH_RERANK)
    for key in ("ids", "dists", "hops", "dist_calc"):
        assert np.array_equal(r[key], g[f"rerank_ef{ef}_{key}"]), key
    low = ix.search(None, g["q_low"], ef, ef, g["entry"], flags=0)
    assert np.array_equal(low["ids"], g[f"rerank_ef{ef}_low_ids"])
    assert np.array_equal(low["dists"], g[f"rerank_ef{ef}_low_dists"])


@pytest.mark.parametrize("ef,k", [(8, 8), (24, 5)])
def test_low_and_plain_golden(ix, g, ef, k):
    r = ix.search(None, g["q_low"], ef, k, g["entry"], flags=0)
    for key in ("ids", "dists", "hops", "dist_calc"):
        assert np.array_equal(r[key], g[f"low_ef{ef}_k{k}_{key}"]), key
    r = ix.search(g["queries"], None, ef, k, g["entry"], flags=capi.SEARCH_PLAIN)
    for key in ("ids", "dists", "hops", "dist_calc"):
        assert np.array_equal(r[key], g[f"plain_ef{ef}_k{k}_{key}"]), key


@pytest.mark.parametrize("variant", ["", "_nr", "_cd"])
def test_gd_prune_golden(g, variant):
    koff, ked = xvecs.adjacency_from_matrix(g["knn_ids"])
    off, ed, _ = capi.gd_prune(koff, ked, g["db_low"], M=int(g["M"]), reverse=variant != "_nr",
                               need_const_degree=variant == "_cd")
    assert np.array_equal(off, g["goff" + variant])
    assert np.array_equal(ed, g["gedges" + variant])


def test_projection_golden(gpu_index_factory, g):
    """fp32-class projection modes: within 2e-6 absolute of GetLowQueryFromNet on unit-norm outputs
    (the reference's own -Ofast build differs from its strict build by up to ~3e-7)."""
    ix = gpu_index_factory()
    ix.set_net(g["l1"], g["l2"], g["l3"])
    for mode in (capi.PROJ_FP32, capi.PROJ_3XTF32):
        ix.set_projection_mode(mode)
        y = ix.project(g["queries"])
        assert np.abs(y - g["q_low"]).max() < 2e-6, mode


@pytest.fixture(scope="module")
def g2():
    return dict(np.load(os.path.join(G, "second.npz")))


@pytest.mark.parametrize("llf", [0, 1])
@pytest.mark.parametrize("hb", [3, 50])
def test_second_graph_golden(ix, g, g2, llf, hb):
    ix.set_aux_graph(g2["aoff"], g2["aedges"], hops_bound=hb, llf=bool(llf))
    for mode, flags, ef, k in ((0, capi.SEARCH_RERANK, 16, 1), (1, 0, 24, 5), (2, capi.SEARCH_PLAIN, 8, 8)):
        r = ix.search(g["queries"], g["q_low"], ef, k, g["entry"], flags=flags | capi.SEARCH_SECOND_GRAPH)
        for key in ("ids", "dists", "hops", "dist_calc"):
            assert np.array_equal(r[key], g2[f"aux_llf{llf}_hb{hb}_m{mode}_{key}"]), (mode, key)


@pytest.mark.parametrize("M", [3, 6])
@pytest.mark.parametrize("cd", [0, 1])
def test_gd_prune_hub_golden(g2, M, cd):
    koff, ked = xvecs.adjacency_from_matrix(g2["hub_knn"])
    off, ed, _ = capi.gd_prune(koff, ked, g2["hub_x"], M=M, reverse=True, need_const_degree=bool(cd))
    assert np.array_equal(off, g2[f"hub_M{M}_cd{cd}_off"]) and np.array_equal(ed, g2[f"hub_M{M}_cd{cd}_edges"])


# ------------------------------------------------------------------------------------------------------------------
# the reference AS SHIPPED (-Ofast build, tests/golden/fast.npz): BASELINE.json's tolerance bars, GPU end to end
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["small", "mid"])
def test_gpu_end_to_end_within_the_bars_of_the_as_shipped_reference(gpu_index_factory, name):
    """Queries go in in the ORIGINAL dimension (projection on the tensor cores, 3xTF32), exactly as the as-shipped
    reference got them; ids equal on >= 99.9 % of queries, distances within 1e-5 relative, recall@1 / @10 within 0.1 pt
    of the reference's (north star).  Recall@10 = re-rank of the ef survivors, top 10 by (dist, id)."""
    from gbnns_dim_red_b200 import workload

    from . import _oracle as O
    from ._data import exact_rerank_topk, mid_case, small_case

    fast = dict(np.load(os.path.join(G, "fast.npz")))
    c = {"small": small_case, "mid": mid_case}[name]()
    goff, ged = c["graph"]
    assert int(ged.astype(np.uint64).sum()) == int(fast[f"{name}_edges_sum"]), "seeded graph drifted"
    ix = gpu_index_factory()
    ix.set_net(*c["net"])
    ix.set_base(c["base"])
    ix.set_low(c["db_low"])
    ix.set_graph(goff, ged)
    truth, _ = capi.knn(c["queries"], c["base"], 10)
    ot, _ = O.orc_knn(c["queries"][:64], c["base"], 10)
    assert np.array_equal(truth[:64], ot)
    for ef in (10, 40, 100):
        r = ix.search(c["queries"], None, ef, 1, c["entry"], flags=capi.SEARCH_RERANK)
        ref_ids, ref_d = fast[f"{name}_ef{ef}_ids"], fast[f"{name}_ef{ef}_dists"]
        same = r["ids"] == ref_ids
        assert same.mean() >= 0.999, (name, ef, same.mean())
        assert np.allclose(r["dists"][same], ref_d[same], rtol=1e-5)
        assert abs(workload.recall_at_1(r["ids"], truth) - workload.recall_at_1(ref_ids, truth)) <= 0.001
        if ef >= 10:
            r10 = ix.search(c["queries"], None, ef, 10, c["entry"], flags=capi.SEARCH_RERANK)
            got = workload.recall_at_k(r10["ids"], truth, 10)
            want = workload.recall_at_k(exact_rerank_topk(fast[f"{name}_ef{ef}_low_ids"], c["queries"], c["base"], 10), truth, 10)
            assert abs(got - want) <= 0.001, (name, ef, got, want)
