"""GPU parity: beam search (K2) + re-rank (K3) through the C ABI vs the CPU oracle.

Bit-exact bar: ids, distances, hops and dist_calc must be IDENTICAL to the oracle (which is itself
bit-identical to the reference's strict-FP build, tests/test_oracle_vs_reference.py)."""
import numpy as np
import pytest

from gbnns_dim_red_b200 import capi

from . import _oracle as O
from ._data import long_link_graph, small_case

pytestmark = pytest.mark.gpu


def _index(make, c, low=True, base=True):
    ix = make()
    if base:
        ix.set_base(c["base"])
    if low:
        ix.set_low(c["db_low"])
    ix.set_graph(*c["graph"])
    return ix


@pytest.mark.parametrize("ef", [1, 3, 8, 24, 25, 40, 56, 57, 100, 120, 121, 180, 248, 249, 400])
def test_search_rerank_matches_oracle(gpu_index_factory, ef):
    c = small_case()
    ix = _index(gpu_index_factory, c)
    goff, ged = c["graph"]
    o = O.orc_search(c["queries"], c["q_low"], c["base"], c["db_low"], goff, ged, ef, 1, 0, c["entry"])
    g = ix.search(c["queries"], c["q_low"], ef, 1, c["entry"], flags=capi.SEARCH_RERANK)
    assert np.array_equal(g["ids"], o["ids"])
    assert np.array_equal(g["dists"], o["dists"])
    assert np.array_equal(g["hops"], o["hops"])
    assert np.array_equal(g["dist_calc"], o["dist_calc"])


@pytest.mark.parametrize("ef,k", [(30, 10), (64, 64), (10, 1), (200, 100)])
def test_lowdim_only_matches_oracle(gpu_index_factory, ef, k):
    c = small_case()
    ix = _index(gpu_index_factory, c)
    goff, ged = c["graph"]
    o = O.orc_search(None, c["q_low"], None, c["db_low"], goff, ged, ef, k, 1, c["entry"])
    g = ix.search(None, c["q_low"], ef, k, c["entry"], flags=0)
    assert np.array_equal(g["ids"], o["ids"])
    assert np.array_equal(g["dists"], o["dists"])
    assert np.array_equal(g["hops"], o["hops"])
    assert np.array_equal(g["dist_calc"], o["dist_calc"])


@pytest.mark.parametrize("ef,k", [(20, 5), (50, 50)])
def test_plain_search_matches_oracle(gpu_index_factory, ef, k):
    c = small_case()
    ix = _index(gpu_index_factory, c, low=False)
    goff, ged = c["graph"]
    o = O.orc_search(c["queries"], None, c["base"], None, goff, ged, ef, k, 2, c["entry"])
    g = ix.search(c["queries"], None, ef, k, c["entry"], flags=capi.SEARCH_PLAIN)
    assert np.array_equal(g["ids"], o["ids"])
    assert np.array_equal(g["dists"], o["dists"])
    assert np.array_equal(g["hops"], o["hops"])
    assert np.array_equal(g["dist_calc"], o["dist_calc"])


@pytest.mark.parametrize("k", [1, 10, 40])
def test_rerank_topk_matches_oracle(gpu_index_factory, k):
    c = small_case()
    ix = _index(gpu_index_factory, c)
    goff, ged = c["graph"]
    o = O.orc_search(c["queries"], c["q_low"], c["base"], c["db_low"], goff, ged, 40, k, 0, c["entry"])
    g = ix.search(c["queries"], c["q_low"], 40, k, c["entry"], flags=capi.SEARCH_RERANK)
    assert np.array_equal(g["ids"], o["ids"])
    assert np.array_equal(g["dists"], o["dists"])


def test_shared_memory_list_variant_matches_oracle(gpu_index_factory, monkeypatch):
    """ef + slack <= 256 normally uses the register-resident list; force the shared-memory list."""
    monkeypatch.setenv("GBDR_BEAM_VARIANT", "smem")
    c = small_case()
    ix = _index(gpu_index_factory, c)
    goff, ged = c["graph"]
    for ef in (5, 40, 100):
        o = O.orc_search(c["queries"], c["q_low"], c["base"], c["db_low"], goff, ged, ef, 1, 0, c["entry"])
        g = ix.search(c["queries"], c["q_low"], ef, 1, c["entry"], flags=capi.SEARCH_RERANK)
        for key in ("ids", "dists", "hops", "dist_calc"):
            assert np.array_equal(g[key], o[key]), (ef, key)


@pytest.mark.parametrize("variant", ["v2", "reg", "smem"])
def test_duplicates_and_ties(gpu_index_factory, monkeypatch, variant):
    monkeypatch.setenv("GBDR_BEAM_VARIANT", variant)
    _duplicates_and_ties(gpu_index_factory)


def _duplicates_and_ties(gpu_index_factory):
    """Exact distance ties: a base with every vector duplicated 3x (the SIFT situation the
    reference special-cases at search_function.h:193-202) must still match id-for-id."""
    c = small_case()
    n0 = 600
    base = np.repeat(c["base"][:n0], 3, axis=0)
    low = np.repeat(c["db_low"][:n0], 3, axis=0)
    knn_ids, _ = O.orc_knn(low, low, 48)
    from gbnns_dim_red_b200 import xvecs

    koff, ked = xvecs.adjacency_from_matrix(knn_ids)
    # plain kNN graph (GD would drop the zero-distance duplicates)
    ix = gpu_index_factory()
    ix.set_base(base)
    ix.set_low(low)
    ix.set_graph(koff, ked)
    entry = c["entry"] % (3 * n0)
    for ef in (4, 16, 50):
        o = O.orc_search(c["queries"], c["q_low"], base, low, koff, ked, ef, 1, 0, entry)
        g = ix.search(c["queries"], c["q_low"], ef, 1, entry, flags=capi.SEARCH_RERANK)
        assert np.array_equal(g["ids"], o["ids"])
        assert np.array_equal(g["hops"], o["hops"])
        assert np.array_equal(g["dist_calc"], o["dist_calc"])
        o = O.orc_search(None, c["q_low"], None, low, koff, ked, ef, ef, 1, entry)
        g = ix.search(None, c["q_low"], ef, ef, entry, flags=0)
        assert np.array_equal(g["ids"], o["ids"])


def test_visited_spill_is_exact(gpu_index_factory, monkeypatch):
    """Force a tiny shared-memory visited table so that queries spill to the HBM table."""
    c = small_case()
    ix = _index(gpu_index_factory, c)
    goff, ged = c["graph"]
    monkeypatch.setenv("GBDR_BEAM_HCAP", "128")
    o = O.orc_search(c["queries"], c["q_low"], c["base"], c["db_low"], goff, ged, 100, 1, 0, c["entry"])
    g = ix.search(c["queries"], c["q_low"], 100, 1, c["entry"], flags=capi.SEARCH_RERANK)
    assert ix.status() & 1, "expected the spill path to be exercised"
    assert np.array_equal(g["ids"], o["ids"])
    assert np.array_equal(g["hops"], o["hops"])
    assert np.array_equal(g["dist_calc"], o["dist_calc"])


def test_ragged_and_edge_cases(gpu_index_factory):
    c = small_case()
    ix = _index(gpu_index_factory, c)
    goff, ged = c["graph"]
    # n_q = 0
    g = ix.search(c["queries"][:0], c["q_low"][:0], 10, 1, c["entry"][:0])
    assert g["ids"].shape == (0, 1)
    # n_q = 1, k == ef
    o = O.orc_search(c["queries"][:1], c["q_low"][:1], c["base"], c["db_low"], goff, ged, 7, 7, 0, c["entry"][:1])
    g = ix.search(c["queries"][:1], c["q_low"][:1], 7, 7, c["entry"][:1])
    assert np.array_equal(g["ids"], o["ids"])
    # bad arguments fail loudly
    with pytest.raises(capi.GbdrError):
        ix.search(c["queries"], c["q_low"], 5, 6, c["entry"])  # k > ef
    with pytest.raises(capi.GbdrError):
        ix.search(None, c["q_low"], 5, 1, c["entry"], flags=capi.SEARCH_RERANK)  # re-rank without queries


def test_isolated_component_pads(gpu_index_factory):
    """A query whose entry vertex has no out-edges returns that vertex and PAD for the rest."""
    c = small_case()
    n = 64
    low = c["db_low"][:n]
    lists = [[(i + 1) % 32, (i + 5) % 32] if i < 32 else [] for i in range(n)]
    from gbnns_dim_red_b200 import xvecs

    off, ed = xvecs.adjacency_from_lists(lists)
    ix = gpu_index_factory()
    ix.set_low(low)
    ix.set_graph(off, ed)
    entry = np.array([40, 3, 63, 0], dtype=np.uint32)
    o = O.orc_search(None, c["q_low"][:4], None, low, off, ed, 8, 8, 1, entry)
    g = ix.search(None, c["q_low"][:4], 8, 8, entry, flags=0)
    assert np.array_equal(g["ids"], o["ids"])
    assert g["ids"][0, 0] == 40 and (g["ids"][0, 1:] == capi.PAD_ID).all()
    assert np.array_equal(g["hops"], o["hops"])


@pytest.mark.parametrize("variant", ["v2", "reg"])
@pytest.mark.parametrize("ef", [1, 2, 9, 24, 53, 56, 57, 100, 180, 248])
def test_register_variants_match_oracle(gpu_index_factory, monkeypatch, variant, ef):
    """Both register-list kernels (batched-merge v2 and the sequential one) on the same inputs."""
    monkeypatch.setenv("GBDR_BEAM_VARIANT", variant)
    c = small_case()
    ix = _index(gpu_index_factory, c)
    goff, ged = c["graph"]
    o = O.orc_search(c["queries"], c["q_low"], c["base"], c["db_low"], goff, ged, ef, 1, 0, c["entry"])
    g = ix.search(c["queries"], c["q_low"], ef, 1, c["entry"], flags=capi.SEARCH_RERANK)
    for key in ("ids", "dists", "hops", "dist_calc"):
        assert np.array_equal(g[key], o[key]), key


@pytest.mark.parametrize("d_low", [16, 32, 48, 64, 24])
def test_row_widths(gpu_index_factory, d_low):
    """d_low 16/32/48/64 take the v2 kernel (C = 4/8/12/16), 24 the generic register kernel."""
    c = small_case(n=2500, d=64, n_q=128, d_low=d_low, dh=64, seed=4, M=10, knn_k=64)
    ix = _index(gpu_index_factory, c)
    goff, ged = c["graph"]
    for ef, k in ((3, 1), (40, 40), (120, 10)):
        o = O.orc_search(None, c["q_low"], None, c["db_low"], goff, ged, ef, k, 1, c["entry"])
        g = ix.search(None, c["q_low"], ef, k, c["entry"], flags=0)
        for key in ("ids", "dists", "hops", "dist_calc"):
            assert np.array_equal(g[key], o[key]), (d_low, ef, key)


@pytest.mark.parametrize("variant", ["v2", "reg", "smem"])
def test_integer_grid_ties(gpu_index_factory, monkeypatch, variant):
    """Vectors on a small integer grid: exact float ties between DIFFERENT vertices at every hop, the
    case where the batched merge must hand over to the sequential accept/evict rules."""
    monkeypatch.setenv("GBDR_BEAM_VARIANT", variant)
    n, d, n_q = 1500, 16, 128
    rng = np.random.default_rng(21)
    base = rng.integers(-2, 3, size=(n, d)).astype(np.float32)
    queries = rng.integers(-2, 3, size=(n_q, d)).astype(np.float32)
    knn_ids, _ = O.orc_knn(base, base, 33)
    from gbnns_dim_red_b200 import xvecs

    off, ed = xvecs.adjacency_from_matrix(knn_ids[:, 1:])
    ix = gpu_index_factory()
    ix.set_low(base)
    ix.set_graph(off, ed)
    entry = rng.integers(0, n, size=n_q).astype(np.uint32)
    for ef, k in ((1, 1), (5, 5), (24, 24), (50, 7), (100, 100)):
        o = O.orc_search(None, queries, None, base, off, ed, ef, k, 1, entry)
        g = ix.search(None, queries, ef, k, entry, flags=0)
        for key in ("ids", "dists", "hops", "dist_calc"):
            assert np.array_equal(g[key], o[key]), (ef, k, key)


@pytest.mark.parametrize("vis16", ["0", "1"])
@pytest.mark.parametrize("ef", [1, 9, 24, 53, 56, 57, 87, 88, 120, 121, 174, 175, 248, 249, 400])
def test_visited_table_formats_match_oracle(gpu_index_factory, monkeypatch, vis16, ef):
    """beam_search_v2 with 32-bit visited slots and with 16-bit tags, across every list capacity (32 ... 512 slots)
    and every tag-table size the plan picks (256 ... 2048 buckets)."""
    monkeypatch.setenv("GBDR_BEAM_VARIANT", "v2")
    monkeypatch.setenv("GBDR_BEAM_VIS16", vis16)
    c = small_case()
    ix = _index(gpu_index_factory, c)
    goff, ged = c["graph"]
    o = O.orc_search(c["queries"], c["q_low"], c["base"], c["db_low"], goff, ged, ef, 1, 0, c["entry"])
    g = ix.search(c["queries"], c["q_low"], ef, 1, c["entry"], flags=capi.SEARCH_RERANK)
    for key in ("ids", "dists", "hops", "dist_calc"):
        assert np.array_equal(g[key], o[key]), key
    o = O.orc_search(c["queries"], None, c["base"], None, goff, ged, ef, ef, 2, c["entry"])
    g = ix.search(c["queries"], None, ef, ef, c["entry"], flags=capi.SEARCH_PLAIN)
    for key in ("ids", "dists", "hops", "dist_calc"):
        assert np.array_equal(g[key], o[key]), ("plain", key)


@pytest.mark.parametrize("lognb", ["3", "5"])
def test_tagged_visited_table_overflow_is_exact(gpu_index_factory, monkeypatch, lognb):
    """A tiny 16-bit-tag table: probe windows fill up (ids diverted to the HBM table one by one) and the
    table closes early (everything diverted); ids, hops and dist_calc must not change."""
    monkeypatch.setenv("GBDR_BEAM_VARIANT", "v2")
    monkeypatch.setenv("GBDR_BEAM_VIS16", "1")
    monkeypatch.setenv("GBDR_BEAM_VIS16_LOGNB", lognb)
    c = small_case()
    ix = _index(gpu_index_factory, c)
    goff, ged = c["graph"]
    for ef in (10, 50, 100, 200):
        o = O.orc_search(c["queries"], c["q_low"], c["base"], c["db_low"], goff, ged, ef, 1, 0, c["entry"])
        g = ix.search(c["queries"], c["q_low"], ef, 1, c["entry"], flags=capi.SEARCH_RERANK)
        assert ix.status() & 1, "expected the HBM visited table to be exercised"
        for key in ("ids", "dists", "hops", "dist_calc"):
            assert np.array_equal(g[key], o[key]), (ef, key)


@pytest.mark.parametrize("llf", [False, True])
@pytest.mark.parametrize("hops_bound", [0, 3, 50])
def test_second_graph_matches_oracle(gpu_index_factory, llf, hops_bound):
    """use_second_graph == true (search_function.h:73-89) for the three performTest branches."""
    c = small_case()
    ix = _index(gpu_index_factory, c)
    goff, ged = c["graph"]
    aux = long_link_graph(c["n"])
    ix.set_aux_graph(*aux, hops_bound=hops_bound, llf=llf)
    for mode, flags, ef, k in ((0, capi.SEARCH_RERANK, 20, 1), (1, 0, 40, 10), (2, capi.SEARCH_PLAIN, 7, 7),
                               (0, capi.SEARCH_RERANK, 300, 5)):
        o = O.orc_search(c["queries"], c["q_low"], c["base"], c["db_low"], goff, ged, ef, k, mode, c["entry"], aux=aux,
                         llf=llf, hops_bound=hops_bound)
        g = ix.search(c["queries"], c["q_low"], ef, k, c["entry"], flags=flags | capi.SEARCH_SECOND_GRAPH)
        for key in ("ids", "dists", "hops", "dist_calc"):
            assert np.array_equal(g[key], o[key]), (mode, ef, key)
    # the flag is per call: without it the same index searches the main graph only
    o = O.orc_search(c["queries"], c["q_low"], c["base"], c["db_low"], goff, ged, 20, 1, 0, c["entry"])
    g = ix.search(c["queries"], c["q_low"], 20, 1, c["entry"], flags=capi.SEARCH_RERANK)
    for key in ("ids", "dists", "hops", "dist_calc"):
        assert np.array_equal(g[key], o[key]), key


def test_second_graph_requires_aux(gpu_index_factory):
    c = small_case()
    ix = _index(gpu_index_factory, c)
    with pytest.raises(capi.GbdrError):
        ix.search(c["queries"], c["q_low"], 20, 1, c["entry"], flags=capi.SEARCH_RERANK | capi.SEARCH_SECOND_GRAPH)
    aux = long_link_graph(c["n"])
    ix.set_aux_graph(*aux)
    ix.search(c["queries"], c["q_low"], 20, 1, c["entry"], flags=capi.SEARCH_RERANK | capi.SEARCH_SECOND_GRAPH)
    ix.set_aux_graph(None, None)
    with pytest.raises(capi.GbdrError):
        ix.search(c["queries"], c["q_low"], 20, 1, c["entry"], flags=capi.SEARCH_RERANK | capi.SEARCH_SECOND_GRAPH)


def test_views_share_the_resident_index(gpu_index_factory):
    """gbdr_index_create_view: same results as the owning handle, read-only, follows parent updates."""
    c = small_case()
    ix = _index(gpu_index_factory, c)
    ix.set_net(*c["net"])
    v = ix.view()
    want = ix.search(c["queries"], c["q_low"], 40, 5, c["entry"], flags=capi.SEARCH_RERANK)
    got = v.search(c["queries"], c["q_low"], 40, 5, c["entry"], flags=capi.SEARCH_RERANK)
    for key in ("ids", "dists", "hops", "dist_calc"):
        assert np.array_equal(got[key], want[key]), key
    # projection through the view's own plan
    a = ix.search(c["queries"], None, 40, 1, c["entry"], flags=capi.SEARCH_RERANK)
    b = v.search(c["queries"], None, 40, 1, c["entry"], flags=capi.SEARCH_RERANK)
    assert np.array_equal(a["ids"], b["ids"]) and np.array_equal(a["hops"], b["hops"])
    assert np.array_equal(ix.project(c["queries"]), v.project(c["queries"]))
    with pytest.raises(capi.GbdrError):
        v.set_low(c["db_low"])
    with pytest.raises(capi.GbdrError):  # live view
        capi._chk(capi.lib().gbdr_index_destroy(ix._h))
    # the parent changes its graph: the view follows at its next call
    koff, ked = c["knn"]
    ix.set_graph(koff, ked)
    want = ix.search(c["queries"], c["q_low"], 20, 1, c["entry"], flags=capi.SEARCH_RERANK)
    got = v.search(c["queries"], c["q_low"], 20, 1, c["entry"], flags=capi.SEARCH_RERANK)
    o = O.orc_search(c["queries"], c["q_low"], c["base"], c["db_low"], koff, ked, 20, 1, 0, c["entry"])
    for key in ("ids", "dists", "hops", "dist_calc"):
        assert np.array_equal(got[key], want[key]) and np.array_equal(got[key], o[key]), key
    v.close()


def test_batches_in_flight_match_blocking_calls(gpu_index_factory):
    """gbdr_search_submit / gbdr_search_wait on the index and two views, pinned buffers, different ef per
    batch: every batch equals the oracle; misuse (double submit, wait without submit) is an error."""
    c = small_case()
    ix = _index(gpu_index_factory, c)
    handles = [ix, ix.view(), ix.view()]
    goff, ged = c["graph"]
    q = capi.pinned_empty(c["queries"].shape, np.float32)
    q[:] = c["queries"]
    ql = capi.pinned_empty(c["q_low"].shape, np.float32)
    ql[:] = c["q_low"]
    en = capi.pinned_empty(c["entry"].shape, np.uint32)
    en[:] = c["entry"]
    efs = [16, 40, 100, 24, 57, 8]
    with pytest.raises(capi.GbdrError):
        ix.search_wait()
    results = {}
    for i, ef in enumerate(efs):
        h = handles[i % 3]
        if h._inflight is not None:
            results[i - 3] = {k: np.array(v) for k, v in h.search_wait().items() if k != "gpu_seconds"}
        h.search_submit(q, ql, ef, 1, en, flags=capi.SEARCH_RERANK)
        if i == 0:
            with pytest.raises(capi.GbdrError):
                capi._chk(capi.lib().gbdr_search_submit(h._h, None, None, 0, 1, 1, 0, capi._ptr(en), capi._ptr(en),
                                                        None, None, None))
    for i in range(len(efs) - 3, len(efs)):
        results[i] = {k: np.array(v) for k, v in handles[i % 3].search_wait().items() if k != "gpu_seconds"}
    for i, ef in enumerate(efs):
        o = O.orc_search(c["queries"], c["q_low"], c["base"], c["db_low"], goff, ged, ef, 1, 0, c["entry"])
        for key in ("ids", "dists", "hops", "dist_calc"):
            assert np.array_equal(results[i][key], o[key]), (i, ef, key)
    for h in handles[1:]:
        h.close()


@pytest.mark.parametrize("variant", ["v2", "reg", "smem"])
def test_wide_and_repeated_adjacency_rows(gpu_index_factory, monkeypatch, variant):
    """Rows of 100 neighbours incl. the vertex itself (the raw kNN graph: two 64-id chunks per row), rows of exactly
    32 / 33 / 64 / 65 ids, and rows that name a neighbour twice (the reference skips the repeat as visited)."""
    monkeypatch.setenv("GBDR_BEAM_VARIANT", variant)
    c = small_case()
    n = c["n"]
    knn = c["knn_ids"]
    rng = np.random.default_rng(8)
    lists = []
    for i in range(n):
        deg = (32, 33, 64, 65, 100, 7)[i % 6]
        row = list(knn[i, :deg])
        if i % 5 == 0:  # repeats: the 3rd id again right away, and the 1st at the end
            row = row[:3] + [row[2]] + row[3:] + [row[0]]
        lists.append(row)
    from gbnns_dim_red_b200 import xvecs

    off, ed = xvecs.adjacency_from_lists(lists)
    ix = gpu_index_factory()
    ix.set_base(c["base"])
    ix.set_low(c["db_low"])
    ix.set_graph(off, ed)
    for ef, k, mode, flags in ((30, 1, 0, capi.SEARCH_RERANK), (60, 10, 1, 0), (100, 5, 0, capi.SEARCH_RERANK)):
        o = O.orc_search(c["queries"], c["q_low"], c["base"], c["db_low"], off, ed, ef, k, mode, c["entry"])
        g = ix.search(c["queries"], c["q_low"], ef, k, c["entry"], flags=flags)
        for key in ("ids", "dists", "hops", "dist_calc"):
            assert np.array_equal(g[key], o[key]), (variant, ef, key)


def test_random_small_graphs_property(gpu_index_factory):
    """Randomised (seeded) GPU-vs-oracle check on arbitrary small digraphs: self loops, repeated neighbours, vertices
    without out-edges, integer-grid vectors (many exact ties), random ef / k / mode, with and without a second graph
    (the oracle itself is pinned on the same generator against the reference, tests/test_oracle_vs_reference.py)."""
    from gbnns_dim_red_b200 import xvecs

    rng = np.random.default_rng(2024)
    ix = gpu_index_factory()
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
                lists.append(rng.integers(0, n, size=deg).astype(np.uint32).tolist())
            return xvecs.adjacency_from_lists(lists)

        off, ed = rand_graph(6)
        aux = rand_graph(3) if trial % 3 == 0 else None
        entry = rng.integers(0, n, size=n_q, dtype=np.uint32)
        ef = int(rng.integers(1, 12))
        k = int(rng.integers(1, ef + 1))
        mode = int(rng.integers(0, 3))
        llf, hb = bool(trial % 2), int(rng.integers(0, 5))
        kw = dict(aux=aux, llf=llf, hops_bound=hb) if aux is not None else {}
        ix.set_base(base)
        ix.set_low(base)
        ix.set_graph(off, ed)
        flags = {0: capi.SEARCH_RERANK, 1: 0, 2: capi.SEARCH_PLAIN}[mode]
        if aux is not None:
            ix.set_aux_graph(*aux, hops_bound=hb, llf=llf)
            flags |= capi.SEARCH_SECOND_GRAPH
        o = O.orc_search(queries, queries, base, base, off, ed, ef, k, mode, entry, **kw)
        g = ix.search(queries, queries, ef, k, entry, flags=flags)
        for key in ("ids", "dists", "hops", "dist_calc"):
            assert np.array_equal(g[key], o[key]), (trial, n, d, ef, k, mode, aux is not None, key)


def test_bad_entry_ids_fail_loudly(gpu_index_factory):
    """An entry id that is not a vertex: the host call refuses it before anything runs; the device-pointer call marks
    the query failed (PAD results, status bit 4) instead of reading out of bounds, and the other queries are intact."""
    c = small_case()
    ix = _index(gpu_index_factory, c)
    goff, ged = c["graph"]
    n, n_q = c["n"], 64
    entry = c["entry"][:n_q].copy()
    entry[5] = n
    with pytest.raises(capi.GbdrError) as ei:
        ix.search(c["queries"][:n_q], c["q_low"][:n_q], 20, 1, entry)
    assert ei.value.code == -1 and "entry[5]" in str(ei.value)
    # device-pointer path
    entry[9] = 0xFFFFFFF0
    bufs = {}
    for name, arr in (("q", c["queries"][:n_q]), ("ql", c["q_low"][:n_q]), ("entry", entry)):
        bufs[name] = capi.DeviceBuffer(arr.nbytes)
        bufs[name].upload(np.ascontiguousarray(arr))
    for name, nb in (("ids", n_q * 3 * 4), ("dists", n_q * 3 * 4), ("hops", n_q * 4), ("dc", n_q * 4)):
        bufs[name] = capi.DeviceBuffer(nb)
    ix.search_dev(bufs["q"].ptr, bufs["ql"].ptr, n_q, 20, 3, bufs["entry"].ptr, bufs["ids"].ptr, bufs["dists"].ptr,
                  bufs["hops"].ptr, bufs["dc"].ptr, flags=capi.SEARCH_RERANK, stream=ix.stream())
    assert ix.status() & 16
    ids = bufs["ids"].download((n_q, 3), np.uint32)
    good = np.ones(n_q, bool)
    good[[5, 9]] = False
    assert (ids[~good] == 0xFFFFFFFF).all()
    ok_entry = c["entry"][:n_q].copy()
    o = O.orc_search(c["queries"][:n_q], c["q_low"][:n_q], c["base"], c["db_low"], goff, ged, 20, 3, 0, ok_entry)
    assert np.array_equal(ids[good], o["ids"][good])
    for b in bufs.values():
        b.free()


def test_overflow_tables_grow_on_demand(gpu_index_factory, monkeypatch):
    """The HBM overflow tables start at the size the beam width needs; a call that exhausts them is re-run by
    gbdr_search_wait with the largest tables, transparently and exactly."""
    c = small_case()
    ix = _index(gpu_index_factory, c)
    goff, ged = c["graph"]
    # a 4-bucket tag table diverts nearly every visited id to HBM, and 256-slot overflow tables close at 192 ids
    monkeypatch.setenv("GBDR_BEAM_VIS16_LOGNB", "2")
    monkeypatch.setenv("GBDR_BEAM_SPILL_LOG", "8")
    o = O.orc_search(c["queries"], c["q_low"], c["base"], c["db_low"], goff, ged, 300, 1, 0, c["entry"])
    g = ix.search(c["queries"], c["q_low"], 300, 1, c["entry"], flags=capi.SEARCH_RERANK)
    for key in ("ids", "dists", "hops", "dist_calc"):
        assert np.array_equal(g[key], o[key]), key


@pytest.mark.parametrize("mode", ["rerank_precomputed", "rerank_projected", "plain"])
def test_repeated_calls_on_pinned_buffers_replay_a_graph(gpu_index_factory, mode):
    """A serving loop: the same call shape on the same page-locked buffers, new queries in them every time.  The library
    captures the second such call into a CUDA graph and replays it afterwards (gbdr_search_submit): every repetition must
    return what the plain path returns for THAT repetition's queries (the oracle), a change of shape in between falls back
    to the plain path, and the per-call device time stays available."""
    c = small_case()
    ix = _index(gpu_index_factory, c)
    if mode == "rerank_projected":
        ix.set_net(*c["net"])
    n_q, d, d_low, ef, k = c["n_q"], c["d"], c["d_low"], 30, 5
    goff, ged = c["graph"]
    q = capi.pinned_empty((n_q, d), np.float32)
    ql = capi.pinned_empty((n_q, d_low), np.float32)
    en = capi.pinned_empty((n_q,), np.uint32)
    out = dict(ids=capi.pinned_empty((n_q, k), np.uint32), dists=capi.pinned_empty((n_q, k), np.float32),
               hops=capi.pinned_empty((n_q,), np.int32), dist_calc=capi.pinned_empty((n_q,), np.int32))
    rng = np.random.default_rng(3)
    flags = capi.SEARCH_PLAIN if mode == "plain" else capi.SEARCH_RERANK
    for rep in range(6):
        perm = rng.permutation(n_q)
        q[:] = c["queries"][perm]
        ql[:] = c["q_low"][perm]
        en[:] = c["entry"][perm]
        if rep == 3:  # another shape in between: drops the graph, the next two calls rebuild it
            other = ix.search(c["queries"][:50], c["q_low"][:50], ef + 7, 1, c["entry"][:50], flags=capi.SEARCH_RERANK)
            o = O.orc_search(c["queries"][:50], c["q_low"][:50], c["base"], c["db_low"], goff, ged, ef + 7, 1, 0, c["entry"][:50])
            assert np.array_equal(other["ids"], o["ids"])
        got = ix.search(q, None if mode == "rerank_projected" else (ql if mode != "plain" else None), ef, k, en, flags=flags, out=out)
        assert got["gpu_seconds"] > 0
        if mode == "plain":
            want = O.orc_search(q, None, c["base"], None, goff, ged, ef, k, 2, en)
        else:
            want = O.orc_search(q, ql, c["base"], c["db_low"], goff, ged, ef, k, 0, en)
        if mode == "rerank_projected":  # 3xTF32 projection: a query or two may walk differently than with the oracle's q_low
            assert (got["ids"][:, 0] == want["ids"][:, 0]).mean() >= 0.98
        else:
            for key in ("ids", "dists", "hops", "dist_calc"):
                assert np.array_equal(got[key], want[key]), (rep, key)
    ms = ix.last_kernel_ms()
    assert ms["search"] > 0
