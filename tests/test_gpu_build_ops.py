"""GPU parity: projection (K1), kNN build (K4), top-k merge (K5) through the C ABI."""
import numpy as np
import pytest

from gbnns_dim_red_b200 import capi, synth

from . import _oracle as O
from ._data import small_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode,tol", [(capi.PROJ_FP32, 2e-6), (capi.PROJ_3XTF32, 1e-5), (capi.PROJ_TF32, 5e-3)])
def test_projection_matches_oracle(gpu_index_factory, mode, tol):
    """fp tolerance (north_star: 1e-5 relative, looser stated bound for single-pass TF32):
    outputs are unit vectors, so the bound is absolute on each component."""
    c = small_case()
    ix = gpu_index_factory()
    ix.set_net(*c["net"])
    ix.set_projection_mode(mode)
    got = ix.project(c["queries"])
    want = O.orc_project(*c["net"], c["queries"])
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= tol
    exact = synth.project_numpy(*c["net"], c["queries"])
    assert np.abs(got - exact).max() <= tol


@pytest.mark.parametrize("shape", [(128, 256, 256, 32, 1000), (96, 128, 128, 16, 777), (960, 1024, 1024, 32, 300)])
def test_projection_baseline_shapes(gpu_index_factory, shape):
    d, dh, dh2, dl, nq = shape
    rng = np.random.default_rng(5)
    q = rng.standard_normal((nq, d), dtype=np.float32)
    net = synth.make_net(d, dh, dl, seed=3, d_hidden2=dh2)
    ix = gpu_index_factory()
    ix.set_net(*net)
    exact = synth.project_numpy(*net, q)
    for mode, tol in ((capi.PROJ_FP32, 2e-6), (capi.PROJ_3XTF32, 1e-5)):
        ix.set_projection_mode(mode)
        got = ix.project(q)
        assert np.abs(got - exact).max() <= tol, (mode, np.abs(got - exact).max())


@pytest.mark.parametrize("shape", [(100, 72, 40, 20, 131), (33, 300, 64, 7, 1), (64, 32, 32, 256, 129), (128, 256, 256, 32, 10000)])
def test_projection_ragged_shapes(gpu_index_factory, shape):
    """K not a multiple of 32, hidden widths that need N padding, d_low % 4 != 0 (normalizeVector
    ignores the tail, support_func.h:636-642), single rows, row counts that leave a partial tile."""
    d, dh, dh2, dl, nq = shape
    rng = np.random.default_rng(d + nq)
    q = rng.standard_normal((nq, d), dtype=np.float32)
    net = synth.make_net(d, dh, dl, seed=8, d_hidden2=dh2)
    ix = gpu_index_factory()
    ix.set_net(*net)
    want = O.orc_project(*net, q)
    for mode, tol in ((capi.PROJ_FP32, 2e-6), (capi.PROJ_3XTF32, 2e-6), (capi.PROJ_TF32, 5e-3)):
        ix.set_projection_mode(mode)
        got = ix.project(q)
        assert np.isfinite(got).all()
        assert np.abs(got - want).max() <= tol, (mode, float(np.abs(got - want).max()))


@pytest.mark.parametrize("n,d,k", [(3000, 16, 100), (2500, 32, 64), (1000, 128, 10), (700, 96, 700), (130, 960, 5)])
def test_knn_self_matches_oracle(n, d, k, gpu_index_factory):
    rng = np.random.default_rng(n + d)
    B = rng.standard_normal((n, d), dtype=np.float32)
    ids, dists, _ = capi.knn(B, B, k, return_dists=True)
    oi, od = O.orc_knn(B, B, k)
    assert np.array_equal(ids, oi)
    assert np.array_equal(dists, od)
    assert (ids[:, 0] == np.arange(n)).all()  # self at rank 0


def test_knn_queries_vs_base_and_ties(gpu_index_factory):
    c = small_case()
    base = np.concatenate([c["base"][:900], c["base"][:300]])  # exact duplicates -> (dist,id) ties
    ids, dists, _ = capi.knn(c["queries"], base, 20, return_dists=True)
    oi, od = O.orc_knn(c["queries"], base, 20)
    assert np.array_equal(ids, oi)
    assert np.array_equal(dists, od)


def test_knn_k_larger_than_n_pads(gpu_index_factory):
    rng = np.random.default_rng(0)
    B = rng.standard_normal((50, 16), dtype=np.float32)
    ids, _ = capi.knn(B, B, 64)
    oi, _ = O.orc_knn(B, B, 64)
    assert np.array_equal(ids, oi)
    assert (ids[:, 50:] == capi.PAD_ID).all()


def test_merge_topk(gpu_index_factory):
    rng = np.random.default_rng(11)
    parts, n_q, k_in, k_out = 4, 37, 25, 30
    d = np.sort(rng.random((parts, n_q, k_in), dtype=np.float32), axis=2)
    d[:, :, ::5] = np.round(d[:, :, ::5], 1)  # many exact ties across parts
    d = np.sort(d, axis=2)
    ids = rng.permutation(parts * n_q * k_in).astype(np.uint32).reshape(parts, n_q, k_in)
    # make lists ascending by (dist, id)
    for p in range(parts):
        for q in range(n_q):
            order = np.lexsort((ids[p, q], d[p, q]))
            ids[p, q] = ids[p, q][order]
            d[p, q] = d[p, q][order]
    ids[3, :, 20:] = capi.PAD_ID  # short lists
    d[3, :, 20:] = np.inf
    b_ids = capi.DeviceBuffer(ids.nbytes).upload(ids)
    b_d = capi.DeviceBuffer(d.nbytes).upload(d)
    o_ids = capi.DeviceBuffer(n_q * k_out * 4)
    o_d = capi.DeviceBuffer(n_q * k_out * 4)
    capi.merge_topk_dev(0, b_ids.ptr, b_d.ptr, parts, n_q, k_in, k_out, o_ids.ptr, o_d.ptr)
    capi.synchronize(0)
    got_i = o_ids.download((n_q, k_out), np.uint32)
    got_d = o_d.download((n_q, k_out), np.float32)
    for q in range(n_q):
        ai = ids[:, q].reshape(-1)
        ad = d[:, q].reshape(-1)
        keep = ai != capi.PAD_ID
        ai, ad = ai[keep], ad[keep]
        order = np.lexsort((ai, ad))[:k_out]
        assert np.array_equal(got_i[q, : order.size], ai[order])
        assert np.array_equal(got_d[q, : order.size], ad[order])


@pytest.mark.parametrize("M", [3, 4, 6, 10])
def test_gd_prune_reverse_pass_with_full_rows(gpu_index_factory, M):
    from gbnns_dim_red_b200 import xvecs

    from ._data import hub_points

    x = hub_points()
    ids, _ = O.orc_knn(x, x, 60)
    koff, ked = xvecs.adjacency_from_matrix(ids)
    ooff, oed = O.orc_gd_prune(koff, ked, x, M=M, reverse=True)
    assert (np.diff(ooff.astype(np.int64)) == 2 * M).sum() >= 50, "the case is meant to fill rows"
    off, ed, _ = capi.gd_prune(koff, ked, x, M=M, reverse=True)
    assert np.array_equal(off, ooff)
    assert np.array_equal(ed, oed)
    off, ed, _ = capi.gd_prune(koff, ked, x, M=M, reverse=True, need_const_degree=True)
    ooff, oed = O.orc_gd_prune(koff, ked, x, M=M, reverse=True, const_degree=True)
    assert np.array_equal(off, ooff) and np.array_equal(ed, oed)


@pytest.mark.parametrize("M,reverse,const", [(12, True, False), (30, True, False), (8, False, False), (10, True, True),
                                             (4, True, False), (44, True, False)])
def test_gd_prune_matches_oracle(gpu_index_factory, M, reverse, const):
    """hnswlikeGD: neighbour ids identical to the oracle (= the reference's strict build)."""
    c = small_case()
    koff, ked = c["knn"]
    off, ed, _ = capi.gd_prune(koff, ked, c["db_low"], M=M, reverse=reverse, need_const_degree=const)
    ooff, oed = O.orc_gd_prune(koff, ked, c["db_low"], M=M, reverse=reverse, const_degree=const)
    assert np.array_equal(off, ooff)
    assert np.array_equal(ed, oed)


def test_gd_prune_with_duplicates_and_ragged_lists(gpu_index_factory):
    """dist <= eps candidates (self, duplicates) are dropped (support_func.h:535); ragged lists."""
    c = small_case()
    low = np.concatenate([c["db_low"][:500], c["db_low"][:100]])
    ids, _ = O.orc_knn(low, low, 40)
    lists = [list(ids[i, : 40 - (i % 7)]) for i in range(low.shape[0])]
    from gbnns_dim_red_b200 import xvecs

    koff, ked = xvecs.adjacency_from_lists(lists)
    off, ed, _ = capi.gd_prune(koff, ked, low, M=10, reverse=True)
    ooff, oed = O.orc_gd_prune(koff, ked, low, M=10, reverse=True)
    assert np.array_equal(off, ooff)
    assert np.array_equal(ed, oed)


@pytest.mark.parametrize("n,nq,d,k", [(40000, 1500, 32, 100), (33000, 700, 16, 64), (50000, 300, 128, 10),
                                      (300000, 300, 32, 1000), (32768, 129, 96, 1), (200000, 256, 64, 500)])
def test_knn_tensor_core_path_is_exact(n, nq, d, k):
    """n >= 32768, d <= 128, k <= 1024 and k/n small takes the tcgen05 filter + exact recompute path (knn_tc.cu): ids and
    distances must equal the oracle's exact (dist,id) ranking, self included at rank 0."""
    rng = np.random.default_rng(n + d + k)
    lat = rng.standard_normal((n, 6), dtype=np.float32) @ rng.standard_normal((6, d), dtype=np.float32)
    B = (lat + 0.05 * rng.standard_normal((n, d), dtype=np.float32)).astype(np.float32)
    rows = rng.choice(n, size=nq, replace=False)
    Q = np.ascontiguousarray(B[rows])
    ids, dists, _ = capi.knn(Q, B, k, return_dists=True)
    oi, od = O.orc_knn(Q, B, k)
    assert np.array_equal(ids, oi)
    assert np.array_equal(dists, od)
    assert np.array_equal(ids[:, 0], rows.astype(np.uint32))


def test_knn_tensor_core_unit_vectors_and_duplicates():
    """Unit-norm rows (what the projection net emits) with every vector duplicated: exact zero
    distances and (dist,id) ties inside the top-k."""
    rng = np.random.default_rng(77)
    x = rng.standard_normal((20000, 32), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    B = np.concatenate([x, x]).astype(np.float32)
    Q = np.ascontiguousarray(B[:500])
    ids, dists, _ = capi.knn(Q, B, 50, return_dists=True)
    oi, od = O.orc_knn(Q, B, 50)
    assert np.array_equal(ids, oi) and np.array_equal(dists, od)


def test_knn_tensor_core_massive_ties_fall_back_to_exact_scan():
    """Integer grid: thousands of equal distances make the candidate set unboundable for the filter;
    those rows are redone by the exact scan kernel and must still match."""
    rng = np.random.default_rng(78)
    B = rng.integers(0, 2, size=(40000, 16)).astype(np.float32)
    Q = np.ascontiguousarray(B[:64])
    ids, dists, _ = capi.knn(Q, B, 100, return_dists=True)
    oi, od = O.orc_knn(Q, B, 100)
    assert np.array_equal(ids, oi) and np.array_equal(dists, od)


def test_knn_variants_agree(monkeypatch):
    rng = np.random.default_rng(79)
    B = rng.standard_normal((40000, 32), dtype=np.float32)
    a, _ = capi.knn(B[:300], B, 20)
    monkeypatch.setenv("GBDR_KNN_VARIANT", "scan")
    b, _ = capi.knn(B[:300], B, 20)
    assert np.array_equal(a, b)


def test_knn_tensor_core_multi_chunk_self_join():
    """More query rows than one chunk of the tensor-core pipeline (75 776): the second chunk reuses the candidate
    buffers and the result chunks are streamed to the host while it is computed."""
    rng = np.random.default_rng(80)
    n, d, k = 80000, 16, 24
    lat = rng.standard_normal((n, 5), dtype=np.float32) @ rng.standard_normal((5, d), dtype=np.float32)
    B = (lat + 0.05 * rng.standard_normal((n, d), dtype=np.float32)).astype(np.float32)
    ids, dists, _ = capi.knn(B, B, k, return_dists=True)
    rows = np.concatenate([np.arange(0, 600), np.arange(75700, 75900), np.arange(n - 400, n)])
    oi, od = O.orc_knn(np.ascontiguousarray(B[rows]), B, k)
    assert np.array_equal(ids[rows], oi)
    assert np.array_equal(dists[rows], od)
    assert np.array_equal(ids[:, 0], np.arange(n, dtype=np.uint32))


def test_knn_tensor_core_scattered_rows_redone_together():
    """A self-join in which a few percent of the rows cannot be bounded by the filter (a clump of 6000 identical points:
    every member has 6000 candidates at distance 0, more than a row buffer holds): those rows are gathered and redone by
    ONE exact scan, the rest keep their tensor-core results.  Small k on many rows also takes the larger-sample branch
    of the threshold estimate."""
    rng = np.random.default_rng(81)
    n, d, k = 90000, 16, 12
    lat = rng.standard_normal((n, 5), dtype=np.float32) @ rng.standard_normal((5, d), dtype=np.float32)
    B = (lat + 0.05 * rng.standard_normal((n, d), dtype=np.float32)).astype(np.float32)
    clump = rng.choice(n, size=6000, replace=False)
    B[clump] = B[clump[0]]
    ids, dists, _ = capi.knn(B, B, k, return_dists=True)
    rows = np.concatenate([np.sort(clump)[:150], np.sort(clump)[-150:], rng.choice(n, size=400, replace=False)])
    oi, od = O.orc_knn(np.ascontiguousarray(B[rows]), B, k)
    assert np.array_equal(ids[rows], oi)
    assert np.array_equal(dists[rows], od)


@pytest.mark.parametrize("knn_size", [1, 32, 100, 150])
def test_knn_cut_matches_oracle(knn_size):
    """gbdr_knn_cut = cutKNNbyK (support_func.h:309-340) on shuffled, ragged lists, in the low and the original dimension."""
    c = small_case()
    koff, ked = c["knn"]
    rng = np.random.default_rng(knn_size)
    rows = [rng.permutation(ked[int(koff[i]):int(koff[i + 1])]) for i in range(koff.size - 1)]
    for i in rng.integers(0, len(rows), 40):
        rows[i] = rows[i][: rng.integers(0, 9)]
    off = np.zeros(len(rows) + 1, np.uint64)
    off[1:] = np.cumsum([r.size for r in rows])
    edges = np.concatenate(rows).astype(np.uint32)
    for db in (c["db_low"], c["base"]):
        goff, ged, _ = capi.knn_cut(off, edges, db, knn_size)
        ooff, oed = O.orc_knn_cut(off, edges, db, knn_size)
        assert np.array_equal(goff, ooff) and np.array_equal(ged, oed)
