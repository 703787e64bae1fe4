"""gbdr_group_* (several GPUs of one node behind the C ABI) and the device-resident graph build, against the CPU oracle.

A group may list one device several times, so the sharded search with its fused peer-load merge and the row-block
sharded build are exercised on a one-GPU box too; with >= 2 GPUs the same tests also run over distinct devices and
compare the two exchange steps (NVLink peer loads vs ncclAllGather + merge kernel)."""
import numpy as np
import pytest

from gbnns_dim_red_b200 import capi, xvecs

from . import _oracle as O
from ._data import hub_points, small_case

pytestmark = pytest.mark.gpu


def _device_sets():
    n = capi.device_count()
    sets = [[0], [0, 0], [0, 0, 0]]
    if n >= 2:
        sets += [[0, 1]]
    if n >= 4:
        sets += [[0, 1, 2, 3]]
    return sets


def _partition(n, world):
    base, rem = divmod(n, world)
    out, b = [], 0
    for r in range(world):
        e = b + base + (1 if r < rem else 0)
        out.append((b, e))
        b = e
    return out


@pytest.mark.parametrize("devs", _device_sets())
def test_replicated_group_equals_one_index(devs):
    c = small_case()
    g = capi.Group(devs, capi.GROUP_REPLICATED)
    try:
        g.set_net(*c["net"])
        g.set_base(c["base"])
        g.set_low(c["db_low"])
        g.set_graph(*c["graph"])
        goff, ged = c["graph"]
        for ef, k, flags, mode, ql in ((40, 1, capi.SEARCH_RERANK, 0, c["q_low"]), (25, 5, 0, 1, c["q_low"]),
                                       (20, 3, capi.SEARCH_PLAIN, 2, None)):
            o = O.orc_search(c["queries"], c["q_low"], c["base"], c["db_low"], goff, ged, ef, k, mode, c["entry"])
            r = g.search(c["queries"], ql, ef, k, c["entry"], flags=flags)
            for key in ("ids", "dists", "hops", "dist_calc"):
                assert np.array_equal(r[key], o[key]), (devs, ef, key)
        # on-the-fly projection (performNetTest): same answers as the single index within projection rounding
        ix = capi.Index(0)
        ix.set_net(*c["net"]); ix.set_base(c["base"]); ix.set_low(c["db_low"]); ix.set_graph(goff, ged)
        a = ix.search(c["queries"], None, 40, 1, c["entry"], flags=capi.SEARCH_RERANK)
        b = g.search(c["queries"], None, 40, 1, c["entry"], flags=capi.SEARCH_RERANK)
        ix.close()
        assert np.array_equal(a["ids"], b["ids"]) and np.array_equal(a["dists"], b["dists"])
    finally:
        g.close()


def _shard_case(c, world, M=10, knn_k=40):
    """Per-shard low-dim kNN + GD graphs over LOCAL ids (built with the oracle), local entry points."""
    parts = _partition(c["n"], world)
    graphs, entry = [], np.empty((world, c["n_q"]), np.uint32)
    rng = np.random.default_rng(3)
    for r, (b, e) in enumerate(parts):
        low = c["db_low"][b:e]
        ids, _ = O.orc_knn(low, low, knn_k)
        koff, ked = xvecs.adjacency_from_matrix(ids)
        graphs.append(O.orc_gd_prune(koff, ked, low, M=M, reverse=True))
        entry[r] = rng.integers(0, e - b, size=c["n_q"], dtype=np.uint32)
    return parts, graphs, entry


def _oracle_sharded(c, parts, graphs, entry, ef, k):
    """Reference semantics per shard + (dist, id) merge of the per-shard top-k lists."""
    ids, dists, hops, dc = [], [], 0, 0
    for r, (b, e) in enumerate(parts):
        o = O.orc_search(c["queries"], c["q_low"], c["base"][b:e], c["db_low"][b:e], graphs[r][0], graphs[r][1], ef, k, 0, entry[r])
        gi = o["ids"].astype(np.int64)
        gi[o["ids"] != capi.PAD_ID] += b
        gi[o["ids"] == capi.PAD_ID] = capi.PAD_ID
        ids.append(gi)
        dists.append(np.where(o["ids"] == capi.PAD_ID, np.inf, o["dists"]))
        hops = hops + o["hops"]
        dc = dc + o["dist_calc"]
    ids, dists = np.concatenate(ids, axis=1), np.concatenate(dists, axis=1)
    order = np.lexsort((ids, dists), axis=1)[:, :k]
    return np.take_along_axis(ids, order, 1).astype(np.uint32), np.take_along_axis(dists, order, 1).astype(np.float32), hops, dc


@pytest.mark.parametrize("devs", [d for d in _device_sets() if len(d) > 1])
def test_sharded_group_equals_per_shard_oracle_and_merge(devs):
    c = small_case()
    world = len(devs)
    parts, graphs, entry = _shard_case(c, world)
    g = capi.Group(devs, capi.GROUP_SHARDED)
    try:
        g.set_net(*c["net"])
        g.set_base(c["base"])
        g.set_low(c["db_low"])
        for r in range(world):
            assert g.shard_rows(r) == parts[r]
            g.set_shard_graph(r, *graphs[r])
        exchanges = [capi.EXCHANGE_PEER] + ([capi.EXCHANGE_NCCL] if len(set(devs)) == world else [])
        for ex in exchanges:
            g.set_exchange(ex)
            for ef, k in ((30, 10), (12, 1), (64, 64)):
                want_ids, want_d, want_h, want_dc = _oracle_sharded(c, parts, graphs, entry, ef, k)
                r = g.search(c["queries"], c["q_low"], ef, k, entry, flags=capi.SEARCH_RERANK)
                assert np.array_equal(r["ids"], want_ids), (devs, ex, ef, k)
                assert np.array_equal(r["dists"], want_d)
                assert np.array_equal(r["hops"], want_h) and np.array_equal(r["dist_calc"], want_dc)
    finally:
        g.close()


def test_sharded_group_rejects_a_whole_graph_and_bad_entries():
    c = small_case()
    g = capi.Group([0, 0], capi.GROUP_SHARDED)
    try:
        g.set_base(c["base"])
        g.set_low(c["db_low"])
        with pytest.raises(capi.GbdrError):
            g.set_graph(*c["graph"])
        parts, graphs, entry = _shard_case(c, 2)
        for r in range(2):
            g.set_shard_graph(r, *graphs[r])
        bad = entry.copy()
        bad[1, 7] = parts[1][1]  # a GLOBAL id where a local one is expected: out of the shard's range
        with pytest.raises(capi.GbdrError) as ei:
            g.search(c["queries"], c["q_low"], 20, 5, bad, flags=capi.SEARCH_RERANK)
        assert ei.value.code == -1
        # the group still works afterwards
        r = g.search(c["queries"], c["q_low"], 20, 5, entry, flags=capi.SEARCH_RERANK)
        assert (r["ids"][:, 0] != capi.PAD_ID).all()
    finally:
        g.close()


@pytest.mark.parametrize("M,reverse,const", [(12, True, False), (8, False, False), (10, True, True), (44, True, False)])
def test_build_graph_in_hbm_equals_oracle(M, reverse, const):
    """kNN self-join + hnswlikeGD chained on the device: same graph (and same kNN lists) as the two oracle steps."""
    c = small_case()
    low = c["db_low"]
    k = 100
    oi, _ = O.orc_knn(low, low, k)
    koff, ked = xvecs.adjacency_from_matrix(oi)
    ooff, oed = O.orc_gd_prune(koff, ked, low, M=M, reverse=reverse, const_degree=const)
    knn_out = np.empty((low.shape[0], k), np.uint32)
    off, ed, t = capi.build_graph(low, knn_k=k, M=M, reverse=reverse, need_const_degree=const, knn_out=knn_out)
    assert np.array_equal(knn_out, oi)
    assert np.array_equal(off, ooff) and np.array_equal(ed, oed)
    assert all(v >= 0 for v in t.values())


@pytest.mark.parametrize("devs", [d for d in _device_sets() if len(d) > 1])
@pytest.mark.parametrize("const", [False, True])
def test_group_build_graph_equals_oracle(devs, const):
    """Row-block sharded build: all-gather of the vectors, per-block kNN and forward prune, reverse pass on the first
    device; hub-heavy data so that the order-dependent part of the reverse pass matters."""
    x = hub_points()
    k, M = 60, 6
    oi, _ = O.orc_knn(x, x, k)
    koff, ked = xvecs.adjacency_from_matrix(oi)
    ooff, oed = O.orc_gd_prune(koff, ked, x, M=M, reverse=True, const_degree=const)
    g = capi.Group(devs, capi.GROUP_REPLICATED)
    try:
        for ex in [capi.EXCHANGE_PEER] + ([capi.EXCHANGE_NCCL] if len(set(devs)) == len(devs) else []):
            g.set_exchange(ex)
            knn_out = np.empty((x.shape[0], k), np.uint32)
            off, ed, t = g.build_graph(x, knn_k=k, M=M, reverse=True, need_const_degree=const, knn_out=knn_out)
            assert np.array_equal(knn_out, oi), (devs, ex)
            assert np.array_equal(off, ooff) and np.array_equal(ed, oed), (devs, ex)
    finally:
        g.close()


def test_forward_prune_by_row_blocks_on_device_buffers():
    """gbdr_gd_prune_dev over two row blocks + gbdr_gd_finish_dev == gbdr_gd_prune."""
    c = small_case()
    low, (koff, ked), M = c["db_low"], c["knn"], 12
    n, k = low.shape[0], int(koff[1] - koff[0])
    knn = ked.reshape(n, k)
    want_off, want_ed, _ = capi.gd_prune(koff, ked, low, M=M, reverse=True)
    d_low = capi.DeviceBuffer(low.nbytes).upload(low)
    d_fwd, d_deg = capi.DeviceBuffer(n * 2 * M * 4), capi.DeviceBuffer(n * 4)
    cut = 1234
    for b, e in ((0, cut), (cut, n)):
        d_knn = capi.DeviceBuffer((e - b) * k * 4).upload(np.ascontiguousarray(knn[b:e]))
        capi.gd_prune_dev(0, d_knn.ptr, k, k, b, e, d_low.ptr, n, low.shape[1], M, d_fwd.ptr + b * 2 * M * 4, d_deg.ptr + b * 4)
        capi.synchronize(0)
        d_knn.free()
    off, ed = capi.gd_finish_dev(0, d_fwd.ptr, d_deg.ptr, n, M, reverse=True)
    assert np.array_equal(off, want_off) and np.array_equal(ed, want_ed)
    for b in (d_low, d_fwd, d_deg):
        b.free()
