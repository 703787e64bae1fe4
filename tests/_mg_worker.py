"""Worker of tests/test_multigpu_host.py: runs under torch.distributed.run with the gloo backend (CPU)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gbnns_dim_red_b200 import multigpu as mg  # noqa: E402

PAD = 0xFFFFFFFF


def numpy_merge(ids, dists, k_out):
    """Checker (tests only): k-way merge by (dist, id) of [parts, n_q, k] lists."""
    ids, dists = ids.numpy().astype(np.int64) & 0xFFFFFFFF, dists.numpy()
    parts, n_q, k = ids.shape
    out_i = np.full((n_q, k_out), PAD, np.int64)
    out_d = np.full((n_q, k_out), np.inf, np.float32)
    for q in range(n_q):
        cand = [(dists[p, q, j], ids[p, q, j]) for p in range(parts) for j in range(k) if ids[p, q, j] != PAD]
        cand.sort()
        for j, (dd, ii) in enumerate(cand[:k_out]):
            out_i[q, j], out_d[q, j] = ii, dd
    return torch.from_numpy(out_i), torch.from_numpy(out_d)


class FakeIndex:
    """Stands in for capi.Index: 'searching' query i with entry e returns ids (e, e+1, ...)."""

    def search(self, q, ql, ef, k, entry, flags=0):
        ids = entry[:, None].astype(np.uint32) + np.arange(k, dtype=np.uint32)[None, :]
        return dict(ids=ids)


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    assert mg.world_info() == (rank, world)

    # ---- replicated: contiguous query partition, gather restores batch order
    n_q = 11
    entry = np.arange(100, 100 + n_q, dtype=np.uint32)
    rs = mg.ReplicatedSearcher(FakeIndex())
    b, e, r = rs.search(None, None, 4, 3, entry, flags=0)
    assert (b, e) == mg.partition(n_q, world, rank) and r["ids"].shape == (e - b, 3)
    full = rs.gather(b, e, r["ids"], n_q)
    want = entry[:, None] + np.arange(3, dtype=np.uint32)[None, :]
    assert np.array_equal(full, want), (full, want)

    # ---- sharded: every rank searches all queries on its rows; gathered lists merge to the global top-k
    n_total, n_q, k = 1000, 7, 5
    rng = np.random.default_rng(3)  # same on every rank
    table = rng.random((n_q, n_total)).astype(np.float32)  # "distance" of query q to row j
    table[:, ::50] = table[:, 1::50]                        # some exact ties across shards

    def search_local(k):
        rb, re = mg.partition(n_total, world, rank)
        loc = table[:, rb:re]
        order = np.lexsort((np.arange(rb, re)[None, :].repeat(n_q, 0), loc), axis=1)[:, :k]
        ids = (order + rb).astype(np.int64)
        dd = np.take_along_axis(loc, order, axis=1)
        if rank == world - 1:  # a short list: PAD / +inf tail
            ids[:, -1], dd[:, -1] = PAD, np.inf
        return torch.from_numpy(ids), torch.from_numpy(dd)

    ss = mg.ShardedSearcher(search_local, numpy_merge, n_total)
    assert (ss.row_begin, ss.row_end) == mg.partition(n_total, world, rank)
    ids, dd = ss.search(k, k)
    # expected: global (dist,id) order, minus the element the last shard dropped
    lb, le = mg.partition(n_total, world, world - 1)
    for q in range(n_q):
        pairs = sorted((table[q, j], j) for j in range(n_total))
        lastshard = sorted((table[q, j], j) for j in range(lb, le))
        dropped = lastshard[k - 1]
        allowed = [p for p in pairs if not (lb <= p[1] < le) or p < dropped]
        assert [int(x) for x in ids[q]] == [p[1] for p in allowed[:k]], (q, ids[q], allowed[:k])

    # ---- ragged row gather (kNN row blocks)
    counts = [pe - pb for pb, pe in mg.knn_row_blocks(10, world)]
    pb, pe = mg.partition(10, world, rank)
    rows = torch.arange(pb, pe)[:, None] * torch.ones((1, 3), dtype=torch.int64)
    allrows = mg.all_gather_rows(rows, counts)
    assert allrows[:, 0].tolist() == list(range(10))
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MG_WORKER_OK")


if __name__ == "__main__":
    main()
