"""Worker of tests/test_gpu_multigpu.py::test_two_gpu_*: torch.distributed.run, NCCL, one rank per GPU.
Sharded search (rows split across ranks, per-shard graphs, NCCL all-gather, GPU merge), replicated
search (query split, no collective on the data path) and row-block sharded kNN build; each checked
against the CPU oracle on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gbnns_dim_red_b200 import capi, multigpu as mg, xvecs  # noqa: E402
from tests import _oracle as O  # noqa: E402
from tests._data import small_case  # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", local)
    c = small_case()
    n, n_q, ef, k = c["n"], c["n_q"], 40, 10

    # ---------------- sharded search
    rb, re = mg.partition(n, world, rank)
    low, base = c["db_low"][rb:re], c["base"][rb:re]
    knn_ids, _ = capi.knn(low, low, 48, device=local)
    goff, ged, _ = capi.gd_prune(*xvecs.adjacency_from_matrix(knn_ids), low, M=10, reverse=True, device=local)
    ix = capi.Index(local)
    ix.set_base(base)
    ix.set_low(low)
    ix.set_graph(goff, ged)
    ix.set_id_offset(rb)
    entry = (c["entry"] % (re - rb)).astype(np.uint32)

    def search_local():
        r = ix.search(c["queries"], c["q_low"], ef, k, entry, flags=capi.SEARCH_RERANK)
        return (torch.from_numpy(r["ids"].view(np.int32)).to(dev), torch.from_numpy(r["dists"]).to(dev))

    ss = mg.ShardedSearcher(search_local, mg.gpu_merge(local), n)
    ids, dd = ss.search(k)
    torch.cuda.synchronize()
    ids = ids.cpu().numpy().view(np.uint32)
    dd = dd.cpu().numpy()
    # oracle: reference semantics per shard (CPU), then (dist,id) merge
    parts = []
    for r in range(world):
        b, e = mg.partition(n, world, r)
        lo_r, ba_r = c["db_low"][b:e], c["base"][b:e]
        ki, _ = O.orc_knn(lo_r, lo_r, 48)
        go, ge = O.orc_gd_prune(*xvecs.adjacency_from_matrix(ki), lo_r, M=10, reverse=True)
        o = O.orc_search(c["queries"], c["q_low"], ba_r, lo_r, go, ge, ef, k, 0, (c["entry"] % (e - b)).astype(np.uint32))
        parts.append((o["ids"].astype(np.int64) + b, o["dists"]))
    for q in range(n_q):
        cand = sorted((float(parts[r][1][q, j]), int(parts[r][0][q, j])) for r in range(world) for j in range(k)
                      if np.isfinite(parts[r][1][q, j]))
        assert [int(x) for x in ids[q]] == [p[1] for p in cand[:k]], (q, ids[q], cand[:k])
        assert np.array_equal(dd[q], np.array([p[0] for p in cand[:k]], np.float32))
    ix.close()

    # ---------------- replicated search: each rank answers its slice, results gathered for the check only
    ix = capi.Index(local)
    ix.set_base(c["base"])
    ix.set_low(c["db_low"])
    ix.set_graph(*c["graph"])
    rs = mg.ReplicatedSearcher(ix)
    b, e, r = rs.search(c["queries"], c["q_low"], ef, 1, c["entry"], flags=capi.SEARCH_RERANK)
    full = rs.gather(b, e, r["ids"], n_q, device=dev)
    o = O.orc_search(c["queries"], c["q_low"], c["base"], c["db_low"], c["graph"][0], c["graph"][1], ef, 1, 0, c["entry"])
    assert np.array_equal(full, o["ids"])
    ix.close()

    # ---------------- row-block sharded kNN build
    allk = mg.sharded_knn(c["db_low"], 32, local)
    oi, _ = O.orc_knn(c["db_low"], c["db_low"], 32)
    assert np.array_equal(allk, oi)

    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MG_GPU_WORKER_OK")


if __name__ == "__main__":
    main()
