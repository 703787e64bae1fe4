"""Sharded-index search on the GPU: per-shard search with global ids + K5 merge vs the oracle
(reference semantics per shard, then a (dist,id) merge).  Single-GPU variant always runs; the NCCL
variant needs >= 2 GPUs (gpurun --gpus 2)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from gbnns_dim_red_b200 import capi, multigpu as mg, xvecs

from . import _oracle as O
from ._data import small_case

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("shards", [2, 3])
def test_sharded_search_on_one_gpu(gpu_index_factory, shards):
    import torch

    c = small_case()
    n, n_q, ef, k = c["n"], c["n_q"], 30, 8
    dev = torch.device("cuda", 0)
    ids_l, dd_l, want = [], [], []
    for r in range(shards):
        b, e = mg.partition(n, shards, r)
        low, base = c["db_low"][b:e], c["base"][b:e]
        ki, _ = O.orc_knn(low, low, 40)
        go, ge = O.orc_gd_prune(*xvecs.adjacency_from_matrix(ki), low, M=8, reverse=True)
        entry = (c["entry"] % (e - b)).astype(np.uint32)
        ix = gpu_index_factory()
        ix.set_base(base)
        ix.set_low(low)
        ix.set_graph(go, ge)
        ix.set_id_offset(b)
        g = ix.search(c["queries"], c["q_low"], ef, k, entry, flags=capi.SEARCH_RERANK)
        o = O.orc_search(c["queries"], c["q_low"], base, low, go, ge, ef, k, 0, entry)
        assert np.array_equal(g["ids"], o["ids"] + np.uint32(b))      # global ids straight out of the kernels
        ids_l.append(torch.from_numpy(g["ids"].view(np.int32)).to(dev))
        dd_l.append(torch.from_numpy(g["dists"]).to(dev))
        want.append((o["ids"].astype(np.int64) + b, o["dists"]))
    ids, dd = mg.gpu_merge(0)(torch.stack(ids_l), torch.stack(dd_l), k)
    torch.cuda.synchronize()
    ids, dd = ids.cpu().numpy().view(np.uint32), dd.cpu().numpy()
    for q in range(n_q):
        cand = sorted((float(want[r][1][q, j]), int(want[r][0][q, j])) for r in range(shards) for j in range(k))
        assert [int(x) for x in ids[q]] == [p[1] for p in cand[:k]]
        assert np.array_equal(dd[q], np.array([p[0] for p in cand[:k]], np.float32))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_gpu_nccl_sharded_replicated_and_knn():
    if capi.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(HERE, "_mg_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MG_GPU_WORKER_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
