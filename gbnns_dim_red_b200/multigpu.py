"""Multi-GPU host logic: one process per GPU, `torch.distributed` for the plumbing (NCCL over NVLink on
GPUs; the same code runs over gloo with CPU tensors in the CPU tests).

The reference's only parallelism is a data-parallel loop over queries (`#pragma omp parallel for`,
search/search_function.h:152).  Three modes follow from that:

  replicated   the whole index on every GPU, queries partitioned contiguously, NO data-path
               collective (results are concatenated only if the caller asks for them);
  sharded      rows [b, e) of base / low-dim base plus a per-shard graph (local ids) on rank r, every
               rank searches ALL queries on its shard, the per-shard top-k (dist, global id) lists are
               all-gathered and k-way merged by (dist, id) with the K5 kernel (gbdr_merge_topk_dev);
  kNN build    output rows partitioned, every rank holds all of Y, row blocks all-gathered.

Nothing here computes distances or merges on the CPU: the merge is an injected callable that defaults
to the GPU kernel (the CPU tests inject a checker to validate the gather layout and id offsets).
"""
from __future__ import annotations

import numpy as np


def partition(n_items: int, world: int, rank: int):
    """Contiguous, balanced [begin, end) of rank `rank`: the first n_items % world ranks get one extra."""
    base, rem = divmod(int(n_items), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def partitions(n_items: int, world: int):
    return [partition(n_items, world, r) for r in range(world)]


def _dist():
    import torch.distributed as dist

    return dist


def world_info():
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def all_gather_equal(t):
    """all_gather of equally-shaped tensors -> tensor [world, *t.shape] on t's device."""
    import torch

    rank, world = world_info()
    if world == 1:
        return t.unsqueeze(0)
    t = t.contiguous()
    flat = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    _dist().all_gather_into_tensor(flat, t)  # concatenation along dim 0 (both NCCL and gloo accept this form)
    return flat.view((world,) + tuple(t.shape))


def all_gather_rows(t, counts):
    """all_gather of row blocks with per-rank row counts `counts` (ragged): pads to the maximum, gathers,
    and returns the concatenation in rank order."""
    import torch

    rank, world = world_info()
    if world == 1:
        return t
    m = max(counts)
    pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    g = all_gather_equal(pad)
    return torch.cat([g[r, : counts[r]] for r in range(world)], dim=0)


# --------------------------------------------------------------------------------------- replicated
class ReplicatedSearcher:
    """Full index on every rank; rank r answers queries [b_r, e_r).  `index` is a capi.Index (or any
    object with the same `search` method)."""

    def __init__(self, index):
        self.index = index
        self.rank, self.world = world_info()

    def my_slice(self, n_q):
        return partition(n_q, self.world, self.rank)

    def search(self, queries, q_low, ef, k, entry, flags):
        """Searches this rank's slice of the batch; returns (begin, end, result dict)."""
        b, e = self.my_slice(len(entry))
        q = None if queries is None else queries[b:e]
        ql = None if q_low is None else q_low[b:e]
        return b, e, self.index.search(q, ql, ef, k, entry[b:e], flags=flags)

    def gather(self, b, e, ids, n_q, device=None):
        """Optional: every rank obtains the ids of the whole batch (one all_gather of n_q*k*4 bytes)."""
        import torch

        t = torch.from_numpy(np.ascontiguousarray(ids).astype(np.int64))
        if device is not None:
            t = t.to(device)
        counts = [pe - pb for pb, pe in partitions(n_q, self.world)]
        return all_gather_rows(t, counts).cpu().numpy().astype(np.uint32)


# --------------------------------------------------------------------------------------- sharded
def gpu_merge(device_index):
    """The product merge: K5 on `device_index`.  Takes torch CUDA tensors ids [parts, n_q, k] (int32
    bit patterns of uint32 ids), dists [parts, n_q, k] float32 -> (ids [n_q, k_out], dists)."""
    import torch

    from . import capi

    def merge(ids, dists, k_out):
        parts, n_q, k_in = ids.shape
        out_ids = torch.empty((n_q, k_out), dtype=torch.int32, device=ids.device)
        out_d = torch.empty((n_q, k_out), dtype=torch.float32, device=ids.device)
        capi.merge_topk_dev(device_index, ids.contiguous().data_ptr(), dists.contiguous().data_ptr(), parts, n_q, k_in, k_out,
                            out_ids.data_ptr(), out_d.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
        return out_ids, out_d

    return merge


class ShardedSearcher:
    """Rank r holds rows [row_begin, row_end) of the database and a graph over LOCAL ids.  Every rank
    searches all queries on its shard with `search_local`, which must return (ids, dists) as tensors
    [n_q, k] whose ids are already GLOBAL (gbdr_index_set_id_offset(row_begin) does that inside the
    kernels) and whose unused slots hold PAD_ID / +inf; the lists are all-gathered and merged."""

    def __init__(self, search_local, merge, n_total):
        self.search_local = search_local
        self.merge = merge
        self.rank, self.world = world_info()
        self.row_begin, self.row_end = partition(n_total, self.world, self.rank)

    def search(self, k_out, *args, **kw):
        ids, dists = self.search_local(*args, **kw)
        g_ids = all_gather_equal(ids)      # [world, n_q, k]
        g_d = all_gather_equal(dists)
        return self.merge(g_ids, g_d, k_out)


# --------------------------------------------------------------------------------------- kNN build
def knn_row_blocks(n, world):
    """Row-block ownership of the kNN-graph build: rank r computes rows [b_r, e_r)."""
    return partitions(n, world)


def sharded_knn(Y, k, device_index):
    """kNN graph of all rows of Y (host float32 [n, d], present on every rank): this rank computes its
    row block on its GPU, the blocks are all-gathered.  Returns ids uint32 [n, k] on every rank."""
    import torch

    from . import capi

    rank, world = world_info()
    n = Y.shape[0]
    b, e = partition(n, world, rank)
    ids, _ = capi.knn(np.ascontiguousarray(Y[b:e]), Y, k, device=device_index)
    t = torch.from_numpy(ids.astype(np.int64))
    if world > 1 and _dist().get_backend() == "nccl":
        t = t.to(torch.device("cuda", device_index))
    counts = [pe - pb for pb, pe in partitions(n, world)]
    return all_gather_rows(t, counts).cpu().numpy().astype(np.uint32)


def equal_blocks(n, world):
    """Equal row blocks of ceil(n / world) rows (the last may be short): what an in-place all-gather wants."""
    per = -(-int(n) // int(world))
    return per, [(min(n, r * per), min(n, (r + 1) * per)) for r in range(world)]


def sharded_build_graph(Y, knn_k, M, device_index, reverse=True, need_const_degree=False):
    """The graph build (kNN-`knn_k` self-join of Y, then hnswlikeGD) row-block sharded over the ranks, everything between
    the upload of Y and the download of the graph in HBM (the one-process form of the same
    flow is gbdr_group_build_graph):

        rank r uploads ITS block of Y           (n * d * 4 / world bytes over its own PCIe link)
        NCCL all-gather of the blocks            -> all of Y on every GPU
        gbdr_knn_dev / gbdr_gd_prune_dev         kNN lists and forward lists of the rank's block
        NCCL all-gather of the forward lists     (2M ids + a degree per row)
        gbdr_gd_finish_dev on rank 0             reverse-edge pass (order dependent) + flattened graph

    Y: host float32 [n, d] present on every rank (each rank reads only its block).  Returns (offsets, edges) on rank 0
    (None, None elsewhere) and a dict of seconds: upload_allgather_s, knn_s, prune_s (device time of this rank, events),
    finish_s (rank 0, host clock)."""
    import time

    import torch

    from . import capi

    rank, world = world_info()
    dev = torch.device("cuda", device_index)
    n, d = Y.shape
    per, blocks = equal_blocks(n, world)
    b, e = blocks[rank]
    st = torch.cuda.current_stream(dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    d_Y = torch.empty((per * world, d), dtype=torch.float32, device=dev)
    ev[0].record(st)
    if e > b:
        d_Y[b:e].copy_(torch.from_numpy(np.ascontiguousarray(Y[b:e])), non_blocking=False)
    if world > 1:
        _dist().all_gather_into_tensor(d_Y, d_Y[rank * per:(rank + 1) * per])
    ev[1].record(st)
    rows = e - b
    d_K = torch.empty((max(rows, 1), knn_k), dtype=torch.int32, device=dev)
    d_F = torch.full((per, 2 * M), -1, dtype=torch.int32, device=dev)
    d_D = torch.zeros((per,), dtype=torch.int32, device=dev)
    if rows:
        capi.knn_dev(device_index, d_Y.data_ptr(), b, e, d_Y.data_ptr(), n, d, knn_k, d_K.data_ptr(), 0, stream=st.cuda_stream)
    ev[2].record(st)
    if rows:
        capi.gd_prune_dev(device_index, d_K.data_ptr(), knn_k, knn_k, b, e, d_Y.data_ptr(), n, d, M, d_F.data_ptr(), d_D.data_ptr(),
                          stream=st.cuda_stream)
    if world > 1:
        g_F = torch.empty((per * world, 2 * M), dtype=torch.int32, device=dev)
        g_D = torch.empty((per * world,), dtype=torch.int32, device=dev)
        _dist().all_gather_into_tensor(g_F, d_F)
        _dist().all_gather_into_tensor(g_D, d_D)
    else:
        g_F, g_D = d_F, d_D
    g_K = None
    if need_const_degree:  # the fill walks every vertex's candidate list on rank 0
        pad = torch.zeros((per, knn_k), dtype=torch.int32, device=dev)
        pad[:rows] = d_K[:rows]
        g_K = all_gather_equal(pad).reshape(per * world, knn_k) if world > 1 else pad
    ev[3].record(st)
    st.synchronize()
    t = dict(upload_allgather_s=ev[0].elapsed_time(ev[1]) * 1e-3, knn_s=ev[1].elapsed_time(ev[2]) * 1e-3,
             prune_s=ev[2].elapsed_time(ev[3]) * 1e-3, finish_s=0.0)
    off = edges = None
    if rank == 0:
        t0 = time.perf_counter()
        off, edges = capi.gd_finish_dev(device_index, g_F.data_ptr(), g_D.data_ptr(), n, M, reverse=reverse,
                                        need_const_degree=need_const_degree, d_knn=g_K.data_ptr() if g_K is not None else 0,
                                        k=knn_k if g_K is not None else 0, kstride=knn_k if g_K is not None else 0,
                                        stream=st.cuda_stream)
        t["finish_s"] = time.perf_counter() - t0
    return off, edges, t
