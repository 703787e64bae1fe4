"""Builds the synthetic workloads of BASELINE.json end to end on the GPU:

    vectors -> projection net -> db_low = net(base) [K1] -> kNN-1k of db_low [K4]
            -> hnswlikeGD(M=30, reverse) [K6] -> ground truth = exact top-n_tr in the original dim [K4]

Everything here goes through the C ABI (capi); nothing touches the CPU oracle.  Results are cached
under ``cache_dir`` so the two bench arms (and repeated runs on one box) share one build.
"""
from __future__ import annotations

import json
import os
import time

import numpy as np

from . import capi, synth, xvecs


def _cache_paths(cache_dir, key):
    d = os.path.join(cache_dir, key)
    return d, os.path.join(d, "meta.json")


# graph per workload: GD-pruned kNN-1k (prepare_graph.cpp:66-70) everywhere except the Deep-1M shape,
# which searches a constant-degree low-dim kNN graph (the 32 nearest non-self neighbours; README "fixed graph" case)
GRAPH_KIND = {"deep1m": "knn32"}


def build_workload(name="sift1m", device=0, cache_dir=None, knn_k=1000, M=30, n_tr=100, latent=8, seed=1234,
                   n=None, n_q=None, log=None, proj_mode=None, graph=None):
    """Returns dict(base, queries, net, db_low, graph=(offsets, edges), truth, entry, shape, timings)."""
    shape = dict(synth.SHAPES[name])
    graph = graph or GRAPH_KIND.get(name, "gd")
    if graph == "knn32":
        knn_k = 33
    if n:
        shape["n"] = n
    if n_q:
        shape["n_q"] = n_q
    knn_k = min(knn_k, shape["n"])
    key = f"{name}_n{shape['n']}_q{shape['n_q']}_k{knn_k}_M{M}_L{latent}_s{seed}" + ("" if graph == "gd" else "_" + graph + "cut")
    shape["graph"] = graph
    log = log or (lambda *a: None)
    if cache_dir:
        cdir, meta = _cache_paths(cache_dir, key)
        if os.path.exists(meta):
            try:
                w = dict(shape=shape, timings=json.load(open(meta))["timings"], cached=True)
                for f in ("base", "queries", "l1", "l2", "l3", "db_low", "goff", "gedges", "truth", "entry"):
                    w[f] = np.load(os.path.join(cdir, f + ".npy"), mmap_mode=None)
                w["net"] = (w.pop("l1"), w.pop("l2"), w.pop("l3"))
                w["graph"] = (w.pop("goff"), w.pop("gedges"))
                log(f"workload {key}: loaded from cache")
                return w
            except Exception as e:  # stale / partial cache: rebuild
                log(f"workload cache unusable ({e}); rebuilding")
    t = {}
    t0 = time.time()
    base, queries = synth.make_vectors(shape["n"], shape["d"], shape["n_q"], latent=latent, seed=seed)
    net = synth.make_net(shape["d"], shape["d_hidden"], shape["d_low"], seed=seed)
    t["generate_s"] = time.time() - t0
    log(f"generated vectors in {t['generate_s']:.1f}s")

    ix = capi.Index(device)
    ix.set_net(*net)
    if proj_mode is not None:
        ix.set_projection_mode(proj_mode)
    t0 = time.time()
    db_low = np.empty((shape["n"], shape["d_low"]), np.float32)
    step = 1 << 18
    for i in range(0, shape["n"], step):
        db_low[i:i + step] = ix.project(base[i:i + step])
    t["project_base_s"] = time.time() - t0
    ix.close()
    log(f"projected base in {t['project_base_s']:.1f}s")

    # the n x k id matrix lands in page-locked host memory (allocated outside the timed call, as a file writer's
    # staging buffer would be), so its chunks stream out over PCIe behind the computation
    pinned = capi.PinnedArray((shape["n"], knn_k), np.uint32)
    t0 = time.time()
    if graph == "knn32":
        # the README's "fixed" constant-degree graph: cutKNNbyK(k = 32) of the kNN file (support_func.h:309-340), which
        # keeps the vertex itself (rank 0, distance 0) like the reference does
        knn_ids, knn_gpu_s = capi.knn(db_low, db_low, knn_k, device=device, out_ids=pinned.array)
        t["knn_build_s"] = knn_gpu_s
        t["knn_build_wall_s"] = time.time() - t0
        log(f"kNN-{knn_k} lists: {knn_gpu_s:.2f}s on GPU ({t['knn_build_wall_s']:.1f}s wall)")
        t0 = time.time()
        koff, kedges = xvecs.adjacency_from_matrix(knn_ids)
        goff, gedges, cut_s = capi.knn_cut(koff, kedges, db_low, 32, device=device)
        del knn_ids, koff, kedges
        pinned.close()
        t["gd_prune_gpu_s"] = cut_s
        t["gd_prune_wall_s"] = time.time() - t0
        log(f"cutKNNbyK(32): {cut_s:.2f}s on GPU, degree {gedges.size / shape['n']:.1f}")
    else:
        # kNN-1k + hnswlikeGD without leaving HBM (gbdr_build_graph): one upload of the vectors, one download of the graph,
        # the kNN lists (the reference's `_knn_1k_` file) streamed to the host behind the computation
        goff, gedges, bt = capi.build_graph(db_low, knn_k=knn_k, M=M, reverse=True, knn_out=pinned.array, device=device)
        pinned.close()
        t["knn_build_s"] = bt["knn_s"]
        t["knn_build_wall_s"] = bt["upload_s"] + bt["knn_s"]
        t["gd_prune_gpu_s"] = bt["prune_s"] + bt["finish_s"]
        t["gd_prune_wall_s"] = time.time() - t0 - t["knn_build_wall_s"]
        t["build_graph"] = bt
        log(f"kNN-{knn_k} {bt['knn_s']:.2f}s + hnswlikeGD {bt['prune_s']:.2f}s forward, {bt['finish_s']:.2f}s reverse pass and "
            f"output ({time.time() - t0:.1f}s wall), avg degree {gedges.size / shape['n']:.1f}")

    t0 = time.time()
    truth, gt_s = capi.knn(queries, base, min(n_tr, shape["n"]), device=device)
    t["ground_truth_s"] = gt_s
    log(f"ground truth: {gt_s:.2f}s on GPU")
    entry = synth.make_entry_points(shape["n"], shape["n_q"], seed=seed)

    w = dict(base=base, queries=queries, net=net, db_low=db_low, graph=(goff, gedges), truth=truth, entry=entry,
             shape=shape, timings=t, cached=False)
    if cache_dir:
        try:
            cdir, meta = _cache_paths(cache_dir, key)
            os.makedirs(cdir, exist_ok=True)
            for f, a in (("base", base), ("queries", queries), ("l1", net[0]), ("l2", net[1]), ("l3", net[2]),
                         ("db_low", db_low), ("goff", goff), ("gedges", gedges), ("truth", truth), ("entry", entry)):
                np.save(os.path.join(cdir, f + ".npy"), a)
            with open(meta, "w") as f:
                json.dump(dict(timings=t, shape=shape), f)
        except OSError as e:
            log(f"could not write workload cache: {e}")
    return w


def recall_at_1(ids, truth, base=None):
    """Scoring loop of performTest (search_function.h:190-203): answer == truth[0], plus the
    SIFT duplicate fix — also counts truth[1] when dist(truth0, truth1) == 0."""
    ids = np.asarray(ids).reshape(-1)
    hit = ids == truth[:, 0]
    if base is not None and truth.shape[1] > 1:
        a, b = base[truth[:, 0]], base[truth[:, 1]]
        dup = (np.abs(a - b).max(axis=1) == 0) & (truth[:, 0] != truth[:, 1])
        hit = hit | (dup & (ids == truth[:, 1]))
    return float(hit.mean())


def recall_at_k(ids, truth, k=10):
    """Recall@k: |answer top-k  ∩  true top-k| / k, averaged over the queries.  The reference scores recall@1 only
    (search_function.h:190-203); this is its natural extension: `ids` [n_q, >= k] is the re-ranked
    list, best first (GBDR_SEARCH_RERANK with k results), `truth` [n_q, >= k] the exact neighbours, nearest first."""
    ids = np.asarray(ids)[:, :k].astype(np.int64)
    tr = np.asarray(truth)[:, :k].astype(np.int64)
    hits = (ids[:, :, None] == tr[:, None, :]).any(axis=2)
    return float(hits.sum(axis=1).mean() / k)
