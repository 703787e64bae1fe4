"""Drop-in for the reference's SWIG module ``wrap.c_support`` (wrap/c_support.i:4,8).

    from gbnns_dim_red_b200.wrap import c_support          # reference: import wrap.c_support
    c_support.get_graphs_and_search_tests("t", "s", 128, 32, 10000, "v", 100000, False)

The reference's wrap/c_support.cpp does not compile against its own headers and swig is not needed
here: this module restates what that function does (wrap/c_support.cpp:254-398) on
top of the C ABI — every distance, prune and search runs on the GPU; there is no CPU fallback.

  1. names: transform 't' -> "triplet_wrap", 'p' -> "pca" (:257-262); dataset 's','g','w','d' ->
     sift/gist/glove/deep with their ef lists (:266-283); val 'v' -> "_valid" suffix and n = n_val,
     otherwise n = 1 000 000 (:285-303); n_tr = 100.
  2. loads <data>/<ds>/<ds>_{base,query,groundtruth}<valid>, the transformed base/query
     `_base_<file><valid>.fvecs` / `_query_<file><valid>.fvecs` (only when d > d_low) and the kNN lists
     <models>/<ds>/knn_1k_<file><valid>.ivecs (:322-380).
  3. hnswlikeGD(M = 20, reverse = reverse_gd) (:384), then one search sweep over the ef list with
     re-ranking, one random entry vertex per query, number_exper = 1, graph label "gd_knn_20"
     (:226-251, :395); the result line is printed and appended to
     <results>/<ds>/train_results_<file>.txt (:331).
  4. returns 0 like the reference (:397).  The measured accuracies of the call are additionally kept in
     ``last_results()`` — the reference's Python side treats the return value as an accuracy
     (dim_red/triplet.py:153-160) although the C++ always returns 0.

The angular trainer passes a 9th positional argument (dim_red/angular.py:190-191); extra arguments
are accepted and ignored.  Directories default to the reference's hard-coded ones and are overridden
with GBDR_DATA_ROOT / GBDR_MODELS_ROOT / GBDR_TRAIN_RESULTS_ROOT; GBDR_SEED fixes the entry points.
"""
from __future__ import annotations

import os
import time

import numpy as np

from .. import capi, xvecs

_TRANSFORMS = {"t": "triplet_wrap", "p": "pca"}
_DATASETS = {"s": ("sift", [150]), "g": ("gist", [250, 500]), "w": ("glove", [500]), "d": ("deep", [100, 150])}
_last = []


def last_results():
    """[(ef, acc, hops, dist_calc, work_time)] of the most recent call."""
    return list(_last)


def _ch(x):
    return x.decode() if isinstance(x, bytes) else str(x)


def search_tests(ds, queries, truth, ds_low, queries_low, knn, efs, M=20, reverse_gd=False, graph_label="gd_knn_20",
                 output_txt=None, seed=None, device=None):
    """In-memory form of the same check: the training loop hands over its arrays instead of
    writing them to the reference's file layout first (dim_red/triplet.py:142-153).

      ds [n,d], queries [n_q,d], truth [n_q,>=1]           original-dimension data and ground truth
      ds_low [n,d_low], queries_low [n_q,d_low]            transformed data (pass ds / queries when d == d_low)
      knn                                                  kNN lists of ds_low: id matrix [n,k] or (offsets, edges)

    hnswlikeGD(M, reverse_gd) on the GPU (wrap/c_support.cpp:384), then one sweep over `efs` with re-ranking and one
    random entry vertex per query (:226-251).  Prints the reference's result lines (appends them to `output_txt`
    when given), records them in last_results() and returns [(ef, acc, hops, dist_calc, work_time)]."""
    ds, queries = np.ascontiguousarray(ds, np.float32), np.ascontiguousarray(queries, np.float32)
    ds_low, queries_low = np.ascontiguousarray(ds_low, np.float32), np.ascontiguousarray(queries_low, np.float32)
    truth = np.asarray(truth)
    n, n_q = ds_low.shape[0], queries_low.shape[0]
    koff, kedges = knn if isinstance(knn, tuple) else xvecs.adjacency_from_matrix(np.ascontiguousarray(knn, np.uint32))
    if device is None:
        device = int(os.environ.get("GBDR_DEVICE", "0"))
    goff, gedges, _ = capi.gd_prune(koff, kedges, ds_low, M=M, reverse=bool(reverse_gd), device=device)
    print("GD_knn_low", int(gedges.size / max(n, 1)))
    print(f"GD knn {M} ")

    if seed is None:
        seed = os.environ.get("GBDR_SEED")
    rng = np.random.default_rng(int(seed) if seed not in (None, "") else None)
    entry = rng.integers(0, n, size=n_q, dtype=np.uint32)  # graph label is not "hnsw*": random entry vertices

    ix = capi.Index(device)
    try:
        ix.set_low(ds_low)
        ix.set_graph(goff, gedges)
        low_dim = ds.shape[1] != ds_low.shape[1]
        if low_dim:
            ix.set_base(ds)
        _last.clear()
        for ef in efs:
            t0 = time.perf_counter()
            if low_dim:
                r = ix.search(queries, queries_low, ef, 1, entry, flags=capi.SEARCH_RERANK)
            else:
                r = ix.search(None, queries_low, ef, 1, entry, flags=0)
            work = time.perf_counter() - t0
            acc = float((r["ids"][:, 0] == truth[:, 0]).mean())  # no duplicate-GT fix here (:213-215)
            hops = int(r["hops"].astype(np.int64).sum()) // n_q
            dist_calc = int(r["dist_calc"].astype(np.int64).sum()) // n_q
            line = xvecs.format_result_line(graph_label, acc, hops, dist_calc, work / n_q)
            print(line)
            if output_txt:
                with open(output_txt, "a") as f:
                    f.write(line + "\n")
            _last.append((ef, acc, hops, dist_calc, work / n_q))
    finally:
        ix.close()
    return list(_last)


def get_graphs_and_search_tests(transform_type, dataset, d_p, d_low_p, n_q_p, val, n_val, reverse_gd, *_ignored):
    file_name = _TRANSFORMS.get(_ch(transform_type), "")
    dataset_name, efs = _DATASETS.get(_ch(dataset), ("", []))
    valid = "_valid" if _ch(val) == "v" else ""
    n = int(n_val) if valid else 1_000_000
    n_q, n_tr, d, d_low = int(n_q_p), 100, int(d_p), int(d_low_p)
    print(n, n_q, n_tr, d, d_low)

    data_root = os.environ.get("GBDR_DATA_ROOT", "/mnt/data/shekhale/data")
    models_root = os.environ.get("GBDR_MODELS_ROOT", "/mnt/data/shekhale/models/nns_graphs")
    results_root = os.environ.get("GBDR_TRAIN_RESULTS_ROOT", "/home/shekhale/results/dim_red")
    path_data = os.path.join(data_root, dataset_name, dataset_name)
    path_models = os.path.join(models_root, dataset_name) + "/"
    output_txt = os.path.join(results_root, dataset_name, f"train_results_{file_name}.txt")

    print("Loading data from", path_data + "_base" + valid + ".fvecs")
    ds = xvecs.read_fvecs(path_data + "_base" + valid + ".fvecs", d=d, n=n)
    queries = xvecs.read_fvecs(path_data + "_query" + valid + ".fvecs", d=d, n=n_q)
    truth = xvecs.read_ivecs(path_data + "_groundtruth" + valid + ".ivecs", d=n_tr, n=n_q)
    if d > d_low:
        ds_low = xvecs.read_fvecs(path_data + "_base_" + file_name + valid + ".fvecs", d=d_low, n=n)
        queries_low = xvecs.read_fvecs(path_data + "_query_" + file_name + valid + ".fvecs", d=d_low, n=n_q)
    else:
        ds_low, queries_low = ds, queries
    koff, kedges = xvecs.read_edges(path_models + "knn_1k_" + file_name + valid + ".ivecs", n=n)
    print("knn_low", int(kedges.size / max(n, 1)))

    os.makedirs(os.path.dirname(output_txt), exist_ok=True)
    search_tests(ds, queries, truth, ds_low, queries_low, (koff, kedges), efs, M=20, reverse_gd=bool(reverse_gd),
                 output_txt=output_txt)
    return 0
