#!/usr/bin/env python
"""Generates tests/golden/*.npz and the IO fixture files from the REFERENCE itself.

Run in the build container only (needs /root/reference and oracle/_ref built from it):

    python tests/golden/make_golden.py

What is recorded (all on small seeded inputs):
  * search.npz  — outputs of the reference's own C++ (oracle/_ref/libgbdr_ref_strict.so, i.e.
                  /root/reference/search/*.h compiled with -O2 -fno-fast-math -ffp-contract=off):
                  L2Metric::Dist / Angular::Dist values, GetLowQueryFromNet outputs, hnswlikeGD graph,
                  getOneSearchResults + getRealNearest results (ids, dists, hops, dist_calc) for
                  several ef in the three performTest branches.
  * second.npz  — the same C++ on the same inputs with use_second_graph == true, and hnswlikeGD on hub-heavy
                  points whose rows fill up during the reverse pass (make_second below).
  * fast.npz    — the same pipeline run end to end by the reference AS SHIPPED (README flags, -Ofast) on the seeded cases
                  of tests/_data.py (make_fast below): the targets of the north star's tolerance bars.
  * knn.npz     — output of the reference's Python get_nearestneighbors_torch
                  (dim_red/support_func.py:54-68; imported with a stub matplotlib).
  * io/         — files written by the reference's writers: dim_red/data.py write_fvecs/write_ivecs
                  and search/support_func.h writeEdges/writeXvec, plus a params file in the
                  parameters_of_databases.txt format with the values the reference's parser returns.
The committed fixtures are what travels to the GPU box; nothing reads /root/reference at test time.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from gbnns_dim_red_b200 import synth, xvecs  # noqa: E402
from tests import _oracle as O  # noqa: E402

REF = os.environ.get("GBDR_REFERENCE", "/root/reference")


def make_second(L, g):
    """second.npz — more outputs of the reference's own C++ on the inputs of search.npz:
      * getOneSearchResults with use_second_graph == true (search_function.h:73-89): a sparse random auxiliary
        graph, llf on/off, hops_bound 3 and 50, in the three performTest branches;
      * hnswlikeGD on hub-heavy points where dozens of rows fill to 2M in addReverseEdgesForGD
        (support_func.h:423-442), with and without the constant-degree pass."""
    from tests._data import hub_points, long_link_graph

    n = g["base"].shape[0]
    aoff, aedges = long_link_graph(n, degree=3, seed=21)
    out = dict(aoff=aoff, aedges=aedges)
    for llf in (0, 1):
        for hb in (3, 50):
            for mode, ef, k in ((0, 16, 1), (1, 24, 5), (2, 8, 8)):
                r = O.ref_search(g["queries"], g["q_low"], g["base"], g["db_low"], g["goff"], g["gedges"], ef, k, mode,
                                 g["entry"], aux=(aoff, aedges), llf=bool(llf), hops_bound=hb)
                for key in ("ids", "dists", "hops", "dist_calc"):
                    out[f"aux_llf{llf}_hb{hb}_m{mode}_{key}"] = r[key]
    x = hub_points(n=1500, d=16, seed=3)
    diff = x[:, None, :].astype(np.float64) - x[None, :, :].astype(np.float64)
    knn_ids = np.argsort((diff * diff).sum(-1), axis=1, kind="stable")[:, :48].astype(np.uint32)
    koff, kedges = xvecs.adjacency_from_matrix(knn_ids)
    out.update(hub_x=x, hub_knn=knn_ids)
    for M in (3, 6):
        for cd in (0, 1):
            off, ed = O.ref_gd_prune(koff, kedges, x, M=M, reverse=True, const_degree=bool(cd))
            out[f"hub_M{M}_cd{cd}_off"], out[f"hub_M{M}_cd{cd}_edges"] = off, ed
        full = int((np.diff(out[f"hub_M{M}_cd0_off"].astype(np.int64)) == 2 * M).sum())
        assert full >= 30, f"hub case M={M}: only {full} full rows"
    np.savez_compressed(os.path.join(HERE, "second.npz"), **out)


def make_fast():
    """fast.npz — outputs of the reference AS SHIPPED (oracle/_ref/libgbdr_ref_fast.so: the README's -Ofast flags) on
    the seeded cases of tests/_data.py, end to end the way final_test.cpp runs them: GetLowQueryFromNet on every query,
    getOneSearchResults on the low-dimensional base, getRealNearest in the original dimension.  These are the targets
    of the north star's tolerance bars (ids on >= 99.9 % of queries, distances 1e-5 relative, recall within 0.1 pt)."""
    from tests._data import mid_case, small_case

    assert O.ref("fast") is not None, "build oracle/_ref first (make -C oracle)"
    out = {}
    for name, c in (("small", small_case()), ("mid", mid_case())):
        goff, ged = c["graph"]
        q_low = O.ref_project(*c["net"], c["queries"], kind="fast")
        out[f"{name}_q_low"] = q_low
        for ef in (10, 40, 100):
            r = O.ref_search(c["queries"], q_low, c["base"], c["db_low"], goff, ged, ef, 1, 0, c["entry"], kind="fast")
            for key in ("ids", "dists", "hops", "dist_calc", "low_ids"):
                out[f"{name}_ef{ef}_{key}"] = r[key]
        # fingerprints of the inputs, so that a drifting generator is noticed rather than misread as a parity failure
        out[f"{name}_base_sum"] = np.array(c["base"].astype(np.float64).sum())
        out[f"{name}_edges_sum"] = np.array(ged.astype(np.uint64).sum())
    np.savez_compressed(os.path.join(HERE, "fast.npz"), **out)


def main():
    assert os.path.isdir(REF), "reference tree not found"
    L = O.ref("strict")
    assert L is not None, "build oracle/_ref first (make -C oracle)"
    os.makedirs(os.path.join(HERE, "io"), exist_ok=True)

    # ------------------------------------------------------------------ inputs
    n, d, n_q, d_low, dh, M, knn_k = 1000, 32, 48, 16, 48, 8, 64
    base, queries = synth.make_vectors(n, d, n_q, latent=6, seed=11)
    l1, l2, l3 = synth.make_net(d, dh, d_low, seed=11)
    entry = synth.make_entry_points(n, n_q, seed=11)

    # ------------------------------------------------------------------ metrics
    rng = np.random.default_rng(5)
    mv_a = rng.standard_normal((16, 37), dtype=np.float32)
    mv_b = rng.standard_normal((16, 37), dtype=np.float32)
    dims = [4, 8, 12, 16, 20, 31, 32, 33, 36, 37]
    l2_vals = np.array([[L.ref_l2(O._p(mv_a[i]), O._p(mv_b[i]), dd) for dd in dims] for i in range(16)], np.float32)
    ang_vals = np.array([[L.ref_angular(O._p(mv_a[i]), O._p(mv_b[i]), dd) for dd in dims] for i in range(16)],
                        np.float32)

    # ------------------------------------------------------------------ projection, graph, search
    db_low = O.ref_project(l1, l2, l3, base)
    q_low = O.ref_project(l1, l2, l3, queries)

    # exact kNN for the GD input: float64 ranking of direct differences, (dist,id) order.  Pairs whose
    # fp32 distances could tie are irrelevant here: the list is an INPUT of the golden GD run.
    diff = db_low[:, None, :].astype(np.float64) - db_low[None, :, :].astype(np.float64)
    dd = (diff * diff).sum(-1)
    knn_ids = np.argsort(dd, axis=1, kind="stable")[:, :knn_k].astype(np.uint32)
    koff, kedges = xvecs.adjacency_from_matrix(knn_ids)
    goff, gedges = O.ref_gd_prune(koff, kedges, db_low, M=M, reverse=True)
    goff_nr, gedges_nr = O.ref_gd_prune(koff, kedges, db_low, M=M, reverse=False)
    goff_cd, gedges_cd = O.ref_gd_prune(koff, kedges, db_low, M=M, reverse=True, const_degree=True)

    out = dict(base=base, queries=queries, l1=l1, l2=l2, l3=l3, entry=entry, db_low=db_low, q_low=q_low,
               knn_ids=knn_ids, goff=goff, gedges=gedges, goff_nr=goff_nr, gedges_nr=gedges_nr, goff_cd=goff_cd,
               gedges_cd=gedges_cd, mv_a=mv_a, mv_b=mv_b, dims=np.array(dims), l2_vals=l2_vals, ang_vals=ang_vals,
               M=np.array(M))
    for ef in (1, 4, 16, 40):
        r = O.ref_search(queries, q_low, base, db_low, goff, gedges, ef, 1, 0, entry)
        for key in ("ids", "dists", "hops", "dist_calc", "low_ids", "low_dists"):
            out[f"rerank_ef{ef}_{key}"] = r[key]
    for ef, k in ((8, 8), (24, 5)):
        r = O.ref_search(None, q_low, None, db_low, goff, gedges, ef, k, 1, entry)
        for key in ("ids", "dists", "hops", "dist_calc"):
            out[f"low_ef{ef}_k{k}_{key}"] = r[key]
        r = O.ref_search(queries, None, base, None, goff, gedges, ef, k, 2, entry)
        for key in ("ids", "dists", "hops", "dist_calc"):
            out[f"plain_ef{ef}_k{k}_{key}"] = r[key]
    np.savez_compressed(os.path.join(HERE, "search.npz"), **out)
    make_second(L, out)
    make_fast()

    # ------------------------------------------------------------------ python kNN of the reference
    sys.modules.setdefault("matplotlib", types.ModuleType("matplotlib"))
    sys.modules.setdefault("matplotlib.pyplot", types.ModuleType("matplotlib.pyplot"))
    sys.path.insert(0, REF)
    from dim_red import data as rdata  # noqa: E402
    from dim_red import support_func as rsf  # noqa: E402

    torch_knn = rsf.get_nearestneighbors_torch(db_low, db_low, 20, "cpu")
    np.savez_compressed(os.path.join(HERE, "knn.npz"), db_low=db_low, torch_knn=torch_knn.astype(np.int64))

    # ------------------------------------------------------------------ IO fixtures
    io = os.path.join(HERE, "io")
    small_f = base[:7, :5].copy()
    small_i = knn_ids[:7, :6].astype(np.int32)
    rdata.write_fvecs(os.path.join(io, "py_writer.fvecs"), small_f)
    rdata.write_ivecs(os.path.join(io, "py_writer.ivecs"), small_i)
    sub_off = goff[:13].copy()
    sub_edges = gedges[: int(sub_off[-1])].copy()
    L.ref_write_edges(os.path.join(io, "cpp_writer_edges.ivecs").encode(), O._p(O._u64(sub_off)), O._p(sub_edges), 12)
    ff = np.ascontiguousarray(small_f)
    L.ref_write_fvecs(os.path.join(io, "cpp_writer.fvecs").encode(), O._p(ff), 5, 7)
    params = os.path.join(io, "params.txt")
    with open(params, "w") as f:
        f.write("toy n 1000\ntoy n_q 48\ntoy d 32\ntoy d_low 16\ntoy d_hidden 48\n"
                "toy efs 1,3,8,15,x7,20\ntoy efs_sp 1, 2\nother n 5\ntoy n_tr 10\nmalformed line with many tokens\ntoy onlytwo\n")
    import ctypes as C

    got = {}
    for key in ("n", "n_q", "d", "d_low", "d_hidden", "efs", "efs_sp", "n_tr", "missing"):
        buf = C.create_string_buffer(256)
        L.ref_read_param(params.encode(), b"toy", key.encode(), buf, 256)
        got[key] = buf.value.decode()
    arr = (C.c_int * 16)()
    m = L.ref_parse_int_list(got["efs"].encode(), arr, 16)
    np.savez(os.path.join(io, "expected.npz"), small_f=small_f, small_i=small_i, sub_off=sub_off, sub_edges=sub_edges,
             param_keys=np.array(list(got.keys())), param_vals=np.array(list(got.values())),
             efs=np.array(list(arr)[:m]),
             avg_degree=np.array(L.ref_find_graph_average_degree(O._p(O._u64(goff)), O._p(gedges), n)))
    print("golden fixtures written under", HERE)


if __name__ == "__main__":
    main()
