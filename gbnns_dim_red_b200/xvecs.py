"""On-disk formats of the reference, byte for byte.

* fvecs / ivecs: per row ``int32 dim`` + ``dim`` x {float32 | int32}, little endian
  (reference: search/support_func.h:176-202 ``readXvec``/``writeXvec``;
  dim_red/data.py:8-21 ``write_fvecs``/``write_ivecs``, :40-72 readers).
* edge lists: per vertex ``uint32 deg`` + ``deg`` x uint32 — the ivecs layout with a
  per-row length (search/support_func.h:205-217 ``writeEdges``, :231-249 ``loadEdges``).
* projection-net matrices: fvecs with rows ``[W[o, 0:in], b[o]]``
  (dim_red/support_func.py:517-555).
* parameters_of_databases.txt: ``<dataset> <key> <value>`` lines
  (search/support_func.h:578-621).
* result lines: ``graph_type <name> acc <f> hops <i> dist_calc <i> work_time <f>``
  (search/search_function.h:206-209).
"""
from __future__ import annotations

import numpy as np


def write_fvecs(path, vecs) -> None:
    """dim_red/data.py:8-13 (vectorised: same bytes)."""
    a = np.ascontiguousarray(vecs, dtype="<f4")
    n, d = a.shape
    out = np.empty((n, d + 1), dtype="<i4")
    out[:, 0] = d
    out[:, 1:] = a.view("<i4")
    out.tofile(path)


def write_ivecs(path, vecs) -> None:
    """dim_red/data.py:16-21."""
    a = np.ascontiguousarray(vecs)
    if a.dtype.kind == "u":
        a = a.astype("<u4").view("<i4")
    a = a.astype("<i4", copy=False)
    n, d = a.shape
    out = np.empty((n, d + 1), dtype="<i4")
    out[:, 0] = d
    out[:, 1:] = a
    out.tofile(path)


def read_fvecs(path, d=None, n=None) -> np.ndarray:
    """loadXvecs<float> (search/support_func.h:220-228).  With ``d`` given, every row header
    must equal ``d`` (readXvec :181-188 exits otherwise -> here ValueError)."""
    raw = np.fromfile(path, dtype="<i4")
    if raw.size == 0:
        return np.zeros((0, d or 0), dtype=np.float32)
    dim = int(raw[0])
    if d is not None and dim != d:
        raise ValueError(f"file error: dim {dim}, d {d}")
    rows = raw.reshape(-1, dim + 1)
    if not (rows[:, 0] == dim).all():
        raise ValueError("file error: inconsistent row headers")
    if n is not None:
        if rows.shape[0] < n:
            raise ValueError(f"file has {rows.shape[0]} rows, expected {n}")
        rows = rows[:n]
    return np.ascontiguousarray(rows[:, 1:]).view(np.float32)


def read_ivecs(path, d=None, n=None) -> np.ndarray:
    """loadXvecs<uint32_t>; dim_red/data.py:40-43 ``ivecs_read``."""
    raw = np.fromfile(path, dtype="<i4")
    if raw.size == 0:
        return np.zeros((0, d or 0), dtype=np.uint32)
    dim = int(raw[0])
    if d is not None and dim != d:
        raise ValueError(f"file error: dim {dim}, d {d}")
    rows = raw.reshape(-1, dim + 1)
    if not (rows[:, 0] == dim).all():
        raise ValueError("file error: inconsistent row headers")
    if n is not None:
        rows = rows[:n]
    return np.ascontiguousarray(rows[:, 1:]).view(np.uint32)


def write_edges(path, offsets, edges) -> None:
    """writeEdges (search/support_func.h:205-217) from flattened adjacency."""
    offsets = np.asarray(offsets, dtype=np.uint64)
    edges = np.asarray(edges, dtype="<u4")
    n = offsets.size - 1
    deg = np.diff(offsets).astype("<u4")
    out = np.empty(n + edges.size, dtype="<u4")
    pos = offsets[:-1] + np.arange(n, dtype=np.uint64)  # slot of each header
    out[pos.astype(np.int64)] = deg
    mask = np.ones(out.size, dtype=bool)
    mask[pos.astype(np.int64)] = False
    out[mask] = edges
    out.tofile(path)


def read_edges(path, n=None):
    """loadEdges (search/support_func.h:231-249) -> (offsets uint64 [n+1], edges uint32)."""
    raw = np.fromfile(path, dtype="<u4")
    degs = []
    p = 0
    total = raw.size
    # fast path: constant degree (kNN-1k files).  A header walk that starts at word 0 and finds the value k
    # at every (k+1)-th word IS the walk loadEdges does (each header sends it exactly k+1 words on), so "every header
    # slot holds k" proves the parse; the n given by the caller must not ask for more rows than the file holds.
    if total and total % (int(raw[0]) + 1) == 0:
        k = int(raw[0])
        rows = raw.reshape(-1, k + 1)
        if (rows[:, 0] == k).all() and (n is None or rows.shape[0] >= n):
            if n is not None:
                rows = rows[:n]
            offsets = np.arange(rows.shape[0] + 1, dtype=np.uint64) * np.uint64(k)
            return offsets, np.ascontiguousarray(rows[:, 1:]).reshape(-1)
    while p < total and (n is None or len(degs) < n):
        dg = int(raw[p])
        if p + dg + 1 > total:
            raise ValueError(f"{path}: edge list of vertex {len(degs)} runs past the end of the file")
        degs.append(dg)
        p += dg + 1
    if n is not None and len(degs) < n:
        raise ValueError(f"{path}: {len(degs)} vertices in the file, {n} expected")
    degs = np.asarray(degs, dtype=np.uint64)
    offsets = np.zeros(degs.size + 1, dtype=np.uint64)
    np.cumsum(degs, out=offsets[1:])
    hdr = (offsets[:-1] + np.arange(degs.size, dtype=np.uint64)).astype(np.int64)
    mask = np.ones(int(offsets[-1]) + degs.size, dtype=bool)
    mask[hdr] = False
    edges = raw[: mask.size][mask]
    return offsets, np.ascontiguousarray(edges)


def read_search_params(path, dataset) -> dict:
    """readSearchParams (search/support_func.h:601-610): lines with exactly three
    space-separated tokens whose first token is the dataset name (:590-598)."""
    out = {}
    with open(path) as f:
        for line in f:
            toks = line.rstrip("\n").split(" ")
            if len(toks) == 3 and toks[0] == dataset:
                out[toks[1]] = toks[2]
    return out


def int_list(s: str):
    """getVectorFromString (search/support_func.h:613-621): atoi semantics per item."""
    out = []
    for tok in s.split(","):
        t = tok.strip()
        num = ""
        for i, ch in enumerate(t):
            if ch.isdigit() or (i == 0 and ch in "+-"):
                num += ch
            else:
                break
        out.append(int(num) if num not in ("", "+", "-") else 0)
    return out


def format_result_line(graph_name, acc, hops, dist_calc, work_time) -> str:
    """search/search_function.h:206-209.  hops and dist_calc are integer divisions there."""
    return f"graph_type {graph_name} acc {acc:g} hops {int(hops)} dist_calc {int(dist_calc)} work_time {work_time:g}"


def adjacency_from_lists(lists):
    """vector<vector<uint32_t>> -> (offsets, edges)."""
    deg = np.fromiter((len(r) for r in lists), dtype=np.uint64, count=len(lists))
    offsets = np.zeros(len(lists) + 1, dtype=np.uint64)
    np.cumsum(deg, out=offsets[1:])
    edges = np.concatenate([np.asarray(r, dtype=np.uint32) for r in lists]) if len(lists) else np.zeros(0, np.uint32)
    return offsets, edges.astype(np.uint32)


def adjacency_from_matrix(mat):
    """Fixed-degree id matrix [n x k] -> (offsets, edges)."""
    mat = np.ascontiguousarray(mat, dtype=np.uint32)
    n, k = mat.shape
    return np.arange(n + 1, dtype=np.uint64) * np.uint64(k), mat.reshape(-1)
