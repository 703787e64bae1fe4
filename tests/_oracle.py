"""ctypes access to the CPU checkers (oracle/liboracle.so and oracle/_ref/*.so).  Tests only."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_STRICT_SO = os.path.join(ROOT, "oracle", "_ref", "libgbdr_ref_strict.so")
REF_FAST_SO = os.path.join(ROOT, "oracle", "_ref", "libgbdr_ref_fast.so")

vp, u32, u64, i32, sz = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_size_t


def _p(a):
    return None if a is None else a.ctypes.data_as(vp)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            import subprocess

            subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), ORACLE_SO], check=True)
        L = C.CDLL(ORACLE_SO)
        L.orc_l2.restype = C.c_float
        L.orc_l2.argtypes = [vp, vp, sz]
        L.orc_angular.restype = C.c_float
        L.orc_angular.argtypes = [vp, vp, sz]
        L.orc_project.argtypes = [vp, vp, vp, vp, sz, sz, sz, sz, sz, vp]
        L.orc_search_batch.argtypes = [vp, vp, vp, vp, u64, u32, u32, vp, vp, u32, i32, i32, i32, vp, vp, vp, vp, vp, vp]
        L.orc_search_batch_aux.argtypes = [vp, vp, vp, vp, u64, u32, u32, vp, vp, vp, vp, i32, u32, u32, i32, i32, i32,
                                           vp, vp, vp, vp, vp, vp]
        L.orc_knn.argtypes = [vp, u64, vp, u64, u32, u32, vp, vp]
        L.orc_gd_prune.argtypes = [vp, vp, vp, u64, u32, i32, i32, i32, vp, vp]
        L.orc_knn_cut.argtypes = [vp, vp, vp, u64, u32, u32, vp, vp]
        _oracle = L
    return _oracle


_refs = {}


def ref(kind="strict"):
    """The reference's own headers compiled by oracle/Makefile; None if not built."""
    path = REF_STRICT_SO if kind == "strict" else REF_FAST_SO
    if kind not in _refs:
        if not os.path.exists(path):
            _refs[kind] = None
        else:
            L = C.CDLL(path)
            L.ref_l2.restype = C.c_float
            L.ref_l2.argtypes = [vp, vp, sz]
            L.ref_angular.restype = C.c_float
            L.ref_angular.argtypes = [vp, vp, sz]
            L.ref_max_threads.restype = i32
            L.ref_project.argtypes = [vp, vp, vp, vp, sz, sz, sz, sz, sz, vp]
            L.ref_search_batch.argtypes = [vp, vp, vp, vp, u64, u32, u32, vp, vp, u32, i32, i32, i32, vp, vp, vp, vp,
                                           vp, vp, vp, i32]
            L.ref_search_batch_aux.argtypes = [vp, vp, vp, vp, u64, u32, u32, vp, vp, vp, vp, i32, u32, u32, i32, i32,
                                               i32, vp, vp, vp, vp, vp, vp, vp, i32]
            L.ref_gd_prune.restype = u64
            L.ref_gd_prune.argtypes = [vp, vp, vp, u64, u32, i32, i32, i32, vp, vp, i32]
            if hasattr(L, "ref_knn_cut"):
                L.ref_knn_cut.restype = u64
                L.ref_knn_cut.argtypes = [vp, vp, vp, u64, u32, i32, vp, vp]
            L.ref_ctx_create.restype = vp
            L.ref_ctx_create.argtypes = [vp, vp, vp, vp, vp, vp, vp, u64, u32, u32, u32, u32]
            L.ref_ctx_set_net.argtypes = [vp, vp, vp, vp, u32]
            L.ref_ctx_destroy.argtypes = [vp]
            dp = C.POINTER(C.c_double)
            L.ref_ctx_perform_test.argtypes = [vp, i32, i32, vp, i32, i32, dp, dp, dp, dp]
            L.ref_ctx_perform_net_test.argtypes = [vp, i32, i32, vp, i32, i32, dp, dp, dp, dp]
            L.ref_load_fvecs.argtypes = [C.c_char_p, sz, sz, vp]
            L.ref_load_ivecs.argtypes = [C.c_char_p, sz, sz, vp]
            L.ref_load_edges.restype = u64
            L.ref_load_edges.argtypes = [C.c_char_p, u32, vp, vp, u64]
            L.ref_write_edges.argtypes = [C.c_char_p, vp, vp, u64]
            L.ref_write_fvecs.argtypes = [C.c_char_p, vp, sz, sz]
            L.ref_read_param.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, sz]
            L.ref_parse_int_list.argtypes = [C.c_char_p, vp, i32]
            L.ref_find_graph_average_degree.argtypes = [vp, vp, u64]
            _refs[kind] = L
    return _refs[kind]


# ------------------------------------------------------------------ oracle wrappers
def orc_project(l1, l2, l3, queries):
    l1, l2, l3, q = _f32(l1), _f32(l2), _f32(l3), _f32(queries)
    d, dh, dh2, dl = l1.shape[1] - 1, l1.shape[0], l2.shape[0], l3.shape[0]
    out = np.empty((q.shape[0], dl), np.float32)
    oracle().orc_project(_p(l1), _p(l2), _p(l3), _p(q), q.shape[0], d, dh, dh2, dl, _p(out))
    return out


def _search(fn, queries, q_low, db, db_low, offsets, edges, ef, k, mode, entry, extra=None):
    queries = None if queries is None else _f32(queries)
    q_low = None if q_low is None else _f32(q_low)
    db = None if db is None else _f32(db)
    db_low = None if db_low is None else _f32(db_low)
    offsets, edges, entry = _u64(offsets), _u32(edges), _u32(entry)
    n = offsets.size - 1
    n_q = entry.size
    d = db.shape[1] if db is not None else 0
    d_low = db_low.shape[1] if db_low is not None else 0
    ids = np.empty((n_q, k), np.uint32)
    dists = np.empty((n_q, k), np.float32)
    hops = np.empty(n_q, np.int32)
    dc = np.empty(n_q, np.int32)
    return queries, q_low, db, db_low, offsets, edges, entry, n, n_q, d, d_low, ids, dists, hops, dc


def orc_search(queries, q_low, db, db_low, offsets, edges, ef, k, mode, entry, aux=None, llf=False, hops_bound=50):
    """aux = (offsets, edges) of the second graph -> use_second_graph == true (search_function.h:73-89)."""
    (queries, q_low, db, db_low, offsets, edges, entry, n, n_q, d, d_low, ids, dists, hops, dc) = _search(
        None, queries, q_low, db, db_low, offsets, edges, ef, k, mode, entry)
    scanned = np.empty(n_q, np.int32)
    aoff = None if aux is None else _u64(aux[0])
    aed = None if aux is None else _u32(aux[1])
    oracle().orc_search_batch_aux(_p(queries), _p(q_low), _p(db), _p(db_low), n, d, d_low, _p(offsets), _p(edges),
                                  _p(aoff), _p(aed), int(llf), hops_bound, n_q, ef, k, mode, _p(entry), _p(ids),
                                  _p(dists), _p(hops), _p(dc), _p(scanned))
    return dict(ids=ids, dists=dists, hops=hops, dist_calc=dc, scanned=scanned)


def ref_search(queries, q_low, db, db_low, offsets, edges, ef, k, mode, entry, kind="strict", threads=1, aux=None,
               llf=False, hops_bound=50):
    L = ref(kind)
    (queries, q_low, db, db_low, offsets, edges, entry, n, n_q, d, d_low, ids, dists, hops, dc) = _search(
        None, queries, q_low, db, db_low, offsets, edges, ef, k, mode, entry)
    low_ids = np.empty((n_q, ef), np.uint32) if mode == 0 else None
    low_dists = np.empty((n_q, ef), np.float32) if mode == 0 else None
    aoff = None if aux is None else _u64(aux[0])
    aed = None if aux is None else _u32(aux[1])
    L.ref_search_batch_aux(_p(queries), _p(q_low), _p(db), _p(db_low), n, d, d_low, _p(offsets), _p(edges), _p(aoff),
                           _p(aed), int(llf), hops_bound, n_q, ef, k, mode, _p(entry), _p(ids), _p(dists), _p(hops),
                           _p(dc), _p(low_ids), _p(low_dists), threads)
    return dict(ids=ids, dists=dists, hops=hops, dist_calc=dc, low_ids=low_ids, low_dists=low_dists)


def ref_project(l1, l2, l3, queries, kind="strict"):
    l1, l2, l3, q = _f32(l1), _f32(l2), _f32(l3), _f32(queries)
    d, dh, dh2, dl = l1.shape[1] - 1, l1.shape[0], l2.shape[0], l3.shape[0]
    out = np.empty((q.shape[0], dl), np.float32)
    ref(kind).ref_project(_p(l1), _p(l2), _p(l3), _p(q), q.shape[0], d, dh, dh2, dl, _p(out))
    return out


def orc_knn(Q, B, k):
    Q, B = _f32(Q), _f32(B)
    ids = np.empty((Q.shape[0], k), np.uint32)
    dists = np.empty((Q.shape[0], k), np.float32)
    oracle().orc_knn(_p(Q), Q.shape[0], _p(B), B.shape[0], B.shape[1], k, _p(ids), _p(dists))
    return ids, dists


def orc_gd_prune(offsets, edges, db_low, M=30, reverse=True, const_degree=False):
    offsets, edges, db_low = _u64(offsets), _u32(edges), _f32(db_low)
    n = offsets.size - 1
    out_off = np.empty(n + 1, np.uint64)
    out_edges = np.empty(n * 2 * M, np.uint32)
    oracle().orc_gd_prune(_p(offsets), _p(edges), _p(db_low), n, db_low.shape[1], M, int(reverse), int(const_degree),
                          _p(out_off), _p(out_edges))
    return out_off, out_edges[: int(out_off[-1])].copy()


def ref_gd_prune(offsets, edges, db_low, M=30, reverse=True, const_degree=False, kind="strict", threads=1):
    offsets, edges, db_low = _u64(offsets), _u32(edges), _f32(db_low)
    n = offsets.size - 1
    out_off = np.empty(n + 1, np.uint64)
    out_edges = np.empty(n * 2 * M, np.uint32)
    ref(kind).ref_gd_prune(_p(offsets), _p(edges), _p(db_low), n, db_low.shape[1], M, int(reverse), int(const_degree),
                           _p(out_off), _p(out_edges), threads)
    return out_off, out_edges[: int(out_off[-1])].copy()


def orc_knn_cut(offsets, edges, db, knn_size):
    offsets, edges, db = _u64(offsets), _u32(edges), _f32(db)
    n = offsets.size - 1
    out_off = np.empty(n + 1, np.uint64)
    out_edges = np.empty(max(1, n * knn_size), np.uint32)
    oracle().orc_knn_cut(_p(offsets), _p(edges), _p(db), n, db.shape[1], knn_size, _p(out_off), _p(out_edges))
    return out_off, out_edges[: int(out_off[-1])].copy()


def ref_knn_cut(offsets, edges, db, knn_size, kind="strict"):
    offsets, edges, db = _u64(offsets), _u32(edges), _f32(db)
    n = offsets.size - 1
    out_off = np.empty(n + 1, np.uint64)
    out_edges = np.empty(max(1, n * knn_size), np.uint32)
    ref(kind).ref_knn_cut(_p(offsets), _p(edges), _p(db), n, db.shape[1], knn_size, _p(out_off), _p(out_edges))
    return out_off, out_edges[: int(out_off[-1])].copy()


class RefContext:
    """Persistent copy of a dataset inside the reference harness (timing legs)."""

    def __init__(self, db, queries, db_low, q_low, truth, offsets, edges, kind="fast"):
        self.L = ref(kind)
        db, queries, db_low = _f32(db), _f32(queries), _f32(db_low)
        q_low = None if q_low is None else _f32(q_low)
        truth, offsets, edges = _u32(truth), _u64(offsets), _u32(edges)
        self.n_q = queries.shape[0]
        self.h = self.L.ref_ctx_create(_p(db), _p(queries), _p(db_low), _p(q_low), _p(truth), _p(offsets), _p(edges),
                                       db.shape[0], db.shape[1], db_low.shape[1], queries.shape[0], truth.shape[1])

    def set_net(self, l1, l2, l3):
        l1, l2, l3 = _f32(l1), _f32(l2), _f32(l3)
        self.L.ref_ctx_set_net(self.h, _p(l1), _p(l2), _p(l3), l1.shape[0])

    def perform_test(self, ef, entry, n_q_use=None, number_exper=1, threads=1, net=False):
        entry = _u32(entry)
        n_q_use = n_q_use or self.n_q
        a, h, dc, w = C.c_double(), C.c_double(), C.c_double(), C.c_double()
        fn = self.L.ref_ctx_perform_net_test if net else self.L.ref_ctx_perform_test
        rc = fn(self.h, ef, n_q_use, _p(entry), number_exper, threads, C.byref(a), C.byref(h), C.byref(dc), C.byref(w))
        if rc != 0:
            raise RuntimeError(f"reference performTest failed: {rc}")
        return dict(acc=a.value, hops=h.value, dist_calc=dc.value, work_time=w.value)

    def close(self):
        if self.h:
            self.L.ref_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
