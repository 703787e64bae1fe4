"""ctypes binding of include/gbdr.h (the C ABI of libgbdr.so).

This is host-side plumbing only: every call goes straight into the CUDA library.  There is no
CPU fallback anywhere in this package — if the library is missing, or no B200-class device is
visible, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgbdr.so")

PAD_ID = 0xFFFFFFFF
SEARCH_RERANK = 1
SEARCH_PLAIN = 2
SEARCH_SECOND_GRAPH = 4
PROJ_3XTF32, PROJ_TF32, PROJ_FP32 = 0, 1, 2

#: every symbol include/gbdr.h declares (tests/test_abi.py checks the header against this and the .so)
SYMBOLS = [
    "gbdr_version", "gbdr_last_error", "gbdr_device_count",
    "gbdr_index_create", "gbdr_index_destroy", "gbdr_index_create_view", "gbdr_index_set_base", "gbdr_index_set_low",
    "gbdr_index_set_graph", "gbdr_index_set_aux_graph", "gbdr_index_set_net", "gbdr_index_set_id_offset",
    "gbdr_index_set_projection_mode", "gbdr_project", "gbdr_project_dev", "gbdr_search",
    "gbdr_search_submit", "gbdr_search_wait", "gbdr_search_dev", "gbdr_last_kernel_ms", "gbdr_kernel_ms", "gbdr_index_status", "gbdr_launch_count", "gbdr_knn",
    "gbdr_knn_dev", "gbdr_gd_prune", "gbdr_knn_cut", "gbdr_merge_topk_dev", "gbdr_dev_malloc", "gbdr_dev_free",
    "gbdr_memcpy_h2d", "gbdr_memcpy_d2h", "gbdr_host_alloc_pinned", "gbdr_host_free_pinned",
    "gbdr_host_register", "gbdr_host_unregister",
    "gbdr_device_synchronize", "gbdr_index_stream", "gbdr_beam_plan_info", "gbdr_index_device_ptrs",
    "gbdr_gd_prune_dev", "gbdr_gd_finish_dev", "gbdr_build_graph",
    "gbdr_group_create", "gbdr_group_destroy", "gbdr_group_size", "gbdr_group_member", "gbdr_group_set_exchange",
    "gbdr_group_set_net", "gbdr_group_set_base", "gbdr_group_set_low", "gbdr_group_set_graph", "gbdr_group_shard_rows",
    "gbdr_group_set_shard_graph", "gbdr_group_search", "gbdr_group_build_graph",
]
GROUP_REPLICATED, GROUP_SHARDED = 0, 1
EXCHANGE_PEER, EXCHANGE_NCCL = 0, 1


class GbdrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"gbdr error {code}: {msg}")
        self.code = code


_lib = None


def _point_at_torch_nccl():
    """libgbdr.so loads NCCL at run time for the group's NCCL exchange.  Shared libraries are shared by soname: if the
    system's libnccl.so.2 is mapped first, a later `import torch` in the same process is served by it too and fails on
    the symbols of the newer NCCL torch is built against.  So name the copy that ships with torch (the nvidia-nccl wheel)
    when there is one; found without importing torch."""
    if os.environ.get("GBDR_NCCL_LIB"):
        return
    try:
        import importlib.util

        spec = importlib.util.find_spec("nvidia.nccl")
        for root in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
            cand = os.path.join(root, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["GBDR_NCCL_LIB"] = cand
                return
    except Exception:
        pass


def lib():
    """Load libgbdr.so (built in-tree by gbnns_dim_red_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA extension first "
            "(python -m gbnns_dim_red_b200.build). There is no CPU fallback."
        )
    _point_at_torch_nccl()
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    L.gbdr_version.restype = i32
    L.gbdr_last_error.restype = C.c_char_p
    L.gbdr_launch_count.restype = u64
    L.gbdr_device_count.argtypes = [C.POINTER(i32)]
    L.gbdr_index_create.argtypes = [i32, C.POINTER(vp)]
    L.gbdr_index_destroy.argtypes = [vp]
    L.gbdr_index_create_view.argtypes = [vp, C.POINTER(vp)]
    L.gbdr_index_set_base.argtypes = [vp, vp, u64, u32]
    L.gbdr_index_set_low.argtypes = [vp, vp, u64, u32]
    L.gbdr_index_set_graph.argtypes = [vp, vp, vp, u64]
    L.gbdr_index_set_aux_graph.argtypes = [vp, vp, vp, u64, u32, i32]
    L.gbdr_index_set_net.argtypes = [vp, vp, vp, vp, u32, u32, u32, u32]
    L.gbdr_index_set_id_offset.argtypes = [vp, u64]
    L.gbdr_index_set_projection_mode.argtypes = [vp, i32]
    L.gbdr_project.argtypes = [vp, vp, u32, vp]
    L.gbdr_project_dev.argtypes = [vp, vp, u32, vp, vp]
    L.gbdr_search.argtypes = [vp, vp, vp, u32, u32, u32, u32, vp, vp, vp, vp, vp, C.POINTER(C.c_double)]
    L.gbdr_search_submit.argtypes = [vp, vp, vp, u32, u32, u32, u32, vp, vp, vp, vp, vp]
    L.gbdr_search_wait.argtypes = [vp, C.POINTER(C.c_double)]
    L.gbdr_search_dev.argtypes = [vp, vp, vp, u32, u32, u32, u32, vp, vp, vp, vp, vp, vp, vp]
    L.gbdr_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.gbdr_kernel_ms.argtypes = [vp, u32, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.gbdr_index_status.argtypes = [vp, C.POINTER(u32)]
    L.gbdr_knn.argtypes = [i32, vp, u64, vp, u64, u32, u32, vp, vp, C.POINTER(C.c_double)]
    L.gbdr_knn_dev.argtypes = [i32, vp, u64, u64, vp, u64, u32, u32, vp, vp, vp]
    L.gbdr_gd_prune.argtypes = [i32, vp, vp, vp, u64, u32, u32, i32, i32, vp, vp, C.POINTER(C.c_double)]
    L.gbdr_knn_cut.argtypes = [i32, vp, vp, vp, u64, u32, u32, vp, vp, C.POINTER(C.c_double)]
    L.gbdr_merge_topk_dev.argtypes = [i32, vp, vp, u32, u32, u32, u32, vp, vp, vp]
    L.gbdr_dev_malloc.argtypes = [i32, C.c_size_t, C.POINTER(vp)]
    L.gbdr_dev_free.argtypes = [i32, vp]
    L.gbdr_memcpy_h2d.argtypes = [i32, vp, vp, C.c_size_t]
    L.gbdr_memcpy_d2h.argtypes = [i32, vp, vp, C.c_size_t]
    L.gbdr_host_alloc_pinned.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.gbdr_host_free_pinned.argtypes = [vp]
    L.gbdr_host_register.argtypes = [vp, C.c_size_t]
    L.gbdr_host_unregister.argtypes = [vp]
    L.gbdr_device_synchronize.argtypes = [i32]
    L.gbdr_index_stream.argtypes = [vp, C.POINTER(vp)]
    L.gbdr_beam_plan_info.argtypes = [u32, u32, u64, i32, C.POINTER(u32)]
    L.gbdr_index_device_ptrs.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(u32)]
    dp = C.POINTER(C.c_double)
    L.gbdr_gd_prune_dev.argtypes = [i32, vp, u32, u32, u64, u64, vp, u64, u32, u32, vp, vp, vp]
    L.gbdr_gd_finish_dev.argtypes = [i32, vp, vp, u64, u32, i32, i32, vp, u32, u32, vp, vp, vp]
    L.gbdr_build_graph.argtypes = [i32, vp, u64, u32, u32, u32, i32, i32, vp, vp, vp, dp]
    L.gbdr_group_create.argtypes = [C.POINTER(i32), i32, i32, C.POINTER(vp)]
    L.gbdr_group_destroy.argtypes = [vp]
    L.gbdr_group_size.argtypes = [vp]
    L.gbdr_group_member.argtypes = [vp, i32, C.POINTER(vp)]
    L.gbdr_group_set_exchange.argtypes = [vp, i32]
    L.gbdr_group_set_net.argtypes = [vp, vp, vp, vp, u32, u32, u32, u32]
    L.gbdr_group_set_base.argtypes = [vp, vp, u64, u32]
    L.gbdr_group_set_low.argtypes = [vp, vp, u64, u32]
    L.gbdr_group_set_graph.argtypes = [vp, vp, vp, u64]
    L.gbdr_group_shard_rows.argtypes = [vp, i32, C.POINTER(u64), C.POINTER(u64)]
    L.gbdr_group_set_shard_graph.argtypes = [vp, i32, vp, vp, u64]
    L.gbdr_group_search.argtypes = [vp, vp, vp, u32, u32, u32, u32, vp, vp, vp, vp, vp, dp]
    L.gbdr_group_build_graph.argtypes = [vp, vp, u64, u32, u32, u32, i32, i32, vp, vp, vp, dp]
    _lib = L
    return L


def _chk(rc):
    if rc != 0:
        raise GbdrError(rc, lib().gbdr_last_error().decode(errors="replace"))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def device_count() -> int:
    n = C.c_int(0)
    _chk(lib().gbdr_device_count(C.byref(n)))
    return n.value


def launch_count() -> int:
    return int(lib().gbdr_launch_count())


def pinned_empty(shape, dtype):
    """numpy array over cudaHostAlloc'ed memory (kept alive by the array's base object)."""
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    p = C.c_void_p()
    _chk(lib().gbdr_host_alloc_pinned(max(nbytes, 16), C.byref(p)))
    buf = (C.c_char * max(nbytes, 16)).from_address(p.value)
    _PINNED.append(p)  # pinned staging buffers live for the life of the process
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


_PINNED = []


class Index:
    """One GPU-resident index: db, db_low, graph and projection net (include/gbdr.h)."""

    def __init__(self, device: int = 0, _parent=None):
        self._h = C.c_void_p()
        self.device = device
        self._parent = _parent
        self._views = []
        self._inflight = None
        if _parent is None:
            _chk(lib().gbdr_index_create(device, C.byref(self._h)))
        else:
            _chk(lib().gbdr_index_create_view(_parent._h, C.byref(self._h)))
        self.n = 0
        self.d = 0
        self.d_low = 0

    def view(self):
        """A second handle on the same resident data with its own stream and workspaces
        (gbdr_index_create_view): lets a second batch be in flight on this GPU."""
        root = self._parent or self
        v = Index(root.device, _parent=root)
        root._views.append(v)
        return v

    @property
    def net_dims(self):
        return (self._parent or self)._net_dims

    @net_dims.setter
    def net_dims(self, v):
        self._net_dims = v

    def close(self):
        for v in list(getattr(self, "_views", [])):
            v.close()
        self._views = []
        if getattr(self, "_h", None) is not None and self._h.value:
            if getattr(self, "_owned", True):
                _chk(lib().gbdr_index_destroy(self._h))
            self._h = C.c_void_p()
            if self._parent is not None and self in self._parent._views:
                self._parent._views.remove(self)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state ----
    def set_base(self, db):
        db = _f32(db)
        _chk(lib().gbdr_index_set_base(self._h, _ptr(db), db.shape[0], db.shape[1]))
        self.n, self.d = db.shape

    def set_low(self, db_low):
        db_low = _f32(db_low)
        _chk(lib().gbdr_index_set_low(self._h, _ptr(db_low), db_low.shape[0], db_low.shape[1]))
        self.d_low = db_low.shape[1]

    def set_graph(self, offsets, edges):
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        edges = _u32(edges)
        _chk(lib().gbdr_index_set_graph(self._h, _ptr(offsets), _ptr(edges), offsets.size - 1))

    def set_aux_graph(self, offsets, edges, hops_bound=50, llf=False):
        """Second graph of use_second_graph searches (flag SEARCH_SECOND_GRAPH); offsets=None removes it."""
        if offsets is None:
            _chk(lib().gbdr_index_set_aux_graph(self._h, None, None, 0, 0, 0))
            return
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        edges = _u32(edges)
        _chk(lib().gbdr_index_set_aux_graph(self._h, _ptr(offsets), _ptr(edges), offsets.size - 1, int(hops_bound),
                                            int(bool(llf))))

    def set_net(self, l1, l2, l3):
        l1, l2, l3 = _f32(l1), _f32(l2), _f32(l3)
        d, dh, dh2, dl = l1.shape[1] - 1, l1.shape[0], l2.shape[0], l3.shape[0]
        assert l2.shape[1] == dh + 1 and l3.shape[1] == dh2 + 1
        _chk(lib().gbdr_index_set_net(self._h, _ptr(l1), _ptr(l2), _ptr(l3), d, dh, dh2, dl))
        self.net_dims = (d, dh, dh2, dl)

    def set_id_offset(self, off):
        _chk(lib().gbdr_index_set_id_offset(self._h, int(off)))

    def set_projection_mode(self, mode):
        _chk(lib().gbdr_index_set_projection_mode(self._h, int(mode)))

    # ---- hot path (host buffers) ----
    def project(self, queries):
        q = _f32(queries)
        out = np.empty((q.shape[0], self.net_dims[3]), dtype=np.float32)
        _chk(lib().gbdr_project(self._h, _ptr(q), q.shape[0], _ptr(out)))
        return out

    def search(self, queries, q_low, ef, k, entry, flags=SEARCH_RERANK, out=None):
        """Returns dict(ids [n_q,k], dists, hops, dist_calc, gpu_seconds)."""
        q = None if queries is None else _f32(queries)
        ql = None if q_low is None else _f32(q_low)
        entry = _u32(entry)
        n_q = entry.shape[0]
        if out is None:
            out = dict(
                ids=np.empty((n_q, k), np.uint32), dists=np.empty((n_q, k), np.float32),
                hops=np.empty(n_q, np.int32), dist_calc=np.empty(n_q, np.int32),
            )
        secs = C.c_double(0)
        _chk(lib().gbdr_search(self._h, _ptr(q), _ptr(ql), n_q, ef, k, flags, _ptr(entry), _ptr(out["ids"]),
                               _ptr(out["dists"]), _ptr(out["hops"]), _ptr(out["dist_calc"]), C.byref(secs)))
        out["gpu_seconds"] = secs.value
        return out

    def search_submit(self, queries, q_low, ef, k, entry, flags=SEARCH_RERANK, out=None):
        """Asynchronous half of search(): enqueue copies + kernels on this handle's stream and return.
        Arrays must already be C-contiguous float32/uint32 (no hidden copies: they must outlive the call);
        page-locked arrays (pinned_empty) make the copies overlap other handles' kernels."""
        for a, dt in ((queries, np.float32), (q_low, np.float32), (entry, np.uint32)):
            if a is not None and not (isinstance(a, np.ndarray) and a.dtype == dt and a.flags["C_CONTIGUOUS"]):
                raise TypeError("search_submit needs C-contiguous float32 / uint32 numpy arrays")
        n_q = entry.shape[0]
        if out is None:
            out = dict(
                ids=np.empty((n_q, k), np.uint32), dists=np.empty((n_q, k), np.float32),
                hops=np.empty(n_q, np.int32), dist_calc=np.empty(n_q, np.int32),
            )
        _chk(lib().gbdr_search_submit(self._h, _ptr(queries), _ptr(q_low), n_q, ef, k, flags, _ptr(entry),
                                      _ptr(out["ids"]), _ptr(out["dists"]), _ptr(out["hops"]), _ptr(out["dist_calc"])))
        self._inflight = (out, queries, q_low, entry)  # keep the buffers alive until search_wait
        return out

    def search_wait(self):
        """Blocks until the submitted call finished; returns its output dict (with gpu_seconds)."""
        secs = C.c_double(0)
        held, self._inflight = self._inflight, None
        _chk(lib().gbdr_search_wait(self._h, C.byref(secs)))
        out = held[0] if held else {}
        out["gpu_seconds"] = secs.value
        return out

    def search_dev(self, d_queries, d_q_low, n_q, ef, k, d_entry, d_out_ids, d_out_dists=0, d_hops=0, d_dist_calc=0,
                   d_scanned=0, flags=SEARCH_RERANK, stream=0):
        """All arguments are raw device addresses (ints), asynchronous on `stream`."""
        _chk(lib().gbdr_search_dev(self._h, d_queries or None, d_q_low or None, n_q, ef, k, flags, d_entry,
                                   d_out_ids, d_out_dists or None, d_hops or None, d_dist_calc or None,
                                   d_scanned or None, stream or None))

    def project_dev(self, d_queries, n_q, d_q_low, stream=0):
        _chk(lib().gbdr_project_dev(self._h, d_queries, n_q, d_q_low, stream or None))

    def last_kernel_ms(self, last_n=1):
        """Average device time (ms) of the projection / beam-search / re-rank kernels over the last calls."""
        a, b, c = C.c_float(0), C.c_float(0), C.c_float(0)
        _chk(lib().gbdr_kernel_ms(self._h, last_n, C.byref(a), C.byref(b), C.byref(c)))
        return dict(project=a.value, search=b.value, rerank=c.value)

    def status(self) -> int:
        f = C.c_uint32(0)
        _chk(lib().gbdr_index_status(self._h, C.byref(f)))
        return f.value

    def stream(self) -> int:
        s = C.c_void_p()
        _chk(lib().gbdr_index_stream(self._h, C.byref(s)))
        return s.value or 0


class PinnedArray:
    """A numpy view over cudaHostAlloc'ed memory that is unpinned and released with close()."""

    def __init__(self, shape, dtype):
        dtype = np.dtype(dtype)
        self.nbytes = max(int(np.prod(shape)) * dtype.itemsize, 16)
        self._p = C.c_void_p()
        _chk(lib().gbdr_host_alloc_pinned(self.nbytes, C.byref(self._p)))
        buf = (C.c_char * self.nbytes).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def close(self):
        if self._p:
            self.array = None
            lib().gbdr_host_free_pinned(self._p)
            self._p = None


def knn(Q, B, k, device=0, return_dists=False, out_ids=None):
    """gbdr_knn: exact kNN ids (and distances) of rows of Q among rows of B.  `out_ids`: caller's [n_q x k]
    uint32 destination (page-locked memory lets the result chunks stream out behind the computation)."""
    B = _f32(B)
    same = Q is B
    Q = B if same else _f32(Q)
    ids = np.empty((Q.shape[0], k), np.uint32) if out_ids is None else out_ids
    assert ids.dtype == np.uint32 and ids.shape == (Q.shape[0], k) and ids.flags.c_contiguous
    dists = np.empty((Q.shape[0], k), np.float32) if return_dists else None
    secs = C.c_double(0)
    _chk(lib().gbdr_knn(device, _ptr(Q), Q.shape[0], _ptr(B), B.shape[0], B.shape[1], k, _ptr(ids), _ptr(dists),
                        C.byref(secs)))
    return (ids, dists, secs.value) if return_dists else (ids, secs.value)


def knn_dev(device, d_Q, q_begin, q_end, d_B, n, d, k, d_out_ids, d_out_dists=0, stream=0):
    _chk(lib().gbdr_knn_dev(device, d_Q, q_begin, q_end, d_B, n, d, k, d_out_ids, d_out_dists or None,
                            stream or None))


def gd_prune(knn_offsets, knn_edges, db_low, M=30, reverse=True, need_const_degree=False, device=0):
    """gbdr_gd_prune -> (offsets, edges, gpu_seconds)."""
    knn_offsets = np.ascontiguousarray(knn_offsets, dtype=np.uint64)
    knn_edges = _u32(knn_edges)
    db_low = _f32(db_low)
    n = knn_offsets.size - 1
    out_off = np.empty(n + 1, np.uint64)
    out_edges = np.empty(n * 2 * M, np.uint32)
    secs = C.c_double(0)
    _chk(lib().gbdr_gd_prune(device, _ptr(knn_offsets), _ptr(knn_edges), _ptr(db_low), n, db_low.shape[1], M,
                             int(reverse), int(need_const_degree), _ptr(out_off), _ptr(out_edges), C.byref(secs)))
    return out_off, out_edges[: int(out_off[-1])].copy(), secs.value


def knn_cut(knn_offsets, knn_edges, db, knn_size, device=0):
    """gbdr_knn_cut (cutKNNbyK) -> (offsets, edges, gpu_seconds)."""
    knn_offsets = np.ascontiguousarray(knn_offsets, dtype=np.uint64)
    knn_edges = _u32(knn_edges)
    db = _f32(db)
    n = knn_offsets.size - 1
    out_off = np.empty(n + 1, np.uint64)
    out_edges = np.empty(n * knn_size, np.uint32)
    secs = C.c_double(0)
    _chk(lib().gbdr_knn_cut(device, _ptr(knn_offsets), _ptr(knn_edges), _ptr(db), n, db.shape[1], knn_size,
                            _ptr(out_off), _ptr(out_edges), C.byref(secs)))
    return out_off, out_edges[: int(out_off[-1])].copy(), secs.value


def gd_prune_dev(device, d_knn, k, kstride, row_begin, row_end, d_db_low, n, d_low, M, d_fwd, d_deg, stream=0):
    """Forward lists of rows [row_begin, row_end) on device buffers (raw addresses)."""
    _chk(lib().gbdr_gd_prune_dev(device, d_knn, k, kstride, row_begin, row_end, d_db_low, n, d_low, M, d_fwd, d_deg,
                                 stream or None))


def gd_finish_dev(device, d_fwd, d_deg, n, M, reverse=True, need_const_degree=False, d_knn=0, k=0, kstride=0, stream=0):
    """Reverse pass (+ constant-degree fill) on the forward lists of all n vertices -> (offsets, edges)."""
    out_off = np.empty(n + 1, np.uint64)
    out_edges = np.empty(n * 2 * M, np.uint32)
    _chk(lib().gbdr_gd_finish_dev(device, d_fwd, d_deg, n, M, int(reverse), int(need_const_degree), d_knn or None, k,
                                  kstride, _ptr(out_off), _ptr(out_edges), stream or None))
    return out_off, out_edges[: int(out_off[-1])].copy()


def _build_graph(fn, head, db_low, knn_k, M, reverse, need_const_degree, knn_out):
    db_low = _f32(db_low)
    n, d_low = db_low.shape
    out_off = np.empty(n + 1, np.uint64)
    out_edges = np.empty(n * 2 * M, np.uint32)
    if knn_out is not None:
        assert knn_out.dtype == np.uint32 and knn_out.shape == (n, knn_k) and knn_out.flags.c_contiguous
    t = (C.c_double * 4)()
    _chk(fn(head, _ptr(db_low), n, d_low, knn_k, M, int(reverse), int(need_const_degree), _ptr(out_off), _ptr(out_edges),
            _ptr(knn_out), t))
    names = ("upload_s", "knn_s", "prune_s", "finish_s")
    return out_off, out_edges[: int(out_off[-1])].copy(), dict(zip(names, [float(x) for x in t]))


def build_graph(db_low, knn_k=1000, M=30, reverse=True, need_const_degree=False, device=0, knn_out=None):
    """gbdr_build_graph: kNN-`knn_k` self-join + hnswlikeGD without leaving HBM -> (offsets, edges, timings dict).
    knn_out: optional [n x knn_k] uint32 destination for the kNN lists (page-locked memory streams)."""
    return _build_graph(lib().gbdr_build_graph, device, db_low, knn_k, M, reverse, need_const_degree, knn_out)


class Group:
    """gbdr_group: several GPUs of one node driven by this process (replicated or sharded index)."""

    def __init__(self, devices, mode=GROUP_REPLICATED):
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        self._g = C.c_void_p()
        _chk(lib().gbdr_group_create(devs, len(devices), int(mode), C.byref(self._g)))
        self.devices = [int(d) for d in devices]
        self.mode = mode
        self._keep = []

    def close(self):
        if self._g:
            lib().gbdr_group_destroy(self._g)
            self._g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return len(self.devices)

    def member(self, i):
        """The per-device Index (borrowed: it dies with the group)."""
        h = C.c_void_p()
        _chk(lib().gbdr_group_member(self._g, i, C.byref(h)))
        ix = Index.__new__(Index)
        ix.device, ix._parent, ix._views, ix._inflight, ix.n, ix.d, ix.d_low = self.devices[i], None, [], None, 0, 0, 0
        ix._h, ix._owned = h, False  # close() / __del__ of the borrowed wrapper must not destroy the member
        return ix

    def set_exchange(self, exchange):
        _chk(lib().gbdr_group_set_exchange(self._g, int(exchange)))

    def set_net(self, l1, l2, l3):
        l1, l2, l3 = _f32(l1), _f32(l2), _f32(l3)
        _chk(lib().gbdr_group_set_net(self._g, _ptr(l1), _ptr(l2), _ptr(l3), l1.shape[1] - 1, l1.shape[0], l2.shape[0], l3.shape[0]))
        self.d, self.d_low = l1.shape[1] - 1, l3.shape[0]

    def set_base(self, db):
        db = _f32(db)
        _chk(lib().gbdr_group_set_base(self._g, _ptr(db), db.shape[0], db.shape[1]))

    def set_low(self, db_low):
        db_low = _f32(db_low)
        _chk(lib().gbdr_group_set_low(self._g, _ptr(db_low), db_low.shape[0], db_low.shape[1]))

    def set_graph(self, offsets, edges):
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        edges = _u32(edges)
        _chk(lib().gbdr_group_set_graph(self._g, _ptr(offsets), _ptr(edges), offsets.size - 1))

    def shard_rows(self, i):
        b, e = C.c_uint64(0), C.c_uint64(0)
        _chk(lib().gbdr_group_shard_rows(self._g, i, C.byref(b), C.byref(e)))
        return int(b.value), int(e.value)

    def set_shard_graph(self, i, offsets, edges):
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        edges = _u32(edges)
        _chk(lib().gbdr_group_set_shard_graph(self._g, i, _ptr(offsets), _ptr(edges), offsets.size - 1))

    def search(self, queries, q_low, ef, k, entry, flags=SEARCH_RERANK, out=None):
        """entry: [n_q] (replicated) or [n_devices, n_q] local entry vertices per shard (sharded)."""
        q = None if queries is None else _f32(queries)
        ql = None if q_low is None else _f32(q_low)
        entry = _u32(entry)
        n_q = entry.shape[-1]
        if out is None:
            out = dict(ids=np.empty((n_q, k), np.uint32), dists=np.empty((n_q, k), np.float32),
                       hops=np.empty(n_q, np.int32), dist_calc=np.empty(n_q, np.int32))
        secs = C.c_double(0)
        _chk(lib().gbdr_group_search(self._g, _ptr(q), _ptr(ql), n_q, ef, k, flags, _ptr(entry), _ptr(out["ids"]),
                                     _ptr(out["dists"]), _ptr(out["hops"]), _ptr(out["dist_calc"]), C.byref(secs)))
        out["gpu_seconds"] = secs.value
        return out

    def build_graph(self, db_low, knn_k=1000, M=30, reverse=True, need_const_degree=False, knn_out=None):
        """gbdr_group_build_graph: the graph build row-block sharded over the group -> (offsets, edges, timings)."""
        return _build_graph(lib().gbdr_group_build_graph, self._g, db_low, knn_k, M, reverse, need_const_degree, knn_out)


def merge_topk_dev(device, d_in_ids, d_in_dists, parts, n_q, k_in, k_out, d_out_ids, d_out_dists=0, stream=0):
    _chk(lib().gbdr_merge_topk_dev(device, d_in_ids, d_in_dists, parts, n_q, k_in, k_out, d_out_ids,
                                   d_out_dists or None, stream or None))


class DeviceBuffer:
    """Raw HBM allocation through the C ABI (for hosts that do not use torch)."""

    def __init__(self, nbytes, device=0):
        self.device = device
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        _chk(lib().gbdr_dev_malloc(device, self.nbytes, C.byref(p)))
        self.ptr = p.value

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        _chk(lib().gbdr_memcpy_h2d(self.device, self.ptr, _ptr(arr), arr.nbytes))
        return self

    def download(self, shape, dtype):
        out = np.empty(shape, dtype)
        assert out.nbytes <= self.nbytes
        _chk(lib().gbdr_memcpy_d2h(self.device, _ptr(out), self.ptr, out.nbytes))
        return out

    def free(self):
        if self.ptr:
            lib().gbdr_dev_free(self.device, self.ptr)
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def beam_plan_info(ef, dim, n_vertices, second_graph=False):
    """Launch plan of the beam-search kernel for this shape (host logic only, works without a GPU)."""
    out = (C.c_uint32 * 10)()
    _chk(lib().gbdr_beam_plan_info(int(ef), int(dim), int(n_vertices), int(bool(second_graph)), out))
    keys = ("variant", "cap", "warps_per_cta", "ctas_per_sm", "smem_per_warp", "vis_bytes", "vis_entries", "tag_bits",
            "disp_bits", "smem_per_sm")
    return dict(zip(keys, [int(x) for x in out]))


def synchronize(device=0):
    _chk(lib().gbdr_device_synchronize(device))
