"""Where the seconds of the graph build go: gbdr_knn (CUDA events) vs the HBM-resident chain gbdr_build_graph (host clock per
stage), first and second call in one process.  usage: python scripts/build_probe.py [n]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from gbnns_dim_red_b200 import capi, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
rng = np.random.default_rng(1)
Y = rng.standard_normal((n, 32), dtype=np.float32)
Y /= np.linalg.norm(Y, axis=1, keepdims=True)
pinned = capi.PinnedArray((n, 1000), np.uint32)
for rep in range(2):
    t0 = time.time()
    ids, gpu_s = capi.knn(Y, Y, 1000, out_ids=pinned.array)
    print(f"gbdr_knn          rep {rep}: gpu {gpu_s:.3f} s, wall {time.time() - t0:.3f} s", flush=True)
for rep in range(2):
    t0 = time.time()
    off, ed, t = capi.build_graph(Y, knn_k=1000, M=30, reverse=True, knn_out=pinned.array)
    print(f"gbdr_build_graph  rep {rep}: {t}, wall {time.time() - t0:.3f} s", flush=True)
for rep in range(2):
    t0 = time.time()
    off, ed, t = capi.build_graph(Y, knn_k=1000, M=30, reverse=True)
    print(f"gbdr_build_graph (no knn_out) rep {rep}: {t}, wall {time.time() - t0:.3f} s", flush=True)
