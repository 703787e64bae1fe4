"""kNN-build timing probe (GPU): 1M x d unit vectors with an 8-dim latent, k = 1000, both variants."""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
from gbnns_dim_red_b200 import capi
n = int(os.environ.get("N", 1000000)); k = int(os.environ.get("K", 1000))
for d in (32, 16, 64):
    rng = np.random.default_rng(d)
    A = rng.standard_normal((8, d), dtype=np.float32)
    x = rng.standard_normal((n, 8), dtype=np.float32) @ A + 0.1 * rng.standard_normal((n, d), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    x = np.ascontiguousarray(x, dtype=np.float32)
    res = {}
    for variant in ("tc", "scan"):
        if variant == "scan" and d != 32:
            continue
        os.environ["GBDR_KNN_VARIANT"] = variant
        t0 = time.time()
        ids, secs = capi.knn(x, x, k)
        print(f"d={d} n={n} k={k} variant={variant}: gpu {secs:.3f} s, wall {time.time()-t0:.2f} s", flush=True)
        res[variant] = ids
    if len(res) == 2:
        print("   variants agree:", bool(np.array_equal(res["tc"], res["scan"])), flush=True)
