import os, sys, time
import numpy as np
sys.path.insert(0, ".")
from gbnns_dim_red_b200 import capi
n = int(os.environ.get("N", 200000)); k = int(os.environ.get("K", 1000)); d = int(os.environ.get("D", 32))
rng = np.random.default_rng(d)
A = rng.standard_normal((8, d), dtype=np.float32)
x = rng.standard_normal((n, 8), dtype=np.float32) @ A + 0.1 * rng.standard_normal((n, d), dtype=np.float32)
x /= np.linalg.norm(x, axis=1, keepdims=True)
x = np.ascontiguousarray(x, dtype=np.float32)
nq = int(os.environ.get("NQ", n))
t0 = time.time(); ids, secs = capi.knn(x[:nq], x, k); print(f"n={n} nq={nq} k={k} d={d}: gpu {secs:.3f}s wall {time.time()-t0:.2f}s", flush=True)
