# usage: bash scripts/gpu_knn_prof.sh TAG — launch list of one 1M x 32, k=1000 kNN build (per-phase kernel times)
TAG=${1:-knnp}
mkdir -p gpurun_out
cat > /tmp/knn_one.py <<'PY'
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
from gbnns_dim_red_b200 import capi
n = int(os.environ.get("N", 1000000)); k = int(os.environ.get("K", 1000)); d = 32
rng = np.random.default_rng(d)
A = rng.standard_normal((8, d), dtype=np.float32)
x = rng.standard_normal((n, 8), dtype=np.float32) @ A + 0.1 * rng.standard_normal((n, d), dtype=np.float32)
x /= np.linalg.norm(x, axis=1, keepdims=True)
x = np.ascontiguousarray(x, dtype=np.float32)
ids, secs = capi.knn(x, x, k)
print(f"knn n={n} d={d} k={k}: gpu {secs:.3f} s", flush=True)
PY
N=1000000 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_knn_launches.csv python /tmp/knn_one.py > gpurun_out/${TAG}_knn_launches.log 2>&1
cat gpurun_out/${TAG}_knn_launches.log
N=1000000 timeout 300 ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_knn_tc python /tmp/knn_one.py > gpurun_out/${TAG}_ncu_knn.log 2>&1
