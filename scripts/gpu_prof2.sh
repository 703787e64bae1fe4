# usage: bash scripts/gpu_prof2.sh TAG — ncu evidence for the tensor-core kernels (projection, kNN build) and re-rank
TAG=${1:-prof2}
mkdir -p gpurun_out
export N=${N:-300000} K=${K:-1000}
cat > /tmp/knn_one.py <<'PY'
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
from gbnns_dim_red_b200 import capi
n = int(os.environ.get("N", 300000)); k = int(os.environ.get("K", 1000)); d = 32
rng = np.random.default_rng(d)
A = rng.standard_normal((8, d), dtype=np.float32)
x = rng.standard_normal((n, 8), dtype=np.float32) @ A + 0.1 * rng.standard_normal((n, d), dtype=np.float32)
x /= np.linalg.norm(x, axis=1, keepdims=True)
x = np.ascontiguousarray(x, dtype=np.float32)
ids, secs = capi.knn(x, x, k)
print(f"knn n={n} d={d} k={k}: gpu {secs:.3f} s", flush=True)
PY
# launch list of one kNN build (per-kernel times)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_knn_launches.csv python /tmp/knn_one.py > gpurun_out/${TAG}_knn_launches.log 2>&1
# full captures: phase-1 scan launch of knn_tc_kernel (2nd launch of that name), the select kernel, a projection layer, re-rank
timeout 300 ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_knn_tc python /tmp/knn_one.py > gpurun_out/${TAG}_ncu_knn.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:knn_select -c 1 -f -o gpurun_out/${TAG}_knn_select python /tmp/knn_one.py >> gpurun_out/${TAG}_ncu_knn.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:linear_tc -s 3 -c 3 -f -o gpurun_out/${TAG}_linear_tc python bench.py --steps 1 --warmup 0 --ef 53 --no-cpu-baseline > gpurun_out/${TAG}_ncu_lin.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rerank -c 1 -f -o gpurun_out/${TAG}_rerank python bench.py --steps 1 --warmup 0 --ef 53 --no-cpu-baseline >> gpurun_out/${TAG}_ncu_lin.log 2>&1
N=1000000 timeout 200 python scripts/knn_probe.py > gpurun_out/${TAG}_knn_probe.txt 2>&1
tail -5 gpurun_out/${TAG}_knn_probe.txt
ls -la gpurun_out | tail -12
