# usage: bash scripts/gpu_r4e.sh TAG N — only the sharded-build legs on N GPUs, twice (is the slow N = 4 build of run r4d the box or the code?)
TAG=${1:-r4e}; N=${2:-4}
mkdir -p gpurun_out
uptime > gpurun_out/${TAG}_host.txt; nvidia-smi --query-gpu=index,utilization.gpu,memory.used,clocks.sm --format=csv >> gpurun_out/${TAG}_host.txt
cat gpurun_out/${TAG}_host.txt
for rep in 1 2; do
timeout 300 python - <<P
import time, numpy as np, sys
sys.path.insert(0, ".")
from gbnns_dim_red_b200 import capi
rng = np.random.default_rng(1)
Y = rng.standard_normal((1_000_000, 32), dtype=np.float32); Y /= np.linalg.norm(Y, axis=1, keepdims=True)
g = capi.Group(list(range($N)), capi.GROUP_REPLICATED)
g.build_graph(Y[:65536], knn_k=64, M=30)
for ex, name in ((capi.EXCHANGE_PEER, "peer"), (capi.EXCHANGE_NCCL, "nccl")):
    g.set_exchange(ex)
    t0 = time.perf_counter(); off, ed, t = g.build_graph(Y, knn_k=1000, M=30); print("rep $rep", name, {k: round(v, 3) for k, v in t.items()}, "wall %.3f" % (time.perf_counter() - t0), flush=True)
g.close()
t0 = time.perf_counter(); off, ed, t = capi.build_graph(Y, knn_k=1000, M=30); print("rep $rep one GPU", {k: round(v, 3) for k, v in t.items()}, "wall %.3f" % (time.perf_counter() - t0), flush=True)
P
done
