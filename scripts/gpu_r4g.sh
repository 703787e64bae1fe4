# usage: bash scripts/gpu_r4g.sh TAG — is the graph-build time of the bench line stable now (pool keeps its memory)?  3 bench runs + the probe
TAG=${1:-r4g}
mkdir -p gpurun_out
for i in 1 2 3; do
timeout 300 python bench.py --steps 20 --warmup 3 --ef 53 --no-cpu-baseline --no-ef-curve > gpurun_out/${TAG}_b$i.json 2> gpurun_out/${TAG}_b$i.log
python - <<P
import json
r=json.load(open("gpurun_out/${TAG}_b$i.json"))
b=r["build"]; print("run $i: value %.2fM e2e %.2fM  knn %.3f s prune %.3f s wall %.3f s identical %s" % (r["value"]/1e6, r["e2e"]["value"]/1e6, b["knn_build_s"], b["gd_prune_s"], b["wall_s"], b["graph_identical_to_the_workload_graph"]))
P
done
timeout 300 python scripts/build_probe.py 2>&1 | tee gpurun_out/${TAG}_build_probe.txt
timeout 600 python -m pytest tests/test_gpu_build_ops.py -q -m gpu 2>&1 | tail -2
