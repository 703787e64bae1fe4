# usage: bash scripts/gpu_dbg2.sh TAG — kNN + search tests with tight timeouts, kNN timing, deep1m bench
TAG=${1:-dbg}
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_build_ops.py tests/test_gpu_search.py tests/test_gpu_golden.py -m gpu -x -q --timeout 120 > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.txt
tail -15 gpurun_out/${TAG}_pytest.txt
N=1000000 timeout 200 python scripts/knn_probe.py > gpurun_out/${TAG}_knn_probe.txt 2>&1
cat gpurun_out/${TAG}_knn_probe.txt
timeout 300 python bench.py --workload deep1m --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_deep1m.json 2> gpurun_out/${TAG}_deep1m.log; echo "deep rc=$?"
tail -2 gpurun_out/${TAG}_deep1m.log
python - <<PY
import json
j = json.loads([l for l in open("gpurun_out/${TAG}_deep1m.json") if l.startswith("{")][-1])
print("deep1m value", round(j["value"]), "ef", j["config"]["ef"], "kernel_ms", j["roofline"]["kernel_ms"], j["roofline"]["other_kernels_ms"], "frac", j["roofline"]["frac"])
PY
