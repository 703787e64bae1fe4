# usage: bash scripts/gpu_r1w.sh TAG — full GPU suite (new GD reverse pass), bench with a fresh workload build (GD stage timing), sharded bench on 1 GPU
TAG=${1:-r1w}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.txt
tail -4 gpurun_out/${TAG}_pytest.txt
rm -rf /tmp/gbdr_bench_cache
GBDR_GD_TIMING=1 timeout 600 python bench.py --steps 60 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.log; echo "bench rc=$?"
grep -v "ef=" gpurun_out/${TAG}_bench.log | tail -14
timeout 600 python bench.py --workload deep-sharded --steps 10 --warmup 3 > gpurun_out/${TAG}_sharded_n1.json 2> gpurun_out/${TAG}_sharded_n1.log; echo "sharded n1 rc=$?"
tail -2 gpurun_out/${TAG}_sharded_n1.log
python - <<PY
import json
for f in ("bench", "sharded_n1"):
    try:
        j = json.loads([l for l in open(f"gpurun_out/${TAG}_{f}.json") if l.startswith("{")][-1])
        print(f, "value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "ms/step", round(j["ms_per_step"], 3), "ef", j["config"]["ef"], j.get("build"), j.get("sharded_knn_build"), j["roofline"]["kernel_ms"], j["roofline"]["other_kernels_ms"])
    except Exception as e:
        print(f, "failed", e)
PY
