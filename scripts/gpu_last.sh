# usage: bash scripts/gpu_last.sh TAG — the driver's round-end sequence on the final commit: GPU suite (-x), smoke(), both bench arms
TAG=${1:-r4y}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.txt
tail -3 gpurun_out/${TAG}_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.txt
timeout 400 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.log; echo "reference rc=$?"
timeout 400 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.log; echo "bench rc=$?"
python - <<PY
import json
for f in ("bench_reference", "bench"):
    j = json.loads([l for l in open(f"gpurun_out/${TAG}_{f}.json") if l.startswith("{")][-1])
    print(f, "value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "ms/step", round(j["ms_per_step"], 4), "frac", j.get("roofline", {}).get("frac"), "launches", j.get("gpu_launches"), "knn_s", (j.get("build") or {}).get("knn_build_s"))
PY
