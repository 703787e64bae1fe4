# usage: bash scripts/gpu_r4h.sh TAG — the two bench arms with default flags (what the driver runs), for profiles/
TAG=${1:-r4h}
mkdir -p gpurun_out
timeout 400 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.log; echo "reference rc=$?"
timeout 400 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.log; echo "bench rc=$?"
python - <<PY
import json
for f in ("bench_reference", "bench"):
    j = json.loads([l for l in open(f"gpurun_out/${TAG}_{f}.json") if l.startswith("{")][-1])
    print(f, "value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "ms/step", round(j["ms_per_step"], 4), "frac", j.get("roofline", {}).get("frac"), "traffic", j.get("roofline", {}).get("traffic"), "build", j.get("build"), "cpu", (j.get("cpu_baseline") or {}).get("value"))
PY
