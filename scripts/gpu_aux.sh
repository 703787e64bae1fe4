# usage: bash scripts/gpu_aux.sh TAG — the other single-GPU BASELINE.json shapes (parity-test cases, reported in DESIGN.md)
TAG=${1:-aux}
mkdir -p gpurun_out
for W in deep1m gist1m; do
  timeout 600 python bench.py --workload $W --steps 20 --warmup 3 > gpurun_out/${TAG}_${W}.json 2> gpurun_out/${TAG}_${W}.log; echo "$W rc=$?"
  tail -4 gpurun_out/${TAG}_${W}.log
  python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/${TAG}_${W}.json") if l.startswith("{")][-1])
    print("$W", "value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "ef", j["config"]["ef"], "frac", round(j["roofline"]["frac"], 3),
          "kernel_ms", j["roofline"]["kernel_ms"], j["roofline"]["other_kernels_ms"], "cpu", j.get("cpu_baseline", {}).get("value"), "build", j["build"])
except Exception as e:
    print("$W failed", e)
PY
done
