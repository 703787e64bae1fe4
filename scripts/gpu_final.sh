# usage: bash scripts/gpu_final.sh TAG — what the driver runs at round end, in its order (GPU tests, smoke(), both bench arms), plus
# the launch list and the ncu --set full capture of the beam kernel that profiles/ cites, the other single-GPU shapes, the
# drop-in final_test binary on the bench workload, and a compute-sanitizer pass over smoke()
TAG=${1:-r4z}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.txt
tail -3 gpurun_out/${TAG}_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.txt
timeout 400 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.log; echo "reference rc=$?"
timeout 400 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.log; echo "bench rc=$?"
EF=$(python -c "import json;print(json.load(open('gpurun_out/${TAG}_bench.json'))['config']['ef'])")
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --ef $EF --no-cpu-baseline --no-ef-curve --in-flight 1 > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:beam_search -c 1 -f -o gpurun_out/${TAG}_beam python bench.py --steps 1 --warmup 0 --ef $EF --no-cpu-baseline --no-ef-curve --in-flight 1 > gpurun_out/${TAG}_ncu_full.log 2>&1
python - <<PY
import json
for f in ("bench_reference", "bench"):
    j = json.loads([l for l in open(f"gpurun_out/${TAG}_{f}.json") if l.startswith("{")][-1])
    print(f, "value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "steps", j["steps"], "ms/step", round(j["ms_per_step"], 4),
          "frac", j.get("roofline", {}).get("frac"), "launches", j.get("gpu_launches"), "clocks", j.get("clocks"), "build", j.get("build"))
PY
for wl in deep1m gist1m; do
timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 > gpurun_out/${TAG}_${wl}.json 2> gpurun_out/${TAG}_${wl}.log
python - <<PY
import json
r=json.load(open("gpurun_out/${TAG}_${wl}.json"))
print("${wl}: ef %d value %.2fM single %.2fM e2e %.2fM beam %.3f ms frac %.3f cpu %s" % (r["config"]["ef"], r["value"]/1e6, r["single_stream"]["value"]/1e6, r["e2e"]["value"]/1e6, r["roofline"]["kernel_ms"], r["roofline"]["frac"], (r.get("cpu_baseline") or {}).get("value")))
PY
done
timeout 400 python scripts/final_test_probe.py gpurun_out/${TAG}_final_test.json > gpurun_out/${TAG}_final_test.txt 2>&1; tail -7 gpurun_out/${TAG}_final_test.txt
timeout 400 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_memcheck_smoke.txt 2>&1; tail -3 gpurun_out/${TAG}_memcheck_smoke.txt
