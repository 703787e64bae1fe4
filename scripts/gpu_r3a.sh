# usage: bash scripts/gpu_r3a.sh TAG — round-3 first pass: GPU tests, then A/B of the beam-kernel tuning flags at the
# SIFT-1M operating point (ef 53), the ef curve with the finer list capacities, and the Deep-1M point
TAG=${1:-r3a}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.txt
tail -5 gpurun_out/${TAG}_pytest.txt
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 40 --warmup 3 --ef 53 --no-cpu-baseline --no-ef-curve > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.log
  python - <<P
import json
try:
    r=json.load(open("gpurun_out/${TAG}_${name}.json"))
    print("${name}: value %.2fM single %.2fM e2e %.2fM sync %.2fM beam %.4f ms" % (r["value"]/1e6, r["single_stream"]["value"]/1e6, r["e2e"]["value"]/1e6, r["e2e"]["sync"]["value"]/1e6, r["roofline"]["kernel_ms"]))
except Exception as e:
    print("${name}: failed", e)
P
}
run pf1 GBDR_BEAM_PF_ROWS=1
run pf3 GBDR_BEAM_PF_ROWS=3
run pf7 GBDR_BEAM_PF_ROWS=7
run pf5 GBDR_BEAM_PF_ROWS=5
run dense_pf3 GBDR_BEAM_DENSE=1 GBDR_BEAM_PF_ROWS=3
run dense_pf7 GBDR_BEAM_DENSE=1 GBDR_BEAM_PF_ROWS=7
run pf3_again GBDR_BEAM_PF_ROWS=3
# ef curve (default flags) and deep1m
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.log; echo "bench rc=$?"
grep "ef curve" gpurun_out/${TAG}_bench.log
timeout 500 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload deep1m > gpurun_out/${TAG}_deep1m.json 2> gpurun_out/${TAG}_deep1m.log; echo "deep1m rc=$?"
grep -E "operating|ef curve" gpurun_out/${TAG}_deep1m.log
python - <<P
import json
for n in ("bench","deep1m"):
    try:
        r=json.load(open("gpurun_out/${TAG}_%s.json"%n)); print(n, "value %.2fM e2e %.2fM ef %d frac %.3f beam %.3f ms"%(r["value"]/1e6, r["e2e"]["value"]/1e6, r["config"]["ef"], r["roofline"]["frac"], r["roofline"]["kernel_ms"]))
    except Exception as e: print(n,"failed",e)
P
