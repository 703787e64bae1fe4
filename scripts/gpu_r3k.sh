# usage: bash scripts/gpu_r3k.sh TAG — workload builder on the HBM-resident chain (full-size + drop-in tests), dense CTAs of 17 vs 18 warps
TAG=${1:-r3k}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_host_dropin.py tests/test_gpu_fullsize.py -q -m gpu > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.txt
tail -6 gpurun_out/${TAG}_pytest.txt
for w in 17 18 17 18; do
GBDR_BEAM_DENSE_WARPS=$w timeout 200 python bench.py --steps 40 --warmup 3 --ef 53 --no-cpu-baseline --no-ef-curve > gpurun_out/${TAG}_dw${w}.json 2> gpurun_out/${TAG}_dw${w}.log
python - <<P
import json
r=json.load(open("gpurun_out/${TAG}_dw${w}.json"))
print("dense warps $w: value %.2fM single %.2fM e2e %.2fM sync %.2fM beam %.4f ms frac %.3f build %s" % (r["value"]/1e6, r["single_stream"]["value"]/1e6, r["e2e"]["value"]/1e6, r["e2e"]["sync"]["value"]/1e6, r["roofline"]["kernel_ms"], r["roofline"]["frac"], r["build"]))
P
done
timeout 300 python bench.py --workload deep1m --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_deep1m.json 2> gpurun_out/${TAG}_deep1m.log
python - <<P
import json
r=json.load(open("gpurun_out/${TAG}_deep1m.json"))
print("deep1m: ef %d value %.2fM single %.2fM e2e %.2fM beam %.3f ms frac %.3f" % (r["config"]["ef"], r["value"]/1e6, r["single_stream"]["value"]/1e6, r["e2e"]["value"]/1e6, r["roofline"]["kernel_ms"], r["roofline"]["frac"]))
P
grep "ef curve" gpurun_out/${TAG}_deep1m.log
