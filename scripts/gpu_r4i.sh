# usage: bash scripts/gpu_r4i.sh TAG — dense build with compile-time tuning flags: parity, then bench x2
TAG=${1:-r4i}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_search.py tests/test_gpu_golden.py -q -m gpu > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.txt
tail -3 gpurun_out/${TAG}_pytest.txt
for i in 1 2; do
timeout 200 python bench.py --steps 40 --warmup 3 --ef 53 --no-cpu-baseline --no-ef-curve > gpurun_out/${TAG}_b$i.json 2> gpurun_out/${TAG}_b$i.log
python - <<P
import json
r=json.load(open("gpurun_out/${TAG}_b$i.json"))
print("run $i: value %.2fM single %.2fM e2e %.2fM sync %.2fM beam %.4f ms frac %.3f" % (r["value"]/1e6, r["single_stream"]["value"]/1e6, r["e2e"]["value"]/1e6, r["e2e"]["sync"]["value"]/1e6, r["roofline"]["kernel_ms"], r["roofline"]["frac"]))
P
done
