# usage: bash scripts/gpu_r4c.sh TAG — CUDA-graph replay of repeated host-buffer calls: parity, then A/B on the end-to-end legs
TAG=${1:-r4c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_fullsize.py tests/test_gpu_group.py tests/test_host_dropin.py -q -m gpu -x > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.txt
tail -5 gpurun_out/${TAG}_pytest.txt
for m in 0 1 0 1; do
GBDR_SEARCH_GRAPH=$m timeout 200 python bench.py --workload gist1m --steps 40 --warmup 3 --no-cpu-baseline --no-ef-curve > gpurun_out/${TAG}_g${m}.json 2> gpurun_out/${TAG}_g${m}.log
python - <<P
import json
r=json.load(open("gpurun_out/${TAG}_g${m}.json"))
print("gist1m graph=$m: value %.2fM single %.2fM e2e %.2fM sync %.2fM" % (r["value"]/1e6, r["single_stream"]["value"]/1e6, r["e2e"]["value"]/1e6, r["e2e"]["sync"]["value"]/1e6))
P
done
for m in 0 1 0 1; do
GBDR_SEARCH_GRAPH=$m timeout 200 python bench.py --steps 40 --warmup 3 --ef 53 --no-cpu-baseline --no-ef-curve > gpurun_out/${TAG}_s${m}.json 2> gpurun_out/${TAG}_s${m}.log
python - <<P
import json
r=json.load(open("gpurun_out/${TAG}_s${m}.json"))
print("sift1m graph=$m: value %.2fM single %.2fM e2e %.2fM sync %.2fM launches %s" % (r["value"]/1e6, r["single_stream"]["value"]/1e6, r["e2e"]["value"]/1e6, r["e2e"]["sync"]["value"]/1e6, r["gpu_launches"]))
P
done
