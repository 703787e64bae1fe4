# usage: bash scripts/gpu_r3q.sh TAG — projection with programmatic dependent launch between its layers (A/B), 32-bit select; parity first
TAG=${1:-r3q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_build_ops.py tests/test_gpu_search.py tests/test_gpu_golden.py -q -m gpu > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.txt
tail -4 gpurun_out/${TAG}_pytest.txt
for m in 0 1 0 1; do
GBDR_PROJ_PDL=$m timeout 200 python bench.py --steps 40 --warmup 3 --ef 53 --no-cpu-baseline --no-ef-curve > gpurun_out/${TAG}_pdl${m}.json 2> gpurun_out/${TAG}_pdl${m}.log
python - <<P
import json
r=json.load(open("gpurun_out/${TAG}_pdl${m}.json"))
print("pdl=$m: value %.2fM single %.2fM e2e %.2fM sync %.2fM beam %.4f ms project %.4f ms" % (r["value"]/1e6, r["single_stream"]["value"]/1e6, r["e2e"]["value"]/1e6, r["e2e"]["sync"]["value"]/1e6, r["roofline"]["kernel_ms"], r["roofline"]["other_kernels_ms"]["project"]))
P
done
for m in 0 1; do
GBDR_PROJ_PDL=$m timeout 200 python bench.py --workload gist1m --steps 20 --warmup 3 --no-cpu-baseline --no-ef-curve > gpurun_out/${TAG}_g${m}.json 2> gpurun_out/${TAG}_g${m}.log
python - <<P
import json
r=json.load(open("gpurun_out/${TAG}_g${m}.json"))
print("gist1m pdl=$m: value %.2fM single %.2fM e2e %.2fM project %.4f ms" % (r["value"]/1e6, r["single_stream"]["value"]/1e6, r["e2e"]["value"]/1e6, r["roofline"]["other_kernels_ms"]["project"]))
P
done
