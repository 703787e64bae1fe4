# usage: bash scripts/gpu_r3g.sh TAG — build-op / search / drop-in / group tests, default bench, blocking-call split sweep, 12.5 M-row shard
TAG=${1:-r3g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_build_ops.py tests/test_gpu_search.py tests/test_host_dropin.py tests/test_gpu_group.py -q -m gpu > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.txt
tail -12 gpurun_out/${TAG}_pytest.txt
timeout 300 python bench.py --steps 20 --warmup 3 --ef 53 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.log; echo "bench rc=$?"
grep "ef curve" gpurun_out/${TAG}_bench.log
for sp in 1 2 3 4; do
GBDR_SEARCH_SPLIT=$sp timeout 200 python bench.py --steps 20 --warmup 3 --ef 53 --no-cpu-baseline --no-ef-curve > gpurun_out/${TAG}_split${sp}.json 2> gpurun_out/${TAG}_split${sp}.log
python - <<P
import json
r=json.load(open("gpurun_out/${TAG}_split${sp}.json"))
print("split=${sp}: value %.2fM single %.2fM e2e %.2fM sync %.2fM beam %.4f ms frac %.3f" % (r["value"]/1e6, r["single_stream"]["value"]/1e6, r["e2e"]["value"]/1e6, r["e2e"]["sync"]["value"]/1e6, r["roofline"]["kernel_ms"], r["roofline"]["frac"]))
P
done
( time timeout 900 python bench.py --workload deep-sharded --shard-n 12500000 --steps 10 --warmup 3 ) > gpurun_out/${TAG}_shard12m.json 2> gpurun_out/${TAG}_shard12m.log; echo "shard12m rc=$?"
grep -E "shard|operating|real" gpurun_out/${TAG}_shard12m.log | tail -8
