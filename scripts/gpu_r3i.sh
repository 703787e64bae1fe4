# usage: bash scripts/gpu_r3i.sh TAG — whole GPU suite, then the single-GPU shapes after the plan changes (visited-count feedback,
# 96-register builds of the 192..320-slot lists, tag format on large shards)
TAG=${1:-r3i}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.txt
tail -8 gpurun_out/${TAG}_pytest.txt
timeout 300 python bench.py --steps 20 --warmup 3 --ef 53 --no-cpu-baseline --efs 53,100,120,140,152,153,184,185,200,248,249,294,312,400 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.log; echo "bench rc=$?"
grep "ef curve" gpurun_out/${TAG}_bench.log
GBDR_BEAM_FEEDBACK=0 timeout 300 python bench.py --steps 20 --warmup 3 --ef 53 --no-cpu-baseline --efs 53,100,140,153,200,294 > gpurun_out/${TAG}_bench_nofb.json 2> gpurun_out/${TAG}_bench_nofb.log
grep "ef curve" gpurun_out/${TAG}_bench_nofb.log
for fb in 1 0; do
GBDR_BEAM_FEEDBACK=$fb timeout 300 python bench.py --workload deep1m --steps 20 --warmup 3 --no-cpu-baseline --no-ef-curve > gpurun_out/${TAG}_deep1m_fb$fb.json 2> gpurun_out/${TAG}_deep1m_fb$fb.log
python - <<P
import json
r=json.load(open("gpurun_out/${TAG}_deep1m_fb$fb.json"))
print("deep1m feedback=$fb: ef %d value %.2fM single %.2fM e2e %.2fM beam %.3f ms frac %.3f" % (r["config"]["ef"], r["value"]/1e6, r["single_stream"]["value"]/1e6, r["e2e"]["value"]/1e6, r["roofline"]["kernel_ms"], r["roofline"]["frac"]))
P
done
timeout 300 python bench.py --workload gist1m --steps 20 --warmup 3 --no-cpu-baseline --no-ef-curve > gpurun_out/${TAG}_gist1m.json 2> gpurun_out/${TAG}_gist1m.log
python - <<P
import json
r=json.load(open("gpurun_out/${TAG}_gist1m.json"))
print("gist1m: ef %d value %.2fM single %.2fM e2e %.2fM beam %.3f ms frac %.3f" % (r["config"]["ef"], r["value"]/1e6, r["single_stream"]["value"]/1e6, r["e2e"]["value"]/1e6, r["roofline"]["kernel_ms"], r["roofline"]["frac"]))
P
( time timeout 900 python bench.py --workload deep-sharded --shard-n 12500000 --steps 10 --warmup 3 ) > gpurun_out/${TAG}_shard12m.json 2> gpurun_out/${TAG}_shard12m.log; echo "shard12m rc=$?"
python - <<P
import json
r=json.load(open("gpurun_out/${TAG}_shard12m.json"))
print("shard12m: ef %d value %.2fM e2e %.2fM beam %.3f ms frac %.3f" % (r["config"]["ef"], r["value"]/1e6, r["e2e"]["value"]/1e6, r["roofline"]["kernel_ms"], r["roofline"]["frac"]))
P
