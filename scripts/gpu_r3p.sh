# usage: bash scripts/gpu_r3p.sh TAG — L2 access-policy window on the searched vectors: A/B (bench + the beam kernel's L2 hit rate / DRAM bytes)
TAG=${1:-r3p}
mkdir -p gpurun_out
for m in 0 1 0 1; do
GBDR_L2_PERSIST=$m timeout 200 python bench.py --steps 40 --warmup 3 --ef 53 --no-cpu-baseline --no-ef-curve > gpurun_out/${TAG}_p${m}.json 2> gpurun_out/${TAG}_p${m}.log
python - <<P
import json
r=json.load(open("gpurun_out/${TAG}_p${m}.json"))
print("persist=$m: value %.2fM single %.2fM e2e %.2fM sync %.2fM beam %.4f ms frac %.3f rerank %.4f ms" % (r["value"]/1e6, r["single_stream"]["value"]/1e6, r["e2e"]["value"]/1e6, r["e2e"]["sync"]["value"]/1e6, r["roofline"]["kernel_ms"], r["roofline"]["frac"], r["roofline"]["other_kernels_ms"]["rerank"]))
P
done
for m in 0 1; do
GBDR_L2_PERSIST=$m timeout 200 ncu --metrics lts__t_sector_hit_rate.pct,dram__bytes_read.sum,gpu__time_duration.sum --clock-control none -k regex:beam_search -c 3 python bench.py --steps 2 --warmup 1 --ef 53 --no-cpu-baseline --no-ef-curve --in-flight 1 2>&1 | grep -E "beam_search|hit_rate|dram__bytes|time_duration" | tail -12
done
for m in 0 1; do
GBDR_L2_PERSIST=$m timeout 200 python bench.py --workload deep1m --steps 20 --warmup 3 --ef 294 --no-cpu-baseline --no-ef-curve > gpurun_out/${TAG}_d${m}.json 2> gpurun_out/${TAG}_d${m}.log
python - <<P
import json
r=json.load(open("gpurun_out/${TAG}_d${m}.json"))
print("deep1m persist=$m: value %.2fM single %.2fM beam %.3f ms" % (r["value"]/1e6, r["single_stream"]["value"]/1e6, r["roofline"]["kernel_ms"]))
P
done
