# usage: bash scripts/gpu_r3j.sh TAG — visited-table load factor x feedback sweep (plan only: results are exact either way)
TAG=${1:-r3j}
mkdir -p gpurun_out
for fb in 0 1; do for load in 75 88 100; do
GBDR_BEAM_FEEDBACK=$fb GBDR_BEAM_LOAD_PCT=$load timeout 200 python bench.py --steps 20 --warmup 3 --ef 53 --no-cpu-baseline --efs 53,100,120,152,184,248,294 > gpurun_out/${TAG}_s_fb${fb}_l${load}.json 2> gpurun_out/${TAG}_s_fb${fb}_l${load}.log
echo "sift fb=$fb load=$load $(grep 'ef curve' gpurun_out/${TAG}_s_fb${fb}_l${load}.log)"
GBDR_BEAM_FEEDBACK=$fb GBDR_BEAM_LOAD_PCT=$load timeout 200 python bench.py --workload deep1m --steps 10 --warmup 3 --ef 294 --no-cpu-baseline --no-ef-curve > gpurun_out/${TAG}_d_fb${fb}_l${load}.json 2> gpurun_out/${TAG}_d_fb${fb}_l${load}.log
python - <<P
import json
r=json.load(open("gpurun_out/${TAG}_d_fb${fb}_l${load}.json"))
print("deep1m fb=$fb load=$load: value %.2fM single %.2fM beam %.3f ms" % (r["value"]/1e6, r["single_stream"]["value"]/1e6, r["roofline"]["kernel_ms"]))
P
done; done
