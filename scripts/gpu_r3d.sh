# usage: bash scripts/gpu_r3d.sh TAG — pair kernel (two walks per warp): parity tests with it forced on, then A/B bench
TAG=${1:-r3e}
mkdir -p gpurun_out
GBDR_BEAM_PAIR=1 timeout 1200 python -m pytest tests/test_gpu_search.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py -x -q > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.txt
tail -25 gpurun_out/${TAG}_pytest.txt
for pair in 1 0; do
GBDR_BEAM_PAIR=$pair timeout 400 python bench.py --steps 20 --warmup 3 --ef 53 --no-cpu-baseline --efs 20,40,53,56,60,80,100,120 > gpurun_out/${TAG}_pair${pair}.json 2> gpurun_out/${TAG}_pair${pair}.log; echo "bench pair=$pair rc=$?"
grep "ef curve" gpurun_out/${TAG}_pair${pair}.log
python - <<P
import json
r=json.load(open("gpurun_out/${TAG}_pair${pair}.json"))
print("pair=${pair}: value %.2fM single %.2fM e2e %.2fM sync %.2fM beam %.4f ms frac %.3f" % (r["value"]/1e6, r["single_stream"]["value"]/1e6, r["e2e"]["value"]/1e6, r["e2e"]["sync"]["value"]/1e6, r["roofline"]["kernel_ms"], r["roofline"]["frac"]))
P
done
GBDR_BEAM_PAIR=1 timeout 300 python bench.py --workload deep1m --steps 10 --warmup 3 --no-cpu-baseline --efs 40,80,100,120 > gpurun_out/${TAG}_deep1m.json 2> gpurun_out/${TAG}_deep1m.log; grep "ef curve" gpurun_out/${TAG}_deep1m.log
GBDR_BEAM_PAIR=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:beam_search -c 1 -f -o gpurun_out/${TAG}_beam python bench.py --steps 1 --warmup 0 --ef 53 --no-cpu-baseline --no-ef-curve > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu rc=$?"
