# usage: bash scripts/gpu_r3b.sh TAG — GPU tests (new: groups on one GPU, HBM-resident graph build, entry validation), a fine ef
# sweep around the list-capacity steps, the Deep-100M shard size on one GPU, and one ncu --set full capture of the beam kernel
TAG=${1:-r3b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.txt
tail -8 gpurun_out/${TAG}_pytest.txt
EFS=100,120,121,140,152,153,160,184,185,200,248,249,280,294,312,313,376,400
timeout 400 python bench.py --steps 20 --warmup 3 --ef 53 --no-cpu-baseline --efs $EFS > gpurun_out/${TAG}_sweep.json 2> gpurun_out/${TAG}_sweep.log; echo "sweep rc=$?"
grep "ef curve" gpurun_out/${TAG}_sweep.log
GBDR_BEAM_WPB=8 timeout 400 python bench.py --steps 20 --warmup 3 --ef 53 --no-cpu-baseline --efs 153,160,184,294,312 > gpurun_out/${TAG}_sweep_w8.json 2> gpurun_out/${TAG}_sweep_w8.log
grep "ef curve" gpurun_out/${TAG}_sweep_w8.log
python - <<P
import json
r=json.load(open("gpurun_out/${TAG}_sweep.json"))
print("default: value %.2fM single %.2fM e2e %.2fM sync %.2fM beam %.4f ms frac %.3f" % (r["value"]/1e6, r["single_stream"]["value"]/1e6, r["e2e"]["value"]/1e6, r["e2e"]["sync"]["value"]/1e6, r["roofline"]["kernel_ms"], r["roofline"]["frac"]))
P
( time timeout 900 python bench.py --workload deep-sharded --shard-n 12500000 --steps 10 --warmup 3 ) > gpurun_out/${TAG}_shard12m.json 2> gpurun_out/${TAG}_shard12m.log; echo "shard12m rc=$?"
grep -E "shard|operating|real" gpurun_out/${TAG}_shard12m.log | tail -12
timeout 300 ncu --set full --clock-control none --import-source on -k regex:beam_search -c 1 -f -o gpurun_out/${TAG}_beam python bench.py --steps 1 --warmup 0 --ef 53 --no-cpu-baseline --no-ef-curve > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu rc=$?"
