# usage: bash scripts/gpu_r3f.sh TAG — pair kernel: parity tests with it forced on, then A/B bench (short timeouts: a hang must not eat the budget)
TAG=${1:-r3f}
mkdir -p gpurun_out
GBDR_BEAM_PAIR=1 timeout 600 python -m pytest tests/test_gpu_search.py tests/test_gpu_golden.py -x -q > gpurun_out/${TAG}_pytest.txt 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/${TAG}_pytest.txt
tail -12 gpurun_out/${TAG}_pytest.txt
if [ $rc -ne 0 ]; then
  GBDR_BEAM_PAIR=1 timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_search.py -x -q -k "tagged_visited_table_overflow" > gpurun_out/${TAG}_memcheck.txt 2>&1
  grep -E "Invalid|at |by thread|Address|beam_search" gpurun_out/${TAG}_memcheck.txt | head -40
  exit 0
fi
GBDR_BEAM_PAIR=1 timeout 150 python bench.py --steps 20 --warmup 3 --ef 53 --no-cpu-baseline --efs 20,40,53,56,60,80,100,120 > gpurun_out/${TAG}_pair1.json 2> gpurun_out/${TAG}_pair1.log; rc=$?; echo "bench pair=1 rc=$rc"
grep "ef curve" gpurun_out/${TAG}_pair1.log
if [ $rc -ne 0 ]; then tail -5 gpurun_out/${TAG}_pair1.log; exit 0; fi
python - <<P
import json
r=json.load(open("gpurun_out/${TAG}_pair1.json"))
print("pair=1: value %.2fM single %.2fM e2e %.2fM sync %.2fM beam %.4f ms frac %.3f" % (r["value"]/1e6, r["single_stream"]["value"]/1e6, r["e2e"]["value"]/1e6, r["e2e"]["sync"]["value"]/1e6, r["roofline"]["kernel_ms"], r["roofline"]["frac"]))
P
GBDR_BEAM_PAIR=1 timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q > gpurun_out/${TAG}_pytest_full.txt 2>&1; echo "pytest fullsize rc=$?"; tail -3 gpurun_out/${TAG}_pytest_full.txt
GBDR_BEAM_PAIR=1 timeout 200 python bench.py --workload deep1m --steps 10 --warmup 3 --no-cpu-baseline --efs 40,80,100,120 > gpurun_out/${TAG}_deep1m.json 2> gpurun_out/${TAG}_deep1m.log; grep "ef curve" gpurun_out/${TAG}_deep1m.log
GBDR_BEAM_PAIR=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:beam_search -c 1 -f -o gpurun_out/${TAG}_beam python bench.py --steps 1 --warmup 0 --ef 53 --no-cpu-baseline --no-ef-curve > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu rc=$?"
