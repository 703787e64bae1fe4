# usage: bash scripts/gpu_r1t.sh TAG — GPU tests + bench (ef curve) with 16-bit visited tags at every list capacity
TAG=${1:-r1t}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.txt
tail -5 gpurun_out/${TAG}_pytest.txt
timeout 400 python bench.py --steps 60 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.log; echo "bench rc=$?"
tail -2 gpurun_out/${TAG}_bench.log
GBDR_BEAM_VIS16=0 timeout 400 python bench.py --steps 20 --warmup 3 --ef 53 --no-cpu-baseline > gpurun_out/${TAG}_bench_vis32.json 2> gpurun_out/${TAG}_bench_vis32.log; echo "bench rc=$?"
tail -1 gpurun_out/${TAG}_bench_vis32.log
for W in deep1m gist1m; do
  timeout 600 python bench.py --workload $W --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_${W}.json 2> gpurun_out/${TAG}_${W}.log; echo "$W rc=$?"
  tail -1 gpurun_out/${TAG}_${W}.log
done
