# usage: bash scripts/gpu_dbg.sh TAG  — search tests with tight timeouts, then a fixed-ef bench
TAG=${1:-dbg}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_search.py -m gpu -x -q --timeout 60 > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.txt
tail -30 gpurun_out/${TAG}_pytest.txt
timeout 200 python bench.py --steps 5 --warmup 3 --ef 8 --no-cpu-baseline > gpurun_out/${TAG}_bench8.json 2> gpurun_out/${TAG}_bench8.log; echo "bench rc=$?"
tail -5 gpurun_out/${TAG}_bench8.log; cat gpurun_out/${TAG}_bench8.json
