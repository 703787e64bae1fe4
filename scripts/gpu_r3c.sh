# usage: bash scripts/gpu_r3c.sh TAG — the whole GPU test suite (no -x), the default bench line, the reference arm
TAG=${1:-r3c}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.txt
tail -15 gpurun_out/${TAG}_pytest.txt
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.log; echo "bench rc=$?"
tail -3 gpurun_out/${TAG}_bench.log
timeout 600 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.log; echo "ref rc=$?"
cat gpurun_out/${TAG}_bench_reference.json
