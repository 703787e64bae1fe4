# usage: bash scripts/gpu_r4t.sh TAG — initcheck over smoke() after the forward-list fill; GD tests
TAG=${1:-r4t}
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool initcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_initcheck_smoke.txt 2>&1; echo "initcheck smoke rc=$?"
grep -E "ERROR SUMMARY|smoke ok" gpurun_out/${TAG}_initcheck_smoke.txt
grep -E "Uninitialized" -B1 -A2 gpurun_out/${TAG}_initcheck_smoke.txt | head -24
timeout 600 python -m pytest tests/test_gpu_build_ops.py tests/test_gpu_golden.py tests/test_gpu_group.py -q -m gpu 2>&1 | tail -2
