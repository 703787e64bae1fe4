# usage: bash scripts/gpu_sanitize.sh TAG — compute-sanitizer over smoke() and a slice of the search tests; reference arm check
TAG=${1:-san}
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_memcheck_smoke.txt 2>&1; echo "memcheck smoke rc=$?"
tail -4 gpurun_out/${TAG}_memcheck_smoke.txt
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_racecheck_smoke.txt 2>&1; echo "racecheck smoke rc=$?"
tail -4 gpurun_out/${TAG}_racecheck_smoke.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_search.py -m gpu -x -q --timeout 600 -k "formats or overflow or ties or spill" > gpurun_out/${TAG}_memcheck_search.txt 2>&1; echo "memcheck search rc=$?"
tail -4 gpurun_out/${TAG}_memcheck_search.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_reference.json 2> gpurun_out/${TAG}_reference.log; echo "reference rc=$?"
cut -c1-600 gpurun_out/${TAG}_reference.json
