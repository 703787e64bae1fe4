# usage: bash scripts/gpu_r1r.sh TAG — compute-sanitizer (memcheck, racecheck) over smoke() and the search tests that
# exercise the 7 KB/warp beam-kernel layout, the second graph and views
TAG=${1:-r1r}
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_memcheck_smoke.txt 2>&1; echo "memcheck smoke rc=$?"
tail -3 gpurun_out/${TAG}_memcheck_smoke.txt
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_racecheck_smoke.txt 2>&1; echo "racecheck smoke rc=$?"
tail -3 gpurun_out/${TAG}_racecheck_smoke.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_search.py -m gpu -x -q --timeout 800 -k "formats or overflow or ties or spill or second_graph or views or in_flight or row_widths" > gpurun_out/${TAG}_memcheck_search.txt 2>&1; echo "memcheck search rc=$?"
tail -3 gpurun_out/${TAG}_memcheck_search.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_search.py -m gpu -x -q --timeout 800 -k "formats or second_graph or row_widths" > gpurun_out/${TAG}_racecheck_search.txt 2>&1; echo "racecheck search rc=$?"
tail -3 gpurun_out/${TAG}_racecheck_search.txt
