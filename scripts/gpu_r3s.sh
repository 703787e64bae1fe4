# usage: bash scripts/gpu_r3s.sh TAG — compute-sanitizer over smoke() (memcheck, racecheck) and over the parts of the suite this
# round added or changed: visited-set formats / spill / ties, kNN redo path, knn_cut, HBM-resident build chain
TAG=${1:-r3s}
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_racecheck_smoke.txt 2>&1; echo "racecheck smoke rc=$?"
tail -3 gpurun_out/${TAG}_racecheck_smoke.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_search.py -m gpu -x -q --timeout 800 -k "formats or overflow or ties or spill" > gpurun_out/${TAG}_memcheck_search.txt 2>&1; echo "memcheck search rc=$?"
tail -3 gpurun_out/${TAG}_memcheck_search.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_build_ops.py -m gpu -x -q --timeout 800 -k "knn_cut or scattered or gd_prune_matches or massive_ties" > gpurun_out/${TAG}_memcheck_build.txt 2>&1; echo "memcheck build rc=$?"
tail -3 gpurun_out/${TAG}_memcheck_build.txt
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_search.py -m gpu -x -q --timeout 500 -k "ties" > gpurun_out/${TAG}_racecheck_search.txt 2>&1; echo "racecheck search rc=$?"
tail -3 gpurun_out/${TAG}_racecheck_search.txt
