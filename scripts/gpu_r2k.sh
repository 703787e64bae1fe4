# usage: bash scripts/gpu_r2k.sh TAG — compute-sanitizer over the final build's search tests (generalised tag tables, joint plan)
TAG=${1:-r2k}
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_search.py -m gpu -x -q --timeout 450 -k "formats or overflow or ties or spill or second_graph or views or in_flight or row_widths or wide" > gpurun_out/${TAG}_memcheck_search.txt 2>&1; echo "memcheck search rc=$?"
tail -3 gpurun_out/${TAG}_memcheck_search.txt
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_search.py -m gpu -x -q --timeout 380 -k "row_widths or wide or second_graph_requires or in_flight" > gpurun_out/${TAG}_racecheck_search.txt 2>&1; echo "racecheck search rc=$?"
tail -3 gpurun_out/${TAG}_racecheck_search.txt
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_build_ops.py -m gpu -x -q -k "projection" > gpurun_out/${TAG}_memcheck_proj.txt 2>&1; echo "memcheck projection rc=$?"
tail -3 gpurun_out/${TAG}_memcheck_proj.txt
