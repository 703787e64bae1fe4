# usage: bash scripts/gpu_r4s.sh TAG — compute-sanitizer synccheck and initcheck over smoke() and a slice of the search tests (final build)
TAG=${1:-r4s}
mkdir -p gpurun_out
for tool in synccheck initcheck; do
timeout 600 compute-sanitizer --tool $tool --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_${tool}_smoke.txt 2>&1; echo "$tool smoke rc=$?"
tail -2 gpurun_out/${TAG}_${tool}_smoke.txt
done
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest tests/test_gpu_search.py -m gpu -x -q --timeout 800 -k "ties or overflow or replay" > gpurun_out/${TAG}_synccheck_search.txt 2>&1; echo "synccheck search rc=$?"
tail -3 gpurun_out/${TAG}_synccheck_search.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_search.py -m gpu -x -q --timeout 800 -k "replay" > gpurun_out/${TAG}_memcheck_replay.txt 2>&1; echo "memcheck replay rc=$?"
tail -3 gpurun_out/${TAG}_memcheck_replay.txt
