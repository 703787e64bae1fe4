# usage: bash scripts/gpu_check.sh TAG [pytest-args...]   — GPU tests, bench, launch list, ncu of the beam kernel
TAG=${1:-run}; shift
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q "$@" > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.txt
tail -15 gpurun_out/${TAG}_pytest.txt
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.log; echo "bench rc=$?" >> gpurun_out/${TAG}_bench.log
EF=$(python -c "import json;print(json.load(open('gpurun_out/${TAG}_bench.json'))['config']['ef'])")
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --ef $EF --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:beam_search -c 1 -f -o gpurun_out/${TAG}_beam python bench.py --steps 1 --warmup 0 --ef $EF --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
cat gpurun_out/${TAG}_bench.json
tail -3 gpurun_out/${TAG}_bench.log
