set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r1b_gpu.txt
nproc >> gpurun_out/r1b_gpu.txt; free -g >> gpurun_out/r1b_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1b_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r1b_pytest.txt
timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/r1b_bench.json 2> gpurun_out/r1b_bench.log; echo "bench rc=$?" >> gpurun_out/r1b_bench.log
EF=$(python -c "import json;print(json.load(open('gpurun_out/r1b_bench.json'))['config']['ef'])")
echo EF=$EF
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1b_launches.csv python bench.py --steps 2 --warmup 1 --ef $EF --no-cpu-baseline > gpurun_out/r1b_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:beam_search -c 1 -f -o gpurun_out/r1b_beam python bench.py --steps 1 --warmup 0 --ef $EF --no-cpu-baseline > gpurun_out/r1b_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rerank -c 1 -f -o gpurun_out/r1b_rerank python bench.py --steps 1 --warmup 0 --ef $EF --no-cpu-baseline >> gpurun_out/r1b_ncu_full.log 2>&1
cat gpurun_out/r1b_bench.json
tail -5 gpurun_out/r1b_pytest.txt
