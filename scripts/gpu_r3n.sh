# usage: bash scripts/gpu_r3n.sh TAG — 2-GPU box: group + NCCL + drop-in tests in ONE process (the order in which torch's import
# once met the system libnccl), incl. the drop-in binaries over GBDR_DEVICES; GIST-1M with the automatic in-flight depth
TAG=${1:-r3n}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_multigpu.py tests/test_host_dropin.py -q -m gpu > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.txt
tail -6 gpurun_out/${TAG}_pytest.txt
timeout 300 python bench.py --workload gist1m --steps 20 --warmup 3 --no-cpu-baseline --no-ef-curve > gpurun_out/${TAG}_gist1m.json 2> gpurun_out/${TAG}_gist1m.log
python - <<PY
import json
r=json.load(open("gpurun_out/${TAG}_gist1m.json"))
print("gist1m: ef %d in-flight %d value %.2fM single %.2fM e2e %.2fM beam %.3f ms frac %.3f build %s" % (r["config"]["ef"], r["config"]["batches_in_flight"], r["value"]/1e6, r["single_stream"]["value"]/1e6, r["e2e"]["value"]/1e6, r["roofline"]["kernel_ms"], r["roofline"]["frac"], r["build"]))
PY
