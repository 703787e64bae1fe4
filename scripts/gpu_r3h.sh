# usage: bash scripts/gpu_r3h.sh TAG N — N-GPU box: group / NCCL tests, the drop-in final_test over GBDR_DEVICES, then the two
# bench arms exactly as the driver launches them (torchrun), timed
TAG=${1:-r3h}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_multigpu.py tests/test_host_dropin.py -q -m gpu > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.txt
tail -6 gpurun_out/${TAG}_pytest.txt
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus $N --steps 20 --warmup 3 ) > gpurun_out/${TAG}_ref_n$N.json 2> gpurun_out/${TAG}_ref_n$N.log; echo "ref rc=$?"
grep real gpurun_out/${TAG}_ref_n$N.log
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 3 ) > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.log; echo "bench rc=$?"
grep -E "real|shard|operating" gpurun_out/${TAG}_bench_n$N.log | tail -12
python - <<PY
import json
for f in ("ref_n$N", "bench_n$N"):
    try:
        j = json.loads([l for l in open(f"gpurun_out/${TAG}_{f}.json") if l.startswith("{")][-1])
        print(f, "value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "cores", (j.get("cpu_baseline") or {}).get("cores"))
        for k in ("strong", "build_sharded", "sharded", "group"):
            if k in j: print("  ", k, json.dumps(j[k])[:700])
    except Exception as e:
        print(f, "failed", e)
PY
