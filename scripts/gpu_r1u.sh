# usage: bash scripts/gpu_r1u.sh TAG N — on N GPUs of one box: the NCCL test, the bench under torchrun (replicated index,
# weak scaling) and the sharded bench (rows sharded, NCCL all-gather + merge; row-block sharded kNN build)
TAG=${1:-r1u}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt
timeout 600 python -m pytest tests/test_gpu_multigpu.py -m gpu -x -q > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.txt
tail -4 gpurun_out/${TAG}_pytest.txt
timeout 400 python bench.py --steps 60 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.log; echo "bench n1 rc=$?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 60 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.log; echo "bench n$N rc=$?"
tail -2 gpurun_out/${TAG}_bench_n$N.log
timeout 600 python bench.py --workload deep-sharded --steps 10 --warmup 3 > gpurun_out/${TAG}_sharded_n1.json 2> gpurun_out/${TAG}_sharded_n1.log; echo "sharded n1 rc=$?"
tail -3 gpurun_out/${TAG}_sharded_n1.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --workload deep-sharded --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_sharded_n$N.json 2> gpurun_out/${TAG}_sharded_n$N.log; echo "sharded n$N rc=$?"
tail -3 gpurun_out/${TAG}_sharded_n$N.log
python - <<PY
import json
for f in ("bench_n1", "bench_n$N", "sharded_n1", "sharded_n$N"):
    try:
        j = json.loads([l for l in open(f"gpurun_out/${TAG}_{f}.json") if l.startswith("{")][-1])
        print(f, "value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "ms/step", round(j["ms_per_step"], 3), "ef", j["config"]["ef"], "recall", j["config"]["recall_at_1"], j.get("sharded_knn_build"))
    except Exception as e:
        print(f, "failed", e)
PY
