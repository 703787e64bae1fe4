# usage: bash scripts/gpu_scale.sh TAG N — N GPUs of one box: the bench under torchrun (replicated index) and the sharded arm
TAG=${1:-scale}; N=${2:-8}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 40 --warmup 3 --no-ef-curve > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.log; echo "bench n$N rc=$?"
tail -2 gpurun_out/${TAG}_bench_n$N.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --workload deep-sharded --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_sharded_n$N.json 2> gpurun_out/${TAG}_sharded_n$N.log; echo "sharded n$N rc=$?"
tail -2 gpurun_out/${TAG}_sharded_n$N.log
python - <<PY
import json
for f in ("bench_n$N", "sharded_n$N"):
    try:
        j = json.loads([l for l in open(f"gpurun_out/${TAG}_{f}.json") if l.startswith("{")][-1])
        print(f, "value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "ms/step", round(j["ms_per_step"], 3), "ef", j["config"]["ef"], "recall", j["config"]["recall_at_1"], j.get("sharded_knn_build"))
    except Exception as e:
        print(f, "failed", e)
PY
