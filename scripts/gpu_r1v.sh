# usage: bash scripts/gpu_r1v.sh TAG N — sharded bench on 1 and N GPUs
TAG=${1:-r1v}; N=${2:-2}
mkdir -p gpurun_out
timeout 600 python bench.py --workload deep-sharded --steps 10 --warmup 3 > gpurun_out/${TAG}_sharded_n1.json 2> gpurun_out/${TAG}_sharded_n1.log; echo "sharded n1 rc=$?"
tail -3 gpurun_out/${TAG}_sharded_n1.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --workload deep-sharded --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_sharded_n$N.json 2> gpurun_out/${TAG}_sharded_n$N.log; echo "sharded n$N rc=$?"
tail -3 gpurun_out/${TAG}_sharded_n$N.log
python - <<PY
import json
for f in ("sharded_n1", "sharded_n$N"):
    try:
        j = json.loads([l for l in open(f"gpurun_out/${TAG}_{f}.json") if l.startswith("{")][-1])
        print(f, "value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "ms/step", round(j["ms_per_step"], 3), "ef", j["config"]["ef"], "recall", j["config"]["recall_at_1"], j.get("sharded_knn_build"), j["roofline"]["kernel_ms"], j["roofline"]["other_kernels_ms"])
    except Exception as e:
        print(f, "failed", e)
PY
