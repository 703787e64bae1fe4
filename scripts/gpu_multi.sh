# usage: bash scripts/gpu_multi.sh TAG N   — GPU tests (incl. NCCL test) and the bench under torchrun on N GPUs
TAG=${1:-mg}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.txt
tail -8 gpurun_out/${TAG}_pytest.txt
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.log; echo "bench rc=$?"
python - <<PY
import json
for n in (1, $N):
    try:
        j = json.loads([l for l in open(f"gpurun_out/${TAG}_bench_n{n}.json") if l.startswith("{")][-1])
        print(n, "GPUs value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "ms/step", j["ms_per_step"], "ef", j["config"]["ef"])
    except Exception as e:
        print(n, "failed", e)
PY
tail -5 gpurun_out/${TAG}_bench_n$N.log
