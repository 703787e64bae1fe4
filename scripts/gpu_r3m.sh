# usage: bash scripts/gpu_r3m.sh TAG N — the two bench arms on N GPUs exactly as the driver launches them (torchrun), timed
TAG=${1:-r3m}; N=${2:-8}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt; nproc >> gpurun_out/${TAG}_gpus.txt; free -g | head -2 >> gpurun_out/${TAG}_gpus.txt
( time timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus $N --steps 20 --warmup 3 ) > gpurun_out/${TAG}_ref_n$N.json 2> gpurun_out/${TAG}_ref_n$N.log; echo "ref rc=$?"
grep real gpurun_out/${TAG}_ref_n$N.log
( time timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 3 ) > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.log; echo "bench rc=$?"
grep -E "real|shard|operating" gpurun_out/${TAG}_bench_n$N.log | tail -12
python - <<PY
import json
for f in ("ref_n$N", "bench_n$N"):
    try:
        j = json.loads([l for l in open(f"gpurun_out/${TAG}_{f}.json") if l.startswith("{")][-1])
        print(f, "value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "cores", (j.get("cpu_baseline") or {}).get("cores"))
        for k in ("strong", "build_sharded", "sharded", "group"):
            if k in j: print("  ", k, json.dumps(j[k])[:900])
    except Exception as e:
        print(f, "failed", e)
PY
