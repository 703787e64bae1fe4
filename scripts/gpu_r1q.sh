# usage: bash scripts/gpu_r1q.sh TAG — full check of the current build: GPU tests, bench (3 batches in flight, CPU baseline,
# fresh workload build with GD stage timing), launch list, ncu --set full of the beam kernel, the other single-GPU shapes
TAG=${1:-r1q}
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.txt
tail -5 gpurun_out/${TAG}_pytest.txt
GBDR_GD_TIMING=1 timeout 600 python bench.py --steps 60 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.log; echo "bench rc=$?"
grep -v "ef=" gpurun_out/${TAG}_bench.log | tail -25
EF=$(python -c "import json;print(json.load(open('gpurun_out/${TAG}_bench.json'))['config']['ef'])")
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.log; echo "reference rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --ef $EF --no-cpu-baseline --in-flight 1 > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:beam_search -c 1 -f -o gpurun_out/${TAG}_beam python bench.py --steps 1 --warmup 0 --ef $EF --no-cpu-baseline --in-flight 1 > gpurun_out/${TAG}_ncu_full.log 2>&1
for W in deep1m gist1m; do
  timeout 600 python bench.py --workload $W --steps 20 --warmup 3 > gpurun_out/${TAG}_${W}.json 2> gpurun_out/${TAG}_${W}.log; echo "$W rc=$?"
  tail -3 gpurun_out/${TAG}_${W}.log
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_*.json')):
    try:
        j=json.loads([l for l in open(f) if l.startswith("{")][-1])
        if j.get('impl')=='reference': print(f, 'reference', round(j['value']), j['cpu_baseline']['cores']); continue
        print(f, 'value', round(j['value']), 'single', round(j['single_stream']['value']), 'e2e', round(j['e2e']['value']), 'sync', round(j['e2e']['sync']['value']), 'kms', j['roofline']['kernel_ms'], 'frac', round(j['roofline']['frac'],3), 'ef', j['config']['ef'], 'cpu', j.get('cpu_baseline',{}).get('value'), j.get('build'))
    except Exception as e: print(f, 'ERR', e)
PY
