# usage: bash scripts/gpu_r1o.sh TAG — GPU tests, then the bench with 2 batches in flight at several CTA sizes
TAG=${1:-r1o}
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.txt
tail -8 gpurun_out/${TAG}_pytest.txt
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_w10.json 2> gpurun_out/${TAG}_bench_w10.log; echo "bench rc=$?"
tail -3 gpurun_out/${TAG}_bench_w10.log
EF=$(python -c "import json;print(json.load(open('gpurun_out/${TAG}_bench_w10.json'))['config']['ef'])")
for W in 5 3 6 10; do
GBDR_BEAM_WPB=$W timeout 300 python bench.py --steps 40 --warmup 3 --ef $EF --no-cpu-baseline > gpurun_out/${TAG}_bench_w${W}b.json 2> gpurun_out/${TAG}_bench_w${W}b.log; echo "bench W=$W rc=$?"
done
GBDR_BEAM_WPB=5 timeout 300 python bench.py --steps 40 --warmup 3 --ef $EF --no-cpu-baseline --in-flight 3 > gpurun_out/${TAG}_bench_w5_f3.json 2> gpurun_out/${TAG}_bench_w5_f3.log; echo "bench f3 rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_bench_*.json')):
    try:
        j=json.load(open(f)); print(f, 'value', round(j['value']), 'single', round(j['single_stream']['value']), 'e2e', round(j['e2e']['value']), 'sync', round(j['e2e']['sync']['value']), 'kms', j['roofline']['kernel_ms'])
    except Exception as e: print(f, 'ERR', e)
PY
