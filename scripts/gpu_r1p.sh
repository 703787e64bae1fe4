# usage: bash scripts/gpu_r1p.sh TAG — GPU tests with the 7 KB/warp layout, bench at in-flight 2/3/4 x CTA size 8/10, ncu
TAG=${1:-r1p}
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.txt
tail -5 gpurun_out/${TAG}_pytest.txt
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_auto_f2.json 2> gpurun_out/${TAG}_bench_auto_f2.log; echo "bench rc=$?"
tail -2 gpurun_out/${TAG}_bench_auto_f2.log
EF=$(python -c "import json;print(json.load(open('gpurun_out/${TAG}_bench_auto_f2.json'))['config']['ef'])")
for F in 2 3 4; do for W in 8 10; do
GBDR_BEAM_WPB=$W timeout 300 python bench.py --steps 60 --warmup 3 --ef $EF --no-cpu-baseline --in-flight $F > gpurun_out/${TAG}_bench_w${W}_f${F}.json 2> gpurun_out/${TAG}_bench_w${W}_f${F}.log; echo "bench W=$W F=$F rc=$?"
done; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:beam_search -c 1 -f -o gpurun_out/${TAG}_beam python bench.py --steps 1 --warmup 0 --ef $EF --no-cpu-baseline --in-flight 1 > gpurun_out/${TAG}_ncu_full.log 2>&1
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_bench_*.json')):
    try:
        j=json.load(open(f)); print(f, 'value', round(j['value']), 'single', round(j['single_stream']['value']), 'e2e', round(j['e2e']['value']), 'sync', round(j['e2e']['sync']['value']), 'kms', j['roofline']['kernel_ms'])
    except Exception as e: print(f, 'ERR', e)
PY
