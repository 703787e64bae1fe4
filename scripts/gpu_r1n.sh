# usage: bash scripts/gpu_r1n.sh TAG — GPU tests, A/B of the speculative row prefetch, launch list, ncu of the beam kernel
TAG=${1:-r1n}
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.txt
tail -8 gpurun_out/${TAG}_pytest.txt
GBDR_BEAM_PF_ROWS=0 timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_pf0.json 2> gpurun_out/${TAG}_bench_pf0.log; echo "bench rc=$?"
EF=$(python -c "import json;print(json.load(open('gpurun_out/${TAG}_bench_pf0.json'))['config']['ef'])")
for i in 1 2; do
GBDR_BEAM_PF_ROWS=1 timeout 300 python bench.py --steps 20 --warmup 3 --ef $EF --no-cpu-baseline > gpurun_out/${TAG}_bench_pf1_$i.json 2> gpurun_out/${TAG}_bench_pf1.log; echo "bench rc=$?"
GBDR_BEAM_PF_ROWS=0 timeout 300 python bench.py --steps 20 --warmup 3 --ef $EF --no-cpu-baseline > gpurun_out/${TAG}_bench_pf0_$i.json 2>> gpurun_out/${TAG}_bench_pf0.log; echo "bench rc=$?"
done
GBDR_BEAM_PF_ROWS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:beam_search -c 1 -f -o gpurun_out/${TAG}_beam_pf1 python bench.py --steps 1 --warmup 0 --ef $EF --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_bench_pf*.json')):
    try:
        j=json.load(open(f)); print(f, j['value'], j['e2e']['value'], j['roofline']['kernel_ms'], j['config']['ef'], j['config']['recall_at_1'])
    except Exception as e: print(f, 'ERR', e)
PY
