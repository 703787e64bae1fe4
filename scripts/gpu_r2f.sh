# usage: bash scripts/gpu_r2f.sh TAG — table load vs resident warps at ef 53: 4x8 (256 buckets), 3x10 (288), 4x7 (320), 3x9 (341)
TAG=${1:-r2f}
mkdir -p gpurun_out
for W in 8 10 7 9; do
GBDR_BEAM_WPB=$W timeout 300 python bench.py --steps 40 --warmup 3 --ef 53 --no-cpu-baseline --no-ef-curve > gpurun_out/${TAG}_w${W}.json 2> gpurun_out/${TAG}_w${W}.log; echo "W=$W rc=$?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_w*.json')):
    try:
        j=json.load(open(f)); print(f, 'value', round(j['value']), 'single', round(j['single_stream']['value']), 'e2e', round(j['e2e']['value']), 'sync', round(j['e2e']['sync']['value']), 'kms', round(j['roofline']['kernel_ms'],4))
    except Exception as e: print(f, 'ERR', e)
PY
