# usage: bash scripts/gpu_r1z.sh TAG — wave quantisation of one batch: resident CTAs per SM capped (single-stream leg)
TAG=${1:-r1z}
mkdir -p gpurun_out
for B in 4 3; do for W in 8 7 6; do
GBDR_BEAM_BPS=$B GBDR_BEAM_WPB=$W timeout 300 python bench.py --steps 40 --warmup 3 --ef 53 --no-cpu-baseline --no-ef-curve > gpurun_out/${TAG}_b${B}_w${W}.json 2> gpurun_out/${TAG}_b${B}_w${W}.log; echo "B=$B W=$W rc=$?"
done; done
grep "hops per query" gpurun_out/${TAG}_b4_w8.log
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_b*.json')):
    try:
        j=json.load(open(f)); print(f, 'value', round(j['value']), 'single', round(j['single_stream']['value']), 'e2e', round(j['e2e']['value']), 'sync', round(j['e2e']['sync']['value']), 'kms', round(j['roofline']['kernel_ms'],4))
    except Exception as e: print(f, 'ERR', e)
PY
