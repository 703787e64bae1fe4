# usage: bash scripts/gpu_r2h.sh TAG — batches in flight 2/3/4/6 at 100 steps (steady state)
TAG=${1:-r2h}
mkdir -p gpurun_out
for F in 3 4 6 2 3 4; do
timeout 300 python bench.py --steps 120 --warmup 3 --ef 53 --no-cpu-baseline --no-ef-curve --in-flight $F > gpurun_out/${TAG}_f${F}.json 2> gpurun_out/${TAG}_f${F}.log; echo "F=$F rc=$?"
python -c "
import json; j=json.load(open('gpurun_out/${TAG}_f${F}.json')); print('F=$F value', round(j['value']), 'e2e', round(j['e2e']['value']), 'single', round(j['single_stream']['value']), 'sync', round(j['e2e']['sync']['value']))"
done
