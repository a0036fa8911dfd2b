set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_metric_gpu.py -m gpu -q -x 2>&1 | tail -3
FO_EXACT_DCE=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-stages --no-latency > gpurun_out/r2_ties.json 2> gpurun_out/r2_ties.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r2_ties.json').read().strip().splitlines()[-1])
print('RESULT ties', '%.4g'%d['value'], '%.3f'%d['ms_per_step'])
PY
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-stages --no-latency > gpurun_out/r2_noties.json 2> gpurun_out/r2_noties.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r2_noties.json').read().strip().splitlines()[-1])
print('RESULT default', '%.4g'%d['value'], '%.3f'%d['ms_per_step'])
PY
python tests/tools/parity_campaign.py > gpurun_out/parity_campaign_r2.json 2> gpurun_out/parity_campaign_r2.err; tail -c 420 gpurun_out/parity_campaign_r2.json
