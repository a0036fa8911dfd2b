# quick check of the metric kernels: parity tests + bench (no CPU baseline)
set -x
mkdir -p gpurun_out
TAG=${1:-q}
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print('RESULT', d['value'], d['ms_per_step'], d['e2e']['value'], d['latency']['p50_us'], d['latency']['graph_p50_us'], d['stages']['detail_kernel']['kernel_ms'])
PY
