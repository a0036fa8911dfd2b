set -x
mkdir -p gpurun_out
[ "${2:-test}" = "test" ] && timeout 600 python -m pytest tests/test_peer_exchange.py tests/test_metric_gpu.py -m gpu -q -x 2>&1 | tail -15
N=${1:-2}
for ex in peer nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 20 --warmup 3 --exchange $ex --no-cpu-baseline --no-stages --no-latency > gpurun_out/r2_peer_${ex}_g$N.json 2> gpurun_out/r2_peer_${ex}_g$N.err; tail -3 gpurun_out/r2_peer_${ex}_g$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_peer_${ex}_g$N.json').read().strip().splitlines()[-1])
print('RESULT $ex', d['n_gpus'], d['value'], d['ms_per_step'], d['roofline'].get('kernel_ms'), d['e2e']['value'], d['e2e']['ms_per_step'], d['run'].get('exchange'))
PY
done
