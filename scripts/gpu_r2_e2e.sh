set -x
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m pytest tests/test_peer_exchange.py tests/test_metric_gpu.py -m gpu -q -x 2>&1 | tail -3
for mode in "" "--e2e-torch"; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 20 --warmup 3 $mode --no-cpu-baseline --no-stages --no-latency > gpurun_out/r2_e2e_g$N.json 2> gpurun_out/r2_e2e_g$N.err; tail -3 gpurun_out/r2_e2e_g$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_e2e_g$N.json').read().strip().splitlines()[-1])
print('RESULT "$mode"', d['n_gpus'], '%.4g'%d['value'], '%.3f'%d['ms_per_step'], '%.4g'%d['e2e']['value'], '%.3f'%d['e2e']['ms_per_step'], d['e2e']['how'][:60])
PY
done
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-stages --no-latency > gpurun_out/r2_e2e_g1.json 2> gpurun_out/r2_e2e_g1.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r2_e2e_g1.json').read().strip().splitlines()[-1])
print('RESULT g1', '%.4g'%d['value'], '%.3f'%d['ms_per_step'], '%.4g'%d['e2e']['value'], '%.3f'%d['e2e']['ms_per_step'])
PY
