# final scaling evidence: the default bench line at 8 / 4 / 2 GPUs (fused exchange) and the NCCL form at 8
set -x
mkdir -p gpurun_out
run() { n=$1; ex=$2; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 20 --warmup 3 --exchange $ex --no-cpu-baseline --no-stages --no-latency > gpurun_out/bench_r2_g${n}_$ex.json 2> gpurun_out/bench_r2_g${n}_$ex.err; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2_g${n}_$ex.json').read().strip().splitlines()[-1])
print('RESULT $ex', d['n_gpus'], '%.4g'%d['value'], '%.3f'%d['ms_per_step'], '%.3f'%d['roofline'].get('kernel_ms'), '%.4g'%d['e2e']['value'], '%.3f'%d['e2e']['ms_per_step'], d['run'].get('exchange'))
PY
}
run 8 peer; run 8 nccl; run 4 peer; run 2 peer
