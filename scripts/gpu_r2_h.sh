set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 --no-stages --no-latency > gpurun_out/r2h_bench_g2.json 2> gpurun_out/r2h_bench_g2.err; tail -c 1500 gpurun_out/r2h_bench_g2.json; tail -5 gpurun_out/r2h_bench_g2.err
python bench.py --gpus 1 --steps 5 --warmup 3 --no-stages --no-latency --no-cpu-baseline > gpurun_out/r2h_bench_g1.json 2> gpurun_out/r2h_bench_g1.err; python -c "
import json
for n in (1,2):
    d=json.load(open('gpurun_out/r2h_bench_g%d.json'%n)); print(n, 'ms %.2f value %.4g e2e %.4g' % (d['ms_per_step'], d['value'], d['e2e']['value']))"
