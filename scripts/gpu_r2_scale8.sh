set -x
mkdir -p gpurun_out
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r2_bench_g$n.json 2> gpurun_out/r2_bench_g$n.err; tail -c 1500 gpurun_out/r2_bench_g$n.json
done
