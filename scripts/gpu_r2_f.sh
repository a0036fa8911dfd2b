set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/r2f_pytest.log 2>&1; tail -8 gpurun_out/r2f_pytest.log
python scripts/bench_visibility.py 10000 2>&1 | tail -1
python scripts/bench_visibility.py 10000 ring 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -c 3000 gpurun_out/r2f_bench.json; tail -3 gpurun_out/r2f_bench.err
