set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/r2n_pytest.log
python scripts/profile_cycle_host.py 5 > gpurun_out/r2n_cycle_host.txt 2>&1; head -3 gpurun_out/r2n_cycle_host.txt
timeout 600 python bench.py > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; tail -c 3000 gpurun_out/r2n_bench.json
