set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/r2j_pytest.log 2>&1; tail -8 gpurun_out/r2j_pytest.log
timeout 900 python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; tail -c 6000 gpurun_out/r2j_bench.json
