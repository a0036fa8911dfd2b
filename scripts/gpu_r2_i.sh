set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_metric_gpu.py tests/test_scenarios.py tests/test_visibility.py -m gpu -q --maxfail=10 > gpurun_out/r2i_pytest.log 2>&1; tail -8 gpurun_out/r2i_pytest.log
bash scripts/ab_variants.sh each 4 python scripts/bench_detail.py 2>&1 | grep -v "^+" | tee gpurun_out/r2i_detail_variants.log
python scripts/bench_detail.py 20000 256 51 2>&1 | tail -1
python scripts/profile_cycle_host.py 3 2>&1 | head -3
