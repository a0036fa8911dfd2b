set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/r2e_pytest.log 2>&1; tail -15 gpurun_out/r2e_pytest.log
python scripts/bench_detail.py 2>&1 | tail -1
python scripts/bench_detail.py 20000 256 51 2>&1 | tail -1
python scripts/bench_detail.py 100000 32 31 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fo_metric_detail -s 1 -c 1 -f -o gpurun_out/prof_r2e_detail python scripts/bench_detail.py > gpurun_out/r2e_prof.log 2>&1; tail -2 gpurun_out/r2e_prof.log
