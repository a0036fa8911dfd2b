# round 2, call A: correctness of the rebuilt detail kernel + sorted agent table, first numbers
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=12 -x -k "not scenario" > gpurun_out/r2a_pytest.log 2>&1; tail -15 gpurun_out/r2a_pytest.log
timeout 600 python -m pytest tests/test_scenarios.py -m gpu -q --maxfail=5 > gpurun_out/r2a_pytest_scen.log 2>&1; tail -5 gpurun_out/r2a_pytest_scen.log
bash scripts/ab_variants.sh each 3 python scripts/bench_detail.py 2>&1 | grep -v "^+" | tee gpurun_out/r2a_detail_variants.log
python scripts/bench_detail.py 20000 256 51 2>&1 | tail -1
python scripts/bench_detail.py 100000 32 31 2>&1 | tail -1
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-stages > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 600 gpurun_out/r2a_bench.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fo_metric_detail -s 1 -c 1 -f -o gpurun_out/prof_r2a_detail python scripts/bench_detail.py > gpurun_out/r2a_prof.log 2>&1; tail -2 gpurun_out/r2a_prof.log
