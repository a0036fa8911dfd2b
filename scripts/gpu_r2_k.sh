set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_visibility.py tests/test_scenarios.py tests/test_stage12_reference.py -m gpu -q 2>&1 | tail -5
python scripts/bench_visibility.py 10000 2>&1 | tail -1 | tee gpurun_out/r2k_vis.json
python scripts/bench_visibility.py 10000 ring 2>&1 | tail -1 | tee gpurun_out/r2k_vis_ring.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fo_visibility_kernel -s 4 -c 1 -f -o gpurun_out/r2k_vis python scripts/bench_visibility.py 2000 > gpurun_out/r2k_ncu.log 2>&1; tail -2 gpurun_out/r2k_ncu.log
