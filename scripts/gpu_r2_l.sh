set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_visibility.py -m gpu -q 2>&1 | tail -3
python scripts/bench_visibility.py 10000 2>&1 | tail -1 | tee gpurun_out/r2l_vis.json
python scripts/bench_visibility.py 10000 ring 2>&1 | tail -1 | tee gpurun_out/r2l_vis_ring.json
python scripts/profile_cycle_host.py 5 > gpurun_out/r2l_cycle_host.txt 2>&1; head -3 gpurun_out/r2l_cycle_host.txt
