set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_scenarios.py tests/test_visibility.py tests/test_stage12_reference.py -m gpu -q -x 2>&1 | tail -15
python scripts/profile_cycle_host.py 5 > gpurun_out/r2u_cycle_host.txt 2>&1; head -3 gpurun_out/r2u_cycle_host.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
