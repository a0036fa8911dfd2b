set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_scenarios.py tests/test_visibility.py tests/test_stage12_reference.py -m gpu -q -x 2>&1 | tail -25
python scripts/profile_cycle_host.py 5 > gpurun_out/r2m_cycle_host.txt 2>&1; head -3 gpurun_out/r2m_cycle_host.txt
