set -x
python scripts/profile_cycle_host.py 3 2>&1 | tail -75
timeout 600 python -m pytest tests/test_visibility.py -m gpu -q 2>&1 | tail -3
