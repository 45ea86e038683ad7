timeout 600 python -m pytest tests/test_policy.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python tools/policy_bench.py 4096 16384 2>&1
GRP_EVENTS=1 timeout 300 python tools/policy_bench.py 4096 2>&1 | tail -2
