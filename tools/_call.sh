set -x
timeout 600 python -m pytest tests/test_policy.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2y_policy_tests.log
timeout 300 python tools/policy_bench.py 4096 16384 2>&1 | tee gpurun_out/r2y_policy_bench.log
GRP_CONV1_ISSUER=1 timeout 300 python tools/policy_bench.py 4096 16384 2>&1 | tee -a gpurun_out/r2y_policy_bench.log
