set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2zc_tests.log
python -m pytest tests/test_policy.py -m gpu -q -s -k variants 2>&1 | grep -E "max \|diff|passed|failed"
export PRE=150 K=40
python tools/steady_diag.py acorn 4096 2>&1 | tee gpurun_out/r2zc_steady.log
