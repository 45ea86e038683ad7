python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/r2zh_tests.log
python -m pytest tests/test_env_parity.py -m gpu -q -s 2>&1 | grep -E "flipped agent"
