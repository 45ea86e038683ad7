python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r2zj_tests.log
export PRE=150 K=40
python tools/steady_diag.py acorn 4096 2>&1 | tee gpurun_out/r2zj_steady.log
